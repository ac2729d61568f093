#!/bin/bash
# SASS evidence for profiles/: per kernel, the counts of the mnemonics that tell how memory moves
# (UTMALDG / UTMASTG = cp.async.bulk.tensor, UBLKCP = 1-D bulk copy, LDGSTS = cp.async, UTMAPF =
# prefetch.tensormap, SYNCS = mbarrier), then the TMA instructions of k_model_step_pair with context.
LIB=${1:-env_build_b200/csrc/libce2e.so}
echo "# cuobjdump -sass $LIB: memory-movement mnemonics per kernel"
echo
echo "| kernel | UTMALDG | UTMASTG | UBLKCP | UTMAPF | LDGSTS | SYNCS | instructions |"
echo "|---|---|---|---|---|---|---|---|"
cuobjdump -sass "$LIB" | awk '
/Function :/ { if (f != "") print_row(); f = $3; delete c; n = 0 }
/^ +\/\*[0-9a-f]{4}\*\// { n++; for (k in pat) if ($0 ~ pat[k]) c[k]++ }
function print_row() { g = f; sub(/^_ZN[0-9]+_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]+/, "", g); printf "| `%s` | %d | %d | %d | %d | %d | %d | %d |\n", substr(g, 1, 60), c["a"], c["b"], c["c"], c["d"], c["e"], c["f"], n }
BEGIN { pat["a"] = "UTMALDG"; pat["b"] = "UTMASTG"; pat["c"] = "UBLKCP"; pat["d"] = "UTMAPF"; pat["e"] = "LDGSTS"; pat["f"] = "SYNCS" }
END { print_row() }'
echo
echo '## TMA instructions of `k_model_step_pair<FAST=0, BAL=0>`'
echo
echo '```'
cuobjdump -sass "$LIB" | awk '/Function :.*k_model_step_pairILb0ELb0/ {on=1} /Function :/ && !/k_model_step_pairILb0ELb0/ {on=0} on' | grep -E "UTMALDG|UTMASTG|UTMAPF|UBLKCP|UTMACMDFLUSH|SYNCS.ARRIVE.TRANS64 |FENCE.VIEW.ASYNC" | sed 's/ *\/\* 0x[0-9a-f]* \*\///' | head -40
echo '```'
