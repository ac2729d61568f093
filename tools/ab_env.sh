#!/bin/bash
# Same-box comparison of environment settings of ONE build (TMA kernel only): tools/ab_env.sh <configs> VAR=val ...
cfg=$1; shift
for kv in "$@"; do
  echo "== $kv"
  env $kv AB_TMA_ONLY=1 AB_CONFIGS=$cfg timeout 200 python tools/ab_step.py 2>&1 | grep "^| [lsr]"
done
