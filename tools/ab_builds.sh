#!/bin/bash
# Same-box comparison of several builds of libce2e.so (TMA kernel only): tools/ab_builds.sh <configs> <lib>...
cfg=$1; shift
for lib in "$@"; do
  echo "== $lib"
  CE2E_LIB=$lib AB_TMA_ONLY=1 AB_CONFIGS=$cfg timeout 150 python tools/ab_step.py 2>&1 | grep "^| [lsr]"
done
