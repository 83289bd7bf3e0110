#!/bin/bash
# round-2 GPU session AB (1 GPU): accumulate kernel with 3 CTAs per SM (148 registers, no spills) against the default 4 (128 registers, 12 B spilled)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
out=gpurun_out/r2ab_acc_ctas.txt
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 200 python scripts/devbench.py 20 "" 2>&1 | grep "^MSM" >> $out; }
run BBG_MSM_ACC_CTAS=4
run BBG_MSM_ACC_CTAS=3
run BBG_MSM_ACC_CTAS=3 BBG_MSM_WAVES=4
run BBG_MSM_ACC_CTAS=3 BBG_MSM_WAVES=2
run BBG_MSM_ACC_CTAS=4
cat $out
