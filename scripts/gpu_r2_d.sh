#!/bin/bash
# round-2 GPU session D (1 GPU): tuning sweep of the bucket-reduction shape (segment lengths, team / wide level 0) and
# of the window width now that the tail is shorter
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
out=gpurun_out/r2d_tail_sweep.txt
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 120 python scripts/devbench.py 16,18,20 "" 2>&1 | grep "^MSM" >> $out; }
run BBG_X=0
run BBG_MSM_SEG0_TEAM=1 BBG_MSM_ELL0_LOG2=4 BBG_MSM_ELL1_LOG2=2
run BBG_MSM_SEG0_TEAM=1 BBG_MSM_ELL0_LOG2=2 BBG_MSM_ELL1_LOG2=4
run BBG_MSM_SEG0_TEAM=0 BBG_MSM_ELL0_LOG2=2 BBG_MSM_ELL1_LOG2=3
run BBG_MSM_SEG0_TEAM=0 BBG_MSM_ELL0_LOG2=2 BBG_MSM_ELL1_LOG2=5
run BBG_MSM_SEG0_TEAM=0 BBG_MSM_ELL0_LOG2=3 BBG_MSM_ELL1_LOG2=3
run BBG_MSM_SEG0_TEAM=0 BBG_MSM_ELL0_LOG2=3 BBG_MSM_ELL1_LOG2=4
run BBG_MSM_SEG0_TEAM=0 BBG_MSM_ELL0_LOG2=1 BBG_MSM_ELL1_LOG2=4
run BBG_MSM_SEG0_TEAM=0 BBG_MSM_ELL0_LOG2=1 BBG_MSM_ELL1_LOG2=5
run BBG_MSM_WAVES=2
run BBG_MSM_WAVES=4
for c in 17 18 19; do run BBG_MSM_C=$c; done
cat $out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_cases.py tests/test_gpu_round2.py -q -m gpu -x > gpurun_out/r2d_pytest.txt 2>&1
tail -3 gpurun_out/r2d_pytest.txt
