#!/bin/bash
# round-2 GPU session U (1 GPU): level-0 segment length of the bucket reduction (any integer now) x worker shape, per MSM size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "msm" 2>&1 | tail -3 | tee gpurun_out/r2u_pytest.txt
BBG_MSM_ELL0=5 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "msm" 2>&1 | tail -3 | tee -a gpurun_out/r2u_pytest.txt
BBG_MSM_ELL0=10 BBG_MSM_SEG0_TEAM=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "msm" 2>&1 | tail -3 | tee -a gpurun_out/r2u_pytest.txt
out=gpurun_out/r2u_ell0.txt
: > $out
run() { lg=$1; shift; echo "== 2^$lg $*" >> $out; env "$@" DEVBENCH_PLAIN=1 timeout 200 python scripts/devbench.py $lg "" 2>&1 | grep "^MSM" >> $out; }
run 20 BBG_X=0
for e in 6 8 10 12 16; do run 20 BBG_MSM_ELL0=$e BBG_MSM_SEG0_TEAM=0; done
run 20 BBG_MSM_ELL0=10 BBG_MSM_SEG0_TEAM=0 BBG_MSM_SEG0_CTAS=4
run 20 BBG_MSM_ELL0=8 BBG_MSM_SEG0_TEAM=0 BBG_MSM_SEG0_CTAS=4
run 18 BBG_X=0
for e in 3 4 5 6 8; do run 18 BBG_MSM_ELL0=$e BBG_MSM_SEG0_TEAM=0; done
for e in 4 6 8 12; do run 18 BBG_MSM_ELL0=$e BBG_MSM_SEG0_TEAM=1; done
run 16 BBG_X=0
for e in 2 3 4 6; do run 16 BBG_MSM_ELL0=$e BBG_MSM_SEG0_TEAM=0; done
for e in 3 4 6 8; do run 16 BBG_MSM_ELL0=$e BBG_MSM_SEG0_TEAM=1; done
cat $out
