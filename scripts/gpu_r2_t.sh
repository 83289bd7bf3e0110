#!/bin/bash
# round-2 GPU session T (1 GPU): wide vs team slot-merge level 0: parity, then timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_fullsize.py -q -m gpu -x -k "msm or pippenger or batch" 2>&1 | tail -4 | tee gpurun_out/r2t_pytest.txt
out=gpurun_out/r2t_merge.txt
: > $out
for v in 32768 1000000000 32768 1000000000; do
  echo "== BBG_MSM_MERGE_WIDE_FROM=$v" >> $out
  BBG_MSM_MERGE_WIDE_FROM=$v DEVBENCH_PLAIN=1 timeout 300 python scripts/devbench.py 20 "" 2>&1 | grep "^MSM" >> $out
  BBG_MSM_MERGE_WIDE_FROM=$v timeout 300 python scripts/devbench.py 20 "" 2>&1 | grep "^MSM" >> $out
done
for v in 32768 1000000000; do
  echo "== 2^18 object, BBG_MSM_MERGE_WIDE_FROM=$v" >> $out
  BBG_MSM_MERGE_WIDE_FROM=$v DEVBENCH_PLAIN=1 timeout 300 python scripts/devbench.py 18 "" 2>&1 | grep "^MSM" >> $out
  BBG_MSM_MERGE_WIDE_FROM=$v timeout 300 python scripts/devbench.py 18 "" 2>&1 | grep "^MSM" >> $out
done
cat $out
