#!/bin/bash
# round-2 GPU session Q (1 GPU): call-time window subdivision for short ranges on a large object: parity + timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_dropin.py -q -m gpu -x -k "msm or pippenger or batch or range" 2>&1 | tail -4 | tee gpurun_out/r2q_pytest.txt
out=gpurun_out/r2q_range.txt
: > $out
for v in 1 0; do
  echo "== BBG_MSM_CALL_WINDOW=$v (ranges of a 2^20-point object)" >> $out
  BBG_MSM_CALL_WINDOW=$v DEVBENCH_PLAIN=1 timeout 300 python scripts/devbench.py 12,14,16,17,18,20 "" 2>&1 | grep "^MSM" >> $out
done
cat $out
