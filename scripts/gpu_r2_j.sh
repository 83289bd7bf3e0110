#!/bin/bash
# round-2 GPU session J (1 GPU): MSM cut into bucket-range parts (pipelined tails): parity, then timings for 1 / 2 / 4 parts
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_fullsize.py -q -m gpu -x -k "msm or pippenger or MSM or batch" 2>&1 | tail -5 | tee gpurun_out/r2j_pytest.txt
out=gpurun_out/r2j_msm.txt
: > $out
for parts in 1 2 4; do
  echo "== BBG_MSM_PARTS=$parts" >> $out
  BBG_MSM_PARTS=$parts DEVBENCH_PLAIN=1 timeout 300 python scripts/devbench.py 18,19,20,21,22 "" 2>&1 | grep "^MSM" >> $out
done
echo "== waves 4, parts 2" >> $out
BBG_MSM_WAVES=4 BBG_MSM_PARTS=2 DEVBENCH_PLAIN=1 timeout 300 python scripts/devbench.py 20 "" 2>&1 | grep "^MSM" >> $out
echo "== waves 2, parts 2" >> $out
BBG_MSM_WAVES=2 BBG_MSM_PARTS=2 DEVBENCH_PLAIN=1 timeout 300 python scripts/devbench.py 20 "" 2>&1 | grep "^MSM" >> $out
echo "== waves 4, parts 4" >> $out
BBG_MSM_WAVES=4 BBG_MSM_PARTS=4 DEVBENCH_PLAIN=1 timeout 300 python scripts/devbench.py 20 "" 2>&1 | grep "^MSM" >> $out
cat $out
