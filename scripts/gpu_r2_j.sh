#!/bin/bash
# round-2 GPU session J (1 GPU): MSM cut into bucket-range parts (pipelined tails): timings for 1 / 2 / 4 parts
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
out=gpurun_out/r2j_msm.txt
: > $out
for parts in 1 2 4; do
  echo "== BBG_MSM_PARTS=$parts" >> $out
  BBG_MSM_PARTS=$parts DEVBENCH_PLAIN=1 timeout 300 python scripts/devbench.py 18,19,20 "" 2>&1 | grep "^MSM\|Error\|error" >> $out
done
echo "== parts 2 with phases" >> $out
BBG_MSM_PARTS=2 timeout 300 python scripts/devbench.py 20 "" 2>&1 | grep "^MSM" >> $out
echo "== parts 1 with phases" >> $out
BBG_MSM_PARTS=1 timeout 300 python scripts/devbench.py 20 "" 2>&1 | grep "^MSM" >> $out
cat $out
