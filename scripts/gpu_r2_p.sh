#!/bin/bash
# round-2 GPU session P (1 GPU): re-tune of the accumulate chunking (waves) and the window width with the team-cooperative merge in place
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
out=gpurun_out/r2p_tune.txt
: > $out
for w in 2 3 4 5 6; do
  echo "== BBG_MSM_WAVES=$w" >> $out
  BBG_MSM_WAVES=$w DEVBENCH_PLAIN=1 timeout 300 python scripts/devbench.py 16,18,20 "" 2>&1 | grep "^MSM" >> $out
done
for c in 19 21; do
  echo "== BBG_MSM_C=$c" >> $out
  BBG_MSM_C=$c DEVBENCH_PLAIN=1 timeout 300 python scripts/devbench.py 20 "" 2>&1 | grep "^MSM" >> $out
done
for c in 15 17; do
  echo "== BBG_MSM_C=$c (2^16)" >> $out
  BBG_MSM_C=$c DEVBENCH_PLAIN=1 timeout 300 python scripts/devbench.py 16 "" 2>&1 | grep "^MSM" >> $out
done
cat $out
