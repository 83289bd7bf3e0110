#!/bin/bash
# round-2 GPU session X (1 GPU): pageable end-to-end MSM vs staging-pool threads
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
out=gpurun_out/r2x_pageable.txt
: > $out
nproc >> $out
for t in default 0 2 4 8 12 16; do
  if [ $t = default ]; then timeout 200 python scripts/dev/pageable_time.py 2>&1 | grep "ms" >> $out; else BBG_STAGING_THREADS=$t timeout 200 python scripts/dev/pageable_time.py 2>&1 | grep "ms" >> $out; fi
done
cat $out
