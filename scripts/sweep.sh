#!/bin/bash
# dev helper (GPU box): window sweep of the MSM at several sizes; one bench JSON per (log_n, c) under gpurun_out/
mkdir -p gpurun_out
for spec in "$@"; do
  lg=${spec%%:*}; cs=${spec#*:}
  for c in ${cs//,/ }; do
    BBG_MSM_C=$c timeout 200 python bench.py --no-cpu --no-ntt --steps 5 --log-n $lg > gpurun_out/sweep_n${lg}_c${c}.json 2>/dev/null
  done
done
