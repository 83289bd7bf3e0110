#!/bin/bash
# round-2 GPU session R (1 GPU): ncu launch list of the headline step (2^20 MSM + 2^22 NTT, no strong record, no CPU leg)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --strong-log-n 0 > gpurun_out/r2r_bench_under_ncu.log 2>&1
wc -l gpurun_out/r2_launches.csv
