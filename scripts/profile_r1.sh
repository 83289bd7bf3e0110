#!/bin/bash
# GPU box: ncu evidence for the current build (B200_PROFILING.md recipe).  Outputs under gpurun_out/.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 3 -c 1 -o gpurun_out/msm_accumulate -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-ntt > gpurun_out/ncu_acc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 9 -c 3 -o gpurun_out/ntt_pass -f \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_ntt.log 2>&1
ls -la gpurun_out/*.ncu-rep
