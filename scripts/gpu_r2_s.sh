#!/bin/bash
# round-2 GPU session S (1 GPU): `ncu --set full` captures of the dominant kernels of the FINAL build (B200_PROFILING.md recipe)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
F="--steps 2 --warmup 3 --no-cpu --no-sweep --strong-log-n 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 3 -c 1 -o gpurun_out/r2_msm_accumulate -f python bench.py $F --no-ntt > gpurun_out/r2s_ncu_acc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 9 -c 3 -o gpurun_out/r2_ntt_pass -f python bench.py $F > gpurun_out/r2s_ncu_ntt.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msm_segments -s 6 -c 2 -o gpurun_out/r2_msm_segments -f python bench.py $F --no-ntt > gpurun_out/r2s_ncu_seg.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msm_merge -s 9 -c 4 -o gpurun_out/r2_msm_merge -f python bench.py $F --no-ntt > gpurun_out/r2s_ncu_merge.log 2>&1
ls -la gpurun_out/*.ncu-rep
