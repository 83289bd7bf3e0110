#!/bin/bash
# round-2 GPU session C (1 GPU): new poly / shim pieces, prover flavours, ncu evidence
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_poly.py tests/test_gpu_round2.py tests/test_gpu_prover.py -q -m gpu > gpurun_out/r2c_pytest.txt 2>&1
tail -15 gpurun_out/r2c_pytest.txt
for bin in js_prover_gpu js_prover_gpu_l1 js_prover_cpu; do
  BBG_STATS=1 timeout 300 oracle/_ref/$bin oracle/_ref/srs_db 5 > gpurun_out/r2c_$bin.txt 2> gpurun_out/r2c_$bin.err
done
BBG_STATS=2 BBG_SHIM_TRACE=1 timeout 300 oracle/_ref/js_prover_gpu oracle/_ref/srs_db 3 > gpurun_out/r2c_trace.txt 2> gpurun_out/r2c_trace.err
python - <<'PY'
import json
for f in ("js_prover_gpu", "js_prover_gpu_l1", "js_prover_cpu"):
    try:
        d = json.loads(open("gpurun_out/r2c_%s.txt" % f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in d if k not in ("first_proof", "last_proof")})
    except Exception as e:
        print(f, "failed", e)
PY
# ---- ncu evidence (B200_PROFILING.md recipe); numbers printed under ncu are never bench values
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --strong-log-n 0 > gpurun_out/r2c_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 3 -c 1 -o gpurun_out/r2_msm_accumulate -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-ntt --no-sweep --strong-log-n 0 > gpurun_out/r2c_ncu_acc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 9 -c 3 -o gpurun_out/r2_ntt_pass -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --strong-log-n 0 > gpurun_out/r2c_ncu_ntt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_msm_segments -s 6 -c 2 -o gpurun_out/r2_msm_segments -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-ntt --no-sweep --strong-log-n 0 > gpurun_out/r2c_ncu_seg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_turbo_quotient -s 4 -c 4 -o gpurun_out/r2_turbo_quotient -f \
    oracle/_ref/js_prover_gpu oracle/_ref/srs_db 2 > gpurun_out/r2c_ncu_turbo.log 2>&1
ls -la gpurun_out/*.ncu-rep
