#!/bin/bash
# round-2 GPU session O (1 GPU): the records of the round -- compute-sanitizer, full GPU suite, default bench + reference arm, ncu launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python scripts/sanitize_run.py small > gpurun_out/r2o_san_plain.txt 2>&1; tail -2 gpurun_out/r2o_san_plain.txt
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_run.py small > gpurun_out/r2o_san_memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2o_san_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_run.py small > gpurun_out/r2o_san_racecheck.txt 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/r2o_san_racecheck.txt
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2o_pytest_gpu.txt 2>&1
tail -4 gpurun_out/r2o_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err
tail -c 400 gpurun_out/r2o_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r2o_bench_ref.json 2> gpurun_out/r2o_bench_ref.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2o_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "pageable", d["e2e_pageable"]["ms_per_step"], "batched", d["batched"]["ms_per_msm"])
print("phases", d["phases_ms"])
print("prover", json.dumps(d.get("prover"))[:2500])
r = json.loads(open("gpurun_out/r2o_bench_ref.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["ms_per_step"], "ratio e2e", d["e2e"]["value"] / r["value"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 2 --warmup 1 --no-sweep > gpurun_out/r2o_ncu_bench.log 2>&1
tail -2 gpurun_out/r2o_launches.csv | cut -c1-300
