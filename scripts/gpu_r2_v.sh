#!/bin/bash
# round-2 GPU session V (1 GPU): records of the FINAL build: full GPU suite, default bench, reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2v_pytest_gpu.txt 2>&1
tail -4 gpurun_out/r2v_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err
tail -c 300 gpurun_out/r2v_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r2v_bench_ref.json 2> gpurun_out/r2v_bench_ref.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2v_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "pageable", d["e2e_pageable"]["ms_per_step"], "batched", d["batched"]["ms_per_msm"])
print("phases", d["phases_ms"])
print("prover", json.dumps(d.get("prover"))[:1200])
r = json.loads(open("gpurun_out/r2v_bench_ref.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["ms_per_step"], "ratio e2e", d["e2e"]["value"] / r["value"])
PY
