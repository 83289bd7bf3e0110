#!/bin/bash
# round-2 GPU session Y (1 GPU): the default bench line of the final build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err
tail -c 300 gpurun_out/r2y_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2y_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "pageable", d["e2e_pageable"]["ms_per_step"], "batched", d["batched"]["ms_per_msm"])
print("prover", json.dumps(d.get("prover"))[:1400])
PY
