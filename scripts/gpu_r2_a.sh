#!/bin/bash
# round-2 GPU session A: new round-2 tests first (fast fail), then the whole GPU suite, then the bench (N = 1)
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -m gpu > gpurun_out/r2a_pytest_round2.txt 2>&1
tail -5 gpurun_out/r2a_pytest_round2.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2a_pytest_all.txt 2>&1
tail -5 gpurun_out/r2a_pytest_all.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
tail -c 600 gpurun_out/r2a_bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2a_bench_n1.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", d["phases_ms"])
    print("batched", d.get("batched"), "pageable", d.get("e2e_pageable"))
    print("strong", {k: v for k, v in d.get("strong_2p26", {}).items() if k in ("ms_per_step", "value", "e2e", "build_s")})
    print("ntt", d["ntt"]["per_kind"], d["ntt"]["e2e"], d["ntt"]["e2e_pageable"])
    for k, v in d.get("sweep", {}).get("msm", {}).items():
        print("sweep msm", k, v.get("ms_per_step"), v.get("e2e", {}).get("ms_per_step"), v.get("error"))
    for k, v in d.get("sweep", {}).get("ntt", {}).items():
        print("sweep ntt", k, v.get("ms"), v.get("e2e", {}).get("ms_per_step"), v.get("error"))
    print("noprecomp", d.get("msm_no_precompute"), "config1", d.get("config1_geometric_2p16"))
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
