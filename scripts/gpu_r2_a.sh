#!/bin/bash
# round-2 GPU session A: smoke, new round-2 tests, the whole GPU suite (no -x: collect every failure), the bench (N = 1),
# the join-split prover with the call trace
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.txt 2>&1
tail -3 gpurun_out/r2a_smoke.txt
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_poly.py -q -m gpu > gpurun_out/r2a_pytest_new.txt 2>&1
tail -30 gpurun_out/r2a_pytest_new.txt
timeout 1200 python -m pytest tests -q -m gpu --deselect tests/test_gpu_round2.py --deselect tests/test_gpu_poly.py > gpurun_out/r2a_pytest_all.txt 2>&1
tail -30 gpurun_out/r2a_pytest_all.txt
BBG_STATS=2 BBG_SHIM_TRACE=1 timeout 300 oracle/_ref/js_prover_gpu oracle/_ref/srs_db 3 > gpurun_out/r2a_prover_gpu.txt 2> gpurun_out/r2a_prover_gpu.err
tail -c 400 gpurun_out/r2a_prover_gpu.txt | cut -c1-400
grep -c bbg_call gpurun_out/r2a_prover_gpu.err
timeout 300 oracle/_ref/js_prover_gpu_l1 oracle/_ref/srs_db 3 > gpurun_out/r2a_prover_gpu_l1.txt 2>&1
timeout 300 oracle/_ref/js_prover_cpu oracle/_ref/srs_db 3 > gpurun_out/r2a_prover_cpu.txt 2>&1
python - <<'PY'
import json
for f in ("r2a_prover_gpu", "r2a_prover_gpu_l1", "r2a_prover_cpu"):
    try:
        d = json.loads(open("gpurun_out/%s.txt" % f).read().strip().splitlines()[-1])
        print(f, "keygen", d["keygen_s"], "proofs", [p["construct_proof_s"] for p in d["proofs"]], "verified", d["verified"])
    except Exception as e:
        print(f, "failed", e)
PY
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
tail -c 800 gpurun_out/r2a_bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2a_bench_n1.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "phases", d["phases_ms"])
    print("batched", d.get("batched"), "pageable", d.get("e2e_pageable"))
    print("strong", {k: v for k, v in d.get("strong_2p26", {}).items() if k in ("ms_per_step", "value", "e2e", "build_s", "phases_ms")})
    print("ntt", d["ntt"]["per_kind"], d["ntt"]["e2e"], d["ntt"]["e2e_pageable"])
    for k, v in d.get("sweep", {}).get("msm", {}).items():
        print("sweep msm", k, v.get("ms_per_step"), v.get("e2e", {}).get("ms_per_step"), v.get("phases_ms"), v.get("error"))
    for k, v in d.get("sweep", {}).get("ntt", {}).items():
        print("sweep ntt", k, v.get("ms"), v.get("e2e", {}).get("ms_per_step"), v.get("error"))
    print("noprecomp", d.get("msm_no_precompute"), "config1", d.get("config1_geometric_2p16"))
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e)
PY
