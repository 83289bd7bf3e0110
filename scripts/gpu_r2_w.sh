#!/bin/bash
# round-2 GPU session W (1 GPU): batched wire IFFTs (pinned staging by the copy pool): parity, byte-identical proofs, proof time
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_poly.py tests/test_gpu_prover.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2w_pytest.txt
for rep in 1 2; do
BBG_STATS=1 timeout 300 oracle/_ref/js_prover_gpu oracle/_ref/srs_db 8 > gpurun_out/r2w_prover_gpu.txt 2> gpurun_out/r2w_prover_gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2w_prover_gpu.txt").read().strip().splitlines()[-1])
print([round(p["construct_proof_s"] * 1e3, 2) for p in d["proofs"]], d["proofs"][-1]["h2d_bytes"], d["proofs"][-1]["d2h_bytes"], d["verified"])
PY
done
