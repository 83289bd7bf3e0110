#!/bin/bash
# round-2 GPU session L (1 GPU): wire coefficients and t(X) stay on the device (deferred write-back): byte-identical proofs, per-proof time and PCIe bytes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_prover.py tests/test_gpu_round2.py tests/test_gpu_poly.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2l_pytest.txt
BBG_STATS=1 timeout 300 oracle/_ref/js_prover_gpu oracle/_ref/srs_db 8 > gpurun_out/r2l_prover_gpu.txt 2> gpurun_out/r2l_prover_gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2l_prover_gpu.txt").read().strip().splitlines()[-1])
print({k: d[k] for k in d if k not in ("first_proof", "last_proof")})
PY
tail -3 gpurun_out/r2l_prover_gpu.err
