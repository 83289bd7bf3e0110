#!/bin/bash
# round-2 GPU session E (1 GPU): prover after the batch fix, MSM tail timings, quick regression of the MSM tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_fullsize.py -q -m gpu -x > gpurun_out/r2e_pytest.txt 2>&1
tail -3 gpurun_out/r2e_pytest.txt
for bin in js_prover_gpu; do
  BBG_STATS=1 timeout 300 oracle/_ref/$bin oracle/_ref/srs_db 6 > gpurun_out/r2e_$bin.txt 2> gpurun_out/r2e_$bin.err
done
BBG_STATS=2 BBG_SHIM_TRACE=1 timeout 300 oracle/_ref/js_prover_gpu oracle/_ref/srs_db 3 > gpurun_out/r2e_trace.txt 2> gpurun_out/r2e_trace.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2e_js_prover_gpu.txt").read().strip().splitlines()[-1])
print({k: d[k] for k in d if k not in ("first_proof", "last_proof")})
PY
timeout 120 python scripts/devbench.py 16 "" 2>&1 | grep "^MSM"
timeout 120 python scripts/devbench.py 18 "" 2>&1 | grep "^MSM"
timeout 120 python scripts/devbench.py 20 "16,22" 2>&1 | grep "^MSM\|^NTT"
