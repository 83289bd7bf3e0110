#!/bin/bash
# round-2 GPU session H (1 GPU): NTT A/B timings (TMA twiddles, 2-CTA variant, E = 8), prover with per-proof PCIe bytes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
out=gpurun_out/r2h_ntt.txt
: > $out
run() { echo "== $*" >> $out; env "$@" timeout 120 python scripts/devbench.py "" 18,20,22,24 2>&1 | grep "^NTT kind 0\|^NTT kind 2" >> $out; }
run BBG_X=0
run BBG_NTT_TMA_TWIDDLES=0
run BBG_NTT_E4_CTAS=2
run BBG_NTT_LOGE=3
run BBG_X=1
run BBG_NTT_TMA_TWIDDLES=0
cat $out
BBG_STATS=1 timeout 300 oracle/_ref/js_prover_gpu oracle/_ref/srs_db 6 > gpurun_out/r2h_prover_gpu.txt 2> gpurun_out/r2h_prover_gpu.err
BBG_STATS=1 timeout 300 oracle/_ref/js_prover_gpu_l1 oracle/_ref/srs_db 4 > gpurun_out/r2h_prover_gpu_l1.txt 2> gpurun_out/r2h_prover_gpu_l1.err
python - <<'PY'
import json
for f in ("r2h_prover_gpu", "r2h_prover_gpu_l1"):
    d = json.loads(open("gpurun_out/%s.txt" % f).read().strip().splitlines()[-1])
    print(f, {k: d[k] for k in d if k not in ("first_proof", "last_proof")})
PY
