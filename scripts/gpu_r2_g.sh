#!/bin/bash
# round-2 GPU session G (1 GPU): NTT experiments -- TMA-staged stage twiddles (guarded: a hang falls back), 2-CTA variant
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
out=gpurun_out/r2g_ntt.txt
: > $out
cat > /tmp/ntt_quick.py <<'PY'
import sys
sys.path[:0] = ['.', 'tests', 'aztec-2.0_b200/python']
import numpy as np, bbg, inputs
from oracle import pyoracle as po
bbg.init(0)
orc = po.Oracle()
for lg in (17, 18):
    x = inputs.fr_elements(lg, 1 << lg)
    y = bbg.coset_fft(x.copy())
    z = bbg.coset_ifft(y.copy())
    assert np.array_equal(orc.reduce(po.FR, z), orc.reduce(po.FR, x)), lg
x = inputs.fr_elements(3, 1 << 17)
assert np.array_equal(orc.reduce(po.FR, bbg.fft(x.copy())), orc.reduce(po.FR, orc.ntt(po.NTT_FFT, x)))
print("ntt quick ok")
PY
if timeout 90 python /tmp/ntt_quick.py >> $out 2>&1; then echo "TMA twiddles: ok" >> $out; else echo "TMA twiddles: FAILED (rc=$?), falling back" >> $out; export BBG_NTT_TMA_TWIDDLES=0; fi
run() { echo "== $*" >> $out; env "$@" timeout 120 python scripts/devbench.py "" 18,20,22,24 2>&1 | grep "^NTT kind 0\|^NTT kind 2" >> $out; }
run BBG_X=0
run BBG_NTT_TMA_TWIDDLES=0
run BBG_NTT_E4_CTAS=2
run BBG_NTT_LOGE=3
cat $out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_reference_cases.py -q -m gpu -k "ntt or fft or NTT" > gpurun_out/r2g_pytest.txt 2>&1
tail -3 gpurun_out/r2g_pytest.txt
