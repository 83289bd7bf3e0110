#!/bin/bash
# round-2 GPU session K (1 GPU): persistent NTT pass with cp.async prefetch: parity (all NTT tests with the variant forced on), then A/B timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
BBG_NTT_PERSIST=1 timeout 1200 python -m pytest tests -q -m gpu -x -k "ntt or fft or NTT or coset or poly or quotient" 2>&1 | tail -5 | tee gpurun_out/r2k_pytest.txt
out=gpurun_out/r2k_ntt.txt
: > $out
for v in 0 1 0 1; do
  echo "== BBG_NTT_PERSIST=$v" >> $out
  BBG_NTT_PERSIST=$v timeout 300 python scripts/devbench.py "" 20,22,24,26 2>&1 | grep "^NTT" >> $out
done
cat $out
