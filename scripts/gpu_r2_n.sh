#!/bin/bash
# round-2 GPU session N (1 GPU): fused small-MSM batches: parity, batch timings fused vs un-fused, prover
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_round2.py tests/test_gpu_prover.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2n_pytest.txt
cat > /tmp/batch_time.py <<'PY'
import sys, os
sys.path[:0] = ['.', 'tests', 'aztec-2.0_b200/python']
import numpy as np, torch, bbg, inputs
from oracle import pyoracle as po
bbg.init(0)
n = 1 << 16
pip = bbg.Pippenger.from_path(po.REF_SRS_DIR, n)
dev = [torch.from_numpy(inputs.fr_elements(70 + i, n).view(np.int64)).cuda() for i in range(4)]
for k in (1, 2, 4):
    tot = 0.0
    for it in range(13):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = pip.pippenger_unsafe_batch(dev[:k], 0, n); e1.record(); torch.cuda.synchronize()
        if it >= 3: tot += e0.elapsed_time(e1)
    print('batch of %d x 2^16 (fused max log2 = %s): %.3f ms per call, %.3f ms per MSM' % (k, os.environ.get('BBG_MSM_FUSED_BATCH_MAX_LOG2', 'default'), tot / 10, tot / 10 / k))
PY
python /tmp/batch_time.py 2>&1 | grep batch | tee gpurun_out/r2n_batch.txt
BBG_MSM_FUSED_BATCH_MAX_LOG2=0 python /tmp/batch_time.py 2>&1 | grep batch | tee -a gpurun_out/r2n_batch.txt
for v in default 0; do
  if [ $v = 0 ]; then export BBG_MSM_FUSED_BATCH_MAX_LOG2=0; fi
  BBG_STATS=1 timeout 300 oracle/_ref/js_prover_gpu oracle/_ref/srs_db 8 > gpurun_out/r2n_prover_$v.txt 2> gpurun_out/r2n_prover_$v.err
  python - $v <<'PY'
import json, sys
d = json.loads(open("gpurun_out/r2n_prover_%s.txt" % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], [round(p["construct_proof_s"] * 1e3, 2) for p in d["proofs"]], d["verified"])
PY
done
