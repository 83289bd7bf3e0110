#!/bin/bash
# round-2 GPU session B (2 GPUs): real-NCCL parity of the sharded paths, then the bench at N = 2 (weak line, strong 2^26
# record with parity, sharded NTT with parity) and the reference arm
set -x
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2b_smi.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/r2b_dist_check.txt 2>&1
tail -4 gpurun_out/r2b_dist_check.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "multi_process or nccl" > gpurun_out/r2b_pytest_nccl.txt 2>&1
tail -3 gpurun_out/r2b_pytest_nccl.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
tail -c 600 gpurun_out/r2b_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2b_bench_n2.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "parity", d.get("parity"))
    s = d.get("strong_2p26", {})
    print("strong", {k: s.get(k) for k in ("ms_per_step", "value", "e2e", "parity", "build_s")})
    print("ntt", d["ntt"].get("per_kind"), d["ntt"].get("parity"))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 oracle/_ref/js_prover_gpu oracle/_ref/srs_db 3 > gpurun_out/r2b_prover_gpu.txt 2> gpurun_out/r2b_prover_gpu.err
tail -c 300 gpurun_out/r2b_prover_gpu.txt | head -c 10
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2b_prover_gpu.txt").read().strip().splitlines()[-1])
print("prover", {k: d[k] for k in d if k not in ("first_proof", "last_proof")})
PY
