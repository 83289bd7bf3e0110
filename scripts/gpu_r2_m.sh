#!/bin/bash
# round-2 GPU session M (8 GPUs): real-NCCL parity of the sharded paths (incl. the fused NTT exchange over 8 peers) and the N = 8 bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 scripts/dist_check.py 2>&1 | grep "rank\|Error\|error" | tee gpurun_out/r2m_dist_check_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --no-sweep 2>gpurun_out/r2m_bench_n$N.err | tail -1 > gpurun_out/r2m_bench_n$N.json
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open("gpurun_out/r2m_bench_n%s.json" % n).read())
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "e2e")})
print("strong", json.dumps(d.get("strong_2p26"))[:900])
print("parity", json.dumps(d.get("parity"))[:600])
print("ntt", json.dumps(d.get("ntt"))[:1600])
PY
tail -3 gpurun_out/r2m_bench_n$N.err
