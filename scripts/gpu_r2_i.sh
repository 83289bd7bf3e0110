#!/bin/bash
# round-2 GPU session I (2 GPUs): fused-exchange NTT -- simulated on one GPU, then real peer stores under torchrun; N = 2 bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused_exchange or data_path_simulated or natural_blocks" 2>&1 | tail -5 | tee gpurun_out/r2i_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/dist_check.py 2>&1 | tail -8 | tee gpurun_out/r2i_dist_check.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --no-sweep 2>gpurun_out/r2i_bench_n2.err | tail -1 > gpurun_out/r2i_bench_n2.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2i_bench_n2.json").read())
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d.get("strong_2p26", {}).get("ms"), json.dumps(d["ntt"])[:1500])
PY
tail -5 gpurun_out/r2i_bench_n2.err
