#!/bin/bash
# round-2 GPU session AA (1 GPU): dedicated Montgomery squaring (100 instead of 128 wide multiply-adds): whole GPU suite, then A/B timing against the old build's record
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/r2aa_pytest_gpu.txt 2>&1
tail -4 gpurun_out/r2aa_pytest_gpu.txt
timeout 300 python scripts/devbench.py 16,20 22 2>&1 | grep "^MSM\|^NTT kind 0" | tee gpurun_out/r2aa_devbench.txt
DEVBENCH_PLAIN=1 timeout 300 python scripts/devbench.py 20 "" 2>&1 | grep "^MSM" | tee -a gpurun_out/r2aa_devbench.txt
