#!/bin/bash
# first GPU run of the training path: regression of the forward tests, stage diagnostics, gradient parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0 OMP_NUM_THREADS=32
nvidia-smi --query-gpu=name,memory.used,memory.total --format=csv,noheader
timeout 400 python -m pytest tests/test_gpu_render.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_fwd.log; tail -3 gpurun_out/pytest_fwd.log
timeout 300 python tools/train_debug.py --stress > gpurun_out/train_debug_stress.log 2>&1; echo "debug rc=$?"; tail -70 gpurun_out/train_debug_stress.log
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q -s 2>&1 | tail -80 > gpurun_out/pytest_train.log; tail -60 gpurun_out/pytest_train.log
