#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -s 2>&1 | tail -30 > gpurun_out/ops_pytest.log
timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -q -s -x 2>&1 | tail -60 > gpurun_out/render_pytest.log
cat gpurun_out/ops_pytest.log gpurun_out/render_pytest.log
