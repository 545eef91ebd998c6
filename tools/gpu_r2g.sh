#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python tools/world_debug.py 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_generator.py -m gpu -q -k "world or sampler" 2>&1 | tail -5
