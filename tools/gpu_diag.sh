#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -1
timeout 200 python bench_train.py --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; cut -c100-330 gpurun_out/bench_train.json; tail -c 300 gpurun_out/bench_train.json
