#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -2
SDB_TIMING=1 timeout 300 python bench_train.py --profile 2>&1 | grep "sdb timing" | tail -5
timeout 200 python bench_train.py --steps 16 --warmup 8 --no-composition | cut -c1-300
