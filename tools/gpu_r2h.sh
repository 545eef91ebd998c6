#!/bin/bash
# full GPU suite + smoke + quick strong-mode sanity on one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | tail -300 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
grep -E "FULL C2|timeline ms|hook stats|FAILED|Error" gpurun_out/pytest_gpu.log | cut -c1-400 | head -20
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extras --mode strong 2>gpurun_out/strong1.err | cut -c1-300; tail -3 gpurun_out/strong1.err
