#!/bin/bash
# round 2, call B: full GPU suite (not -x), train-step stage timing, ray stats, quick bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | tail -250 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
grep -E "FULL C2|timeline ms|hook stats|C2 window|C4 window|RenderCNN|FAILED|Error|error" gpurun_out/pytest_gpu.log | cut -c1-400 | head -40
SDB_TIMING=1 timeout 300 python bench_train.py --steps 4 --warmup 3 --no-composition > gpurun_out/train.json 2> gpurun_out/train.err
grep "sdb timing" gpurun_out/train.err | tail -3; cat gpurun_out/train.json | cut -c1-400
