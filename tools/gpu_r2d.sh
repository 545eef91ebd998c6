#!/bin/bash
# round 2, call D: full GPU suite, train step (stage timing + host/device profile), zero-edit timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | tail -250 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
grep -E "FULL C2|timeline ms|hook stats|RenderCNN 570|FAILED|Error" gpurun_out/pytest_gpu.log | cut -c1-400 | head -20
timeout 300 python bench_train.py --steps 10 --warmup 3 > gpurun_out/train.json 2> gpurun_out/train.err; cut -c1-700 gpurun_out/train.json
SDB_TIMING=1 timeout 300 python bench_train.py --steps 3 --warmup 3 --no-composition 2>&1 | grep "sdb timing" | tail -2
timeout 300 python bench_train.py --profile 2>&1 | tail -8
for b in ref dropin; do timeout 600 python -m oracle.refgen --backend $b --frames 4 --warm 1 --out /tmp/x_$b.npz --workdir /tmp/x_refgen 2>&1 | tail -1; done
