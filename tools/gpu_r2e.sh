#!/bin/bash
# round 2, call E (1 GPU): the new tests only (f2 Adam, f3 world builder, f4 sampler) + a quick bench_train with the hook-less sync-free backward
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
timeout 1200 python -m pytest tests/test_gpu_generator.py tests/test_gpu_train.py tests/test_gpu_ops.py -m gpu -q -s 2>&1 | tail -120 > gpurun_out/pytest_e.log
tail -8 gpurun_out/pytest_e.log; grep -E "FAILED|Error|Mismatch|Max " gpurun_out/pytest_e.log | head -20
