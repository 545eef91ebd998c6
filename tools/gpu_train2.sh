#!/bin/bash
# full GPU suite + train bench + ncu launch list / full captures of the training kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0 OMP_NUM_THREADS=32
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -120 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench_train.py --steps 10 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; tail -3 gpurun_out/bench_train.err; cat gpurun_out/bench_train.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; cat gpurun_out/bench_1gpu.json | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/train_launches.csv python bench_train.py --steps 1 --warmup 3 --no-composition > gpurun_out/ncu_train.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_kernel -s 6 -c 2 -o gpurun_out/prof_train python bench_train.py --steps 1 --warmup 3 --no-composition > gpurun_out/ncu_train_full.log 2>&1
ls -la gpurun_out
