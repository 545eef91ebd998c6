#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -45 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fp16x3.json 2> gpurun_out/bench_fp16x3.err
timeout 300 python bench.py --steps 10 --warmup 3 --precision bf16x3 --no-cpu > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err
timeout 300 python bench.py --steps 10 --warmup 3 --precision fp16 --no-cpu > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 3 -c 1 -o gpurun_out/prof_render python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_*.json; tail -3 gpurun_out/bench_fp16x3.err
