#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_kernel -s 9 -c 1 -o gpurun_out/prof_render_inf python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_inf.log 2>&1
tail -3 gpurun_out/ncu_inf.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
