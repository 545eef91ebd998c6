#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"table3_backward|mlp_kernel|composite_backward" -c 200 --csv --log-file gpurun_out/train_scatter.csv python bench_train.py --steps 8 --warmup 3 --no-composition > gpurun_out/ncu_train.log 2>&1
grep -E "table3_backward" gpurun_out/train_scatter.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
grep -E "mlp_kernel<2, 0, 0, 1>" gpurun_out/train_scatter.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
grep -E "mlp_kernel<1, 0, 2, 0>" gpurun_out/train_scatter.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
