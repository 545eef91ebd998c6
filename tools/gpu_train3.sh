#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0 OMP_NUM_THREADS=32
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_dropin.py -m gpu -x -q -s 2>&1 | grep -E "passed|failed|Error|rel-L2 [0-9.e-]+ " | awk '{ if ($0 ~ /passed|failed|Error/) print; else if ($4+0 > 4e-3) print }' | tail -30
for a in 0 6 9 12 16; do
  echo "agg_levels=$a"; SDB_TABLE_AGG_LEVELS=$a timeout 200 python bench_train.py --steps 10 --warmup 3 --no-composition 2>&1 | tail -1 | cut -c1-330
done
SDB_TABLE_AGG_LEVELS=12 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/train_launches_agg12.csv python bench_train.py --steps 2 --warmup 3 --no-composition > gpurun_out/ncu_train.log 2>&1
