#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0 OMP_NUM_THREADS=32
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -12
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_dda.json 2> gpurun_out/bench_dda.err; tail -2 gpurun_out/bench_dda.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_dda.json').read().strip().splitlines()[-1])
print('value %.1f exact %.1f frame %.2f ms e2e %.1f kernel_ms %.2f'%(d['value'], d['value_exact_march'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms']))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dda_perspective|height_bound" -c 12 --csv --log-file gpurun_out/dda.csv python bench.py --steps 3 --warmup 3 --no-cpu > /dev/null 2>&1
grep -E "dda_perspective|height_bound" gpurun_out/dda.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -6
timeout 200 python bench_train.py --steps 32 --warmup 16 --no-composition | cut -c1-300
