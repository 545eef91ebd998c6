#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0 OMP_NUM_THREADS=32
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_ops.py tests/test_gpu_dropin.py tests/test_gpu_render.py -m gpu -x -q -s 2>&1 | grep -E "passed|failed|Error|error|assert|sky .*rel-L2|sky_net" | tail -40
timeout 300 python bench_train.py --steps 10 --warmup 3 --no-composition 2>&1 | tail -1 | cut -c1-330
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -2 gpurun_out/bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
print('value %.1f %s frame %.2f ms e2e %.1f kernel_ms %.2f'%(d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms']))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/train_launches2.csv python bench_train.py --steps 2 --warmup 3 --no-composition > gpurun_out/ncu_train.log 2>&1
