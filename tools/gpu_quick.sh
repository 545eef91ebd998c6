#!/bin/bash
# quick regression + numbers: render/train/dropin tests, headline bench, train bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0 OMP_NUM_THREADS=32
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_train.py tests/test_gpu_dropin.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -2 gpurun_out/bench_q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print('value %.1f exact %.1f frame %.2f ms e2e %.1f kernel_ms %.2f'%(d['value'], d['value_exact_march'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms']), d['clocks'])
PY
SDB_TIMING=1 timeout 200 python bench_train.py --steps 16 --warmup 8 --no-composition 2> gpurun_out/train_timing.err | cut -c1-330; grep "sdb timing" gpurun_out/train_timing.err | tail -3
