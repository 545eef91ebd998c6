#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
PREC=2 timeout 120 python gpu_debug.py 2>&1 | tail -12
timeout 240 python -m pytest tests/test_gpu_render.py -m gpu -q -x 2>&1 | tail -5
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_q_fp16x3.json 2> gpurun_out/bench_q.err
timeout 200 python bench.py --steps 10 --warmup 3 --precision fp16 --no-cpu > gpurun_out/bench_q_fp16.json 2>> gpurun_out/bench_q.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_q_fp16x3.json','gpurun_out/bench_q_fp16.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'value %.1f Msamples/s  frame %.2f ms  kernel %.2f ms  e2e %.1f'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value']))
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/bench_q.err
