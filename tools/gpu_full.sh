#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0 OMP_NUM_THREADS=32
timeout 600 python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_fp16x3.json 2> gpurun_out/bench_fp16x3.err
timeout 200 python bench.py --steps 20 --warmup 3 --precision fp16 --no-cpu > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
timeout 200 python bench.py --steps 20 --warmup 3 --precision bf16x3 --no-cpu > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mlp_kernel -s 9 -c 1 -o gpurun_out/prof_render3 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full3.log 2>&1
python - <<'PY'
import json
for f in ('bench_fp16x3','bench_fp16','bench_bf16x3','bench_reference'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, 'value %.3f %s frame %.2f ms'%(d['value'], d['unit'], d['ms_per_step']), d.get('roofline',{}).get('kernel_ms'), d.get('cpu_baseline'), d.get('clocks'))
    except Exception as e: print(f, 'ERR', e)
PY
