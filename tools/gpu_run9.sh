#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0 OMP_NUM_THREADS=32
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|Error|error|C2 window|C4 window|assert" | tail -15
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -2 gpurun_out/bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_1gpu.json').read().strip().splitlines()[-1])
print('value %.1f %s frame %.2f ms e2e %.1f kernel_ms %.2f'%(d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms']), d['roofline_tensor'], d['clocks'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1
grep -E "dda_perspective|sky_mean|mlp_kernel|prepass" gpurun_out/launches.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,200-260 | tail -12
