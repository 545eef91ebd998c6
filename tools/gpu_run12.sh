#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0 OMP_NUM_THREADS=32
timeout 300 python -m pytest tests/test_gpu_render.py tests/test_gpu_fullsize.py -m gpu -x -q -s 2>&1 | grep -E "passed|failed|Error|error|assert|early stop" | tail -8
for f in "" "--no-early-stop"; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu $f > gpurun_out/bench_es$f.json 2> gpurun_out/bench_es.err; tail -2 gpurun_out/bench_es.err
python - "$f" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_es%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1] or 'early-stop', 'value %.1f %s frame %.2f ms e2e %.1f kernel_ms %.2f'%(d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms']), {k:d['roofline_tensor'][k] for k in ('achieved','frac','tile_steps_executed_per_frame','tile_steps_without_early_termination')}, d['clocks'])
PY
done
