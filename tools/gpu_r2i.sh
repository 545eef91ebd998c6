#!/bin/bash
# ray-slot kernel bring-up: guarded small runs first, then the suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
timeout 90 python __graft_entry__.py smoke 2>&1 | tail -2 || echo "SMOKE TIMED OUT / FAILED"
timeout 150 python tools/ray_stats.py --frames 2 2>&1 | tail -5 || echo "RAY_STATS TIMED OUT"
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_rq.json 2> gpurun_out/bench_rq.err || echo "BENCH TIMED OUT"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_rq.json').read().strip().splitlines()[-1])
    print('value %.1f e2e %.1f exact %.1f ms/step %.2f kernel %.2f shaded %.2fM variant %s' % (d['value'], d['e2e']['value'], d['value_exact_march'], d['ms_per_step'], d['roofline']['kernel_ms'], d['samples_shaded_per_frame']/1e6, d['roofline'].get('kernel_variant','')[:20]))
except Exception as e: print('ERR', e)
PY
SDB_RAY_SLOTS=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TILE kernel: value %.1f exact %.1f kernel %.2f' % (d['value'], d['value_exact_march'], d['roofline']['kernel_ms']))"
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
