#!/bin/bash
# round 2, call A: full GPU test suite (incl. the real reference Generator), smoke, default bench with all legs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }   # no-op when the shipped .so matches the sources
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
timeout 1500 python -m pytest tests -m gpu -q -s -x 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
grep -E "FULL C2|timeline ms|hook stats|C2 window|C4 window" gpurun_out/pytest_gpu.log
timeout 300 python tools/ray_stats.py --frames 4 > gpurun_out/ray_stats.log 2>&1; cat gpurun_out/ray_stats.log | cut -c1-400
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value %.1f e2e %.1f exact %.1f ms/step %.2f launches %s' % (d['value'], d['e2e']['value'], d['value_exact_march'], d['ms_per_step'], d['gpu_launches']))
    print('roofline', {k: d['roofline'][k] for k in ('bound','achieved','peak','frac','executed_frac','kernel_ms')})
    print('shaded', d['samples_shaded_per_frame'], 'credited', d['samples_credited_per_frame'])
    print('c4', d.get('c4')); print('c5', d.get('c5_train_step')); print('refcuda', d.get('reference_cuda_b200'))
    print('cpu', d.get('cpu_baseline')); print('clocks', d.get('clocks')); print('per_rank', d.get('per_rank_ms'))
except Exception as e:
    print('ERR', e)
PY
