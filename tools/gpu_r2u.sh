#!/bin/bash
# last check of the round at HEAD: full GPU suite, smoke, the default bench line and the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -400 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log; grep -E "FAILED|Error" gpurun_out/pytest_gpu.log | head
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value %.1f e2e %.1f exact %.1f ms/step %.2f launches %s kernel_ms %.3f frac %.3f' % (d['value'], d['e2e']['value'], d['value_exact_march'], d['ms_per_step'], d['gpu_launches'], d['roofline']['kernel_ms'], d['roofline']['frac']))
    print('c4', {k: d['c4'].get(k) for k in ('value','value_exact_march','roofline_frac')}); print('c5', json.dumps(d.get('c5_train_step'))[:230]); print('refcuda', {k: v for k, v in d.get('reference_cuda_b200', {}).items() if k != 'what'})
    print('cpu', d.get('cpu_baseline')['value'], 'clocks', d.get('clocks'))
    r=json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', r['value'])
except Exception as e:
    print('ERR', e)
PY
