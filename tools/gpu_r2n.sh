#!/bin/bash
# round 2 final validation (1 GPU): suite, smoke, bench with all legs, reference arm, launch list, train bench, boundary-op timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | tail -400 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log; grep -E "FAILED|Error" gpurun_out/pytest_gpu.log | head
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value %.1f e2e %.1f exact %.1f ms/step %.2f launches %s cnn_ms %s' % (d['value'], d['e2e']['value'], d['value_exact_march'], d['ms_per_step'], d['gpu_launches'], d.get('rendercnn_ms')))
    print('roofline', {k: d['roofline'][k] for k in ('bound','achieved','peak','frac','executed_frac','kernel_ms','traffic')})
    print('c4', {k: d['c4'].get(k) for k in ('value','value_exact_march','ms_per_step','e2e','roofline_frac')}); print('c5', json.dumps(d.get('c5_train_step'))[:260]); print('refcuda', {k: v for k, v in d.get('reference_cuda_b200', {}).items() if k != 'what'})
    print('cpu', d.get('cpu_baseline')); print('clocks', d.get('clocks'))
    r=json.loads(open('gpurun_out/bench_reference.json').read().strip().splitlines()[-1]); print('reference arm', r['value'], r['cpu_baseline']['cores'])
except Exception as e:
    print('ERR', e)
PY
timeout 400 ncu --clock-control none --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_launches.log 2>&1
timeout 300 python bench_train.py --steps 10 --warmup 3 > gpurun_out/train.json 2> gpurun_out/train.err; cut -c1-400 gpurun_out/train.json
timeout 600 python tests/ops_timing.py > gpurun_out/ops_timing.json 2> gpurun_out/ops_timing.err; cut -c1-1500 gpurun_out/ops_timing.json; tail -3 gpurun_out/ops_timing.err
timeout 300 ncu --set full --clock-control none -k 'regex:grid_|input_backward' -s 3 -c 3 -f -o gpurun_out/prof_gridenc env PYTHONPATH=. python tools/gridenc_run.py > gpurun_out/prof_gridenc.log 2>&1
ncu -i gpurun_out/prof_gridenc.ncu-rep --page raw --csv > gpurun_out/prof_gridenc_raw.csv 2>/dev/null; rm -f gpurun_out/prof_gridenc.ncu-rep; tail -2 gpurun_out/prof_gridenc.log
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-extras --mode strong > gpurun_out/strong_n1.json 2> gpurun_out/strong_n1.err; cut -c1-200 gpurun_out/strong_n1.json; tail -2 gpurun_out/strong_n1.err
timeout 400 ncu --clock-control none --set full -k regex:mlp_kernel -s 9 -c 1 -f -o gpurun_out/prof_render python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_render.log 2>&1
ncu -i gpurun_out/prof_render.ncu-rep --page raw --csv > gpurun_out/prof_render_raw.csv 2>/dev/null; rm -f gpurun_out/prof_render.ncu-rep; tail -1 gpurun_out/ncu_render.log
