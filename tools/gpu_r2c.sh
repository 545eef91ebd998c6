#!/bin/bash
# round 2, call C: generator test, default bench with all legs, ncu launch list + full captures (text summaries kept)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
timeout 900 python -m pytest tests/test_gpu_generator.py -m gpu -q -s 2>&1 | tail -60 > gpurun_out/pytest_generator.log
grep -E "FULL C2|timeline ms|hook stats|depth error|passed|failed" gpurun_out/pytest_generator.log | cut -c1-400
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print('value %.1f e2e %.1f exact %.1f ms/step %.2f launches %s cnn_ms %s' % (d['value'], d['e2e']['value'], d['value_exact_march'], d['ms_per_step'], d['gpu_launches'], d.get('rendercnn_ms')))
    print('roofline', {k: d['roofline'][k] for k in ('bound','achieved','peak','frac','executed_frac','kernel_ms')})
    print('c4', d.get('c4')); print('c5', json.dumps(d.get('c5_train_step'))[:300]); print('refcuda', {k: v for k, v in d.get('reference_cuda_b200', {}).items() if k != 'what'})
    print('cpu', d.get('cpu_baseline')); print('clocks', d.get('clocks'))
except Exception as e:
    print('ERR', e)
PY
NCU="ncu --clock-control none"
timeout 400 $NCU --metrics gpu__time_duration.sum -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_launches.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:mlp_kernel -s 9 -c 1 -o gpurun_out/prof_render python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_render.log 2>&1
timeout 400 $NCU --set full -k regex:conv_kernel -s 8 -c 1 -o gpurun_out/prof_conv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_conv.log 2>&1
timeout 400 $NCU --set full -k regex:dda_perspective -s 3 -c 1 -o gpurun_out/prof_dda python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_dda.log 2>&1
timeout 400 $NCU --set full -k regex:wgrad_kernel -s 2 -c 1 -o gpurun_out/prof_wgrad python bench_train.py --steps 2 --warmup 3 --no-composition > gpurun_out/ncu_wgrad.log 2>&1
timeout 400 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/train_launches.csv python bench_train.py --steps 2 --warmup 2 --no-composition > gpurun_out/ncu_train_launches.log 2>&1
for k in render conv dda wgrad; do
  if [ -f gpurun_out/prof_$k.ncu-rep ]; then
    ncu -i gpurun_out/prof_$k.ncu-rep --page details > gpurun_out/prof_${k}_details.txt 2>/dev/null
    ncu -i gpurun_out/prof_$k.ncu-rep --page raw --csv > gpurun_out/prof_${k}_raw.csv 2>/dev/null
    [ $k != render ] && rm -f gpurun_out/prof_$k.ncu-rep
  fi
done
ls -la gpurun_out | head -40
