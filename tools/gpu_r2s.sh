#!/bin/bash
# gather loads in flight (SDB_GATHER_UNROLL): parity of a thin variant, then bench of each (the 4- and 1-corner variants existed for this A/B only:
# the shipped library keeps 2 (default) and 8; profiles/r02_exp_gather_unroll.json)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
SDB_GATHER_UNROLL=2 timeout 600 python -m pytest tests -m gpu -q -x -k "render or fullsize" 2>&1 | tail -3
for v in 8 4 2 1; do
  SDB_GATHER_UNROLL=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_gu$v.json 2> gpurun_out/bench_gu$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_gu$v.json').read().strip().splitlines()[-1])
print('gather_unroll=$v value %.1f e2e %.1f exact %.1f ms/step %.2f kernel_ms %.3f frac %.3f clocks %s' % (d['value'], d['e2e']['value'], d['value_exact_march'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']['sm_mhz']))
PY
done
