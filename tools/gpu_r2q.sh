#!/bin/bash
# paired gather A/B: parity (ray slots == tile kernel bit for bit), bench both ways, timeline
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
timeout 600 python -m pytest tests -m gpu -q -x -k "render or fullsize" 2>&1 | tail -4
for v in 1 0; do
  SDB_PAIR_GATHER=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_pg$v.json 2> gpurun_out/bench_pg$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pg$v.json').read().strip().splitlines()[-1])
print('pair_gather=$v value %.1f e2e %.1f exact %.1f ms/step %.2f kernel_ms %.3f frac %.3f clocks %s' % (d['value'], d['e2e']['value'], d['value_exact_march'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks']['sm_mhz']))
PY
done
[ -n "$SKIP_TIMELINE" ] && exit 0
# diagnostics build for the timeline (stamps compiled in), then back to the product build
SDB_NVCC_EXTRA=-DSDB_TIMELINE python -m scenedreamer_b200.build > gpurun_out/build_tl.log 2>&1
for v in 1 0; do
SDB_PAIR_GATHER=$v PYTHONPATH=. timeout 300 python tools/render_timeline.py 60 > gpurun_out/render_timeline_pg$v.txt 2> gpurun_out/render_timeline.err; tail -10 gpurun_out/render_timeline_pg$v.txt; tail -3 gpurun_out/render_timeline.err
done
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1
