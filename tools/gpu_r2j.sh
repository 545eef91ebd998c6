#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], 'value %.1f e2e %.1f exact %.1f ms/step %.2f kernel %.2f shaded %.2fM frac %.3f exec %.3f' % (d['value'], d['e2e']['value'], d['value_exact_march'], d['ms_per_step'], d['roofline']['kernel_ms'], d['samples_shaded_per_frame']/1e6, d['roofline']['frac'], d['roofline']['executed_frac']))" "$1" "$2"; }
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_rq.json 2> gpurun_out/bench_rq.err; show gpurun_out/bench_rq.json RAYSLOTS
SDB_RAY_SLOTS=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_tile.json 2> gpurun_out/bench_tile.err; show gpurun_out/bench_tile.json TILES
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-extras --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; show gpurun_out/bench_c4.json C4
timeout 400 ncu --clock-control none --set full --import-source on -k regex:mlp_kernel -s 9 -c 1 -o gpurun_out/prof_render python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_render.log 2>&1
ncu -i gpurun_out/prof_render.ncu-rep --page raw --csv > gpurun_out/prof_render_raw.csv 2>/dev/null; ncu -i gpurun_out/prof_render.ncu-rep --page details > gpurun_out/prof_render_details.txt 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/prof_render_raw.csv'))); d=dict(zip(rows[0],rows[2]))
for m in ('Kernel Name','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed_op_local_ld.sum'): print(m, d.get(m))
PY
