#!/bin/bash
# round 2, call F (8 GPUs): scaling of the headline bench -- N=1 baseline on the same box, N=8 weak, N=8 strong, N=8 C3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@"; }
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
timeout 400 bash -c "$(declare -f run); run 8 --steps 20 --warmup 3" > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err
timeout 400 bash -c "$(declare -f run); run 4 --steps 20 --warmup 3" > gpurun_out/scale_n4.json 2> gpurun_out/scale_n4.err
timeout 400 bash -c "$(declare -f run); run 8 --steps 20 --warmup 3 --mode strong" > gpurun_out/scale_n8_strong.json 2> gpurun_out/scale_n8_strong.err
timeout 400 bash -c "$(declare -f run); run 8 --steps 5 --warmup 3 --workload c3" > gpurun_out/scale_n8_c3.json 2> gpurun_out/scale_n8_c3.err
python - <<'PY'
import json
def load(f):
    try: return json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
    except Exception as e: return None
b = load('scale_n1')
for f in ('scale_n1', 'scale_n4', 'scale_n8', 'scale_n8_strong', 'scale_n8_c3'):
    d = load(f)
    if not d: print(f, 'NO RESULT'); continue
    print('%-16s n=%d %-6s value %.1f (x%.2f of n1) e2e %.1f (x%.2f) exact %.1f  coll %s' % (f, d['n_gpus'], d['scaling'], d['value'], d['value'] / b['value'] if b else 0, d['e2e']['value'], d['e2e']['value'] / b['e2e']['value'] if b else 0, d['value_exact_march'], (d.get('collective') or {}).get('ms_per_step_incl_wait_for_slowest_rank')))
    print('    per-rank', [[round(v, 2) for v in r] for r in d['per_rank_ms']['rows']])
PY
tail -2 gpurun_out/scale_n8.err
