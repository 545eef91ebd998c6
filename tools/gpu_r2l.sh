#!/bin/bash
# 8 GPUs: scaling with the ray-slot kernel (same box, back to back)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@"; }
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu --no-extras > gpurun_out/scale2_n1.json 2> gpurun_out/scale2_n1.err
for n in 2 4 8; do timeout 400 bash -c "$(declare -f run); run $n --steps 20 --warmup 3" > gpurun_out/scale2_n$n.json 2> gpurun_out/scale2_n$n.err; done
timeout 400 bash -c "$(declare -f run); run 8 --steps 20 --warmup 3 --mode strong" > gpurun_out/scale2_n8_strong.json 2> gpurun_out/scale2_n8_strong.err
python - <<'PY'
import json
def load(f):
    try: return json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
    except Exception as e: return None
b = load('scale2_n1')
for f in ('scale2_n1', 'scale2_n2', 'scale2_n4', 'scale2_n8', 'scale2_n8_strong'):
    d = load(f)
    if not d: print(f, 'NO RESULT'); continue
    print('%-18s n=%d %-6s value %.1f (x%.2f) e2e %.1f (x%.2f) exact %.1f  coll %s' % (f, d['n_gpus'], d['scaling'], d['value'], d['value'] / b['value'] if b else 0, d['e2e']['value'], d['e2e']['value'] / b['e2e']['value'] if b else 0, d['value_exact_march'], (d.get('collective') or {}).get('ms_per_step_incl_wait_for_slowest_rank')))
    print('    per-rank', [[round(v, 2) for v in r[:5]] for r in d['per_rank_ms']['rows']])
PY
tail -2 gpurun_out/scale2_n8_strong.err
