#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
N=${1:-2}
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@"; }
[ "$N" = 1 ] && timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-extras --mode strong > gpurun_out/strong_n1.json 2> gpurun_out/strong_n1.err
timeout 400 bash -c "$(declare -f run); run $N --steps 10 --warmup 3 --mode strong" > gpurun_out/strong_nN.json 2> gpurun_out/strong_nN.err
python - <<'PY'
import json
for f in ('strong_n1', 'strong_nN'):
    try:
        d = json.loads(open('gpurun_out/%s.json' % f).read().strip().splitlines()[-1])
        print(f, 'n=%d value %.1f e2e %.1f ms/step %.2f' % (d['n_gpus'], d['value'], d['e2e']['value'], d['ms_per_step']), [[round(v, 2) for v in r[:5]] for r in d['per_rank_ms']['rows']])
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -3 gpurun_out/strong_nN.err
