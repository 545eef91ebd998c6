#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
python - <<'PY'
import json
for f in ('bench_1gpu','bench_2gpu'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, 'n_gpus', d['n_gpus'], 'value %.1f exact %.1f %s frame %.2f ms e2e %.1f'%(d['value'], d['value_exact_march'], d['unit'], d['ms_per_step'], d['e2e']['value']), d['clocks'], d.get('collective'))
    except Exception as e: print(f, 'ERR', e)
PY
