#!/bin/bash
# fused style-modulation fold: training tests + train bench (fused fold on / off)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
timeout 900 python -m pytest tests -m gpu -q -x -k "train or generator" 2>&1 | tail -5
for v in 1 0; do
  SDB200_FUSED_MOD=$v timeout 300 python bench_train.py --steps 10 --warmup 3 --no-composition > gpurun_out/train_mod$v.json 2> gpurun_out/train_mod$v.err; echo "fused_mod=$v $(cut -c1-330 gpurun_out/train_mod$v.json)"; tail -1 gpurun_out/train_mod$v.err
done
