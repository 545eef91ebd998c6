#!/bin/bash
# targeted re-check: grid encoder (float16 tables, single-pass dy_dx), sampler, banded raycast, ray-slot test
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
timeout 900 python -m pytest tests -m gpu -q -s -k "grid_encode or gridencoder or ray_slots or sampler or row_bands or dda" 2>&1 | tail -60 > gpurun_out/pytest_gpu_o.log
tail -5 gpurun_out/pytest_gpu_o.log; grep -E "FAILED|Error|f16 table" gpurun_out/pytest_gpu_o.log | head -20
for m in 1 2; do SDB_GRIDENC_MINB=$m PYTHONPATH=. timeout 120 python tools/gridenc_run.py 2>&1 | tail -1; done | tee gpurun_out/gridenc_minb.log
timeout 600 python tests/ops_timing.py > gpurun_out/ops_timing.json 2> gpurun_out/ops_timing.err; cut -c1-1600 gpurun_out/ops_timing.json; tail -3 gpurun_out/ops_timing.err
