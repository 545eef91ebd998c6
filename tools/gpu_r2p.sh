#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
export OMP_WAIT_POLICY=PASSIVE GOMP_SPINCOUNT=0
PYTHONPATH=. timeout 300 python tools/render_timeline.py 60 > gpurun_out/render_timeline.txt 2> gpurun_out/render_timeline.err; cat gpurun_out/render_timeline.txt; tail -3 gpurun_out/render_timeline.err
timeout 600 python -m pytest tests -m gpu -q -k "sp_trilinear or render or fullsize" 2>&1 | tail -4
timeout 300 python tests/ops_timing.py > gpurun_out/ops_timing.json 2> gpurun_out/ops_timing.err; cut -c1-1800 gpurun_out/ops_timing.json; tail -2 gpurun_out/ops_timing.err
