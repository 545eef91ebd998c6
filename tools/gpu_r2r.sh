#!/bin/bash
# diagnostics build (clock stamps compiled in) -> per-layer timeline of the render kernel; the product build is restored afterwards
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SDB_NVCC_EXTRA=-DSDB_TIMELINE python -m scenedreamer_b200.build > gpurun_out/build_tl.log 2>&1 || { tail -20 gpurun_out/build_tl.log; exit 1; }
PYTHONPATH=. timeout 300 python tools/render_timeline.py 60 > gpurun_out/render_timeline.txt 2> gpurun_out/render_timeline.err; cat gpurun_out/render_timeline.txt; tail -3 gpurun_out/render_timeline.err
python -m scenedreamer_b200.build > gpurun_out/build.log 2>&1
