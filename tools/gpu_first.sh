#!/bin/bash
# first GPU smoke: boundary ops parity + tcgen05 layout diagnostics
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/first_pytest.log
timeout 120 python - > gpurun_out/tc_variants.log 2>&1 <<'PY'
import torch
from scenedreamer_b200 import ops
for bf16 in (False, True):
    for variant in (0, 1):
        a = torch.randn(128, 64, device='cuda'); b = torch.randn(32, 64, device='cuda')
        try:
            c = ops.tc_selftest(a, b, bf16=bf16, variant=variant); torch.cuda.synchronize()
            lo = torch.bfloat16 if bf16 else torch.float16
            ref = a.to(lo).float() @ b.to(lo).float().t()
            print('bf16', bf16, 'variant', variant, 'maxerr', float((c - ref).abs().max()), flush=True)
        except Exception as e:
            print('bf16', bf16, 'variant', variant, 'EXC', e, flush=True)
PY
cat gpurun_out/first_pytest.log; cat gpurun_out/tc_variants.log
