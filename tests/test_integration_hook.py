"""Zero-edit arming of the fused path (CPU, no compute): with dropin/ on the path, importing the reference's generator
module and making the first drop-in call patches `Generator` at class level; the frame detector recognises the tiles
`inference_givenstyle` cuts (scenedreamer.py:600-617) as windows of one raycast result."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))

SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
from oracle import refgen
refgen.setup('dropin')
import torch
import imaginaire.generators.scenedreamer as m          # imports `voxlib` (= dropin/voxlib.py) at its line 13
from scenedreamer_b200 import integration
import voxlib
assert voxlib.__file__.startswith(%(root)r + '/dropin/')
try:                                                     # first drop-in call of a run (refuses the CPU tensor, after arming)
    voxlib.positional_encoding(torch.zeros(2, 3), 2, -1, True)
except RuntimeError:
    pass
assert m.Generator._forward_perpix is integration.fused_forward_perpix
assert hasattr(m.Generator.inference_givenstyle, '_sdb200_wrapped') and hasattr(m.Generator.forward, '_sdb200_wrapped')
gen, cfg = refgen.build_generator(64, 'cpu')             # the real class instantiates from configs/scenedreamer_inference.yaml
assert type(gen) is m.Generator and integration.supported(gen, type('T', (), {'is_cuda': True})(), torch.zeros(1, 256), torch.zeros(1, 2))
integration.uninstall(m.Generator)
assert m.Generator._forward_perpix is m.Generator.__dict__['_forward_perpix'] and not hasattr(m.Generator.forward, '_sdb200_wrapped')
# the import-hook route: `voxlib` imported first (any module that imports gancraft voxlib earlier), generator module afterwards
del sys.modules['imaginaire.generators.scenedreamer']
integration._installed = False
integration.install_import_hook()
import imaginaire.generators.scenedreamer as m2
assert m2.Generator._forward_perpix is integration.fused_forward_perpix
print('HOOK-OK')
'''


def test_zero_edit_hook_arms_itself():
    from oracle import refgen
    if refgen.reference_python_root() is None:
        pytest.skip('reference Python not available (oracle/_ref/py, /root/reference)')
    p = subprocess.run([sys.executable, '-c', SCRIPT % {'root': ROOT}], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0 and 'HOOK-OK' in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]


def test_frame_window_detection():
    import torch
    from scenedreamer_b200 import integration
    H, W, M = 57, 99, 6
    vid = torch.zeros(H, W, M, 1, dtype=torch.int32)
    dep = torch.zeros(2, H, W, M, 1)
    rd = torch.zeros(H, W, 1, 3)
    v, d, r = vid.unsqueeze(0), dep.unsqueeze(0), rd.unsqueeze(0)
    for (h0, h1, w0, w1) in ((10, 40, 20, 60), (0, 30, 0, 40), (27, 57, 59, 99)):
        win = integration._frame_window(v[:, h0:h1, w0:w1], d[:, :, h0:h1, w0:w1], r[:, h0:h1, w0:w1])
        assert win[1:] == (h0, w0, h1 - h0, w1 - w0)
        assert win[0][0] is vid and win[0][1] is dep and win[0][2] is rd
    assert integration._frame_window(v, d, r) is None                                             # the whole frame
    assert integration._frame_window(v[:, 10:40, 20:60].contiguous(), d[:, :, 10:40, 20:60], r[:, 10:40, 20:60]) is None
    assert integration._frame_window(v[:, 10:40, 20:60], d[:, :, 11:41, 20:60], r[:, 10:40, 20:60]) is None   # not the same window
    assert integration._frame_window(v[:, 10:40:2, 20:60], d[:, :, 10:40:2, 20:60], r[:, 10:40:2, 20:60]) is None   # strided rows
