"""Host side of the tensor-core RenderCNN (libsdb200: sdb_cnn_pack / sdb_cnn_forward).

Mirrors Base3DGenerator._forward_global (imaginaire/generators/gancraft_base.py:588-603): per-pixel feature map
[N,H,W,64] + style code -> (tanh image, raw image) [N,3,H,W], with RenderCNN.forward (:201-225) evaluated once on the
whole frame.  torch allocates; the style modulation vector fc_z_cond(z) (one 256x1024 GEMV) is computed in torch.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import _lib

PRECISION_FP16 = 0      # one fp16 pass per product (the class of the reference's default, cuDNN TF32)
PRECISION_FP16X3 = 2    # fp16 hi/lo split, 3 passes: fp32-grade (parity default)

_NAMES = ('conv1.weight', 'conv1.bias', 'conv2a.weight', 'conv2a.bias', 'conv2b.weight', 'conv3a.weight', 'conv3a.bias',
          'conv3b.weight', 'conv4a.weight', 'conv4a.bias', 'conv4b.weight', 'conv4b.bias', 'conv4.weight', 'conv4.bias')


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def supported(P, prefix='denoiser.'):
    """The kernel is specialised for SceneDreamer's RenderCNN: 64 -> 256 hidden -> 3."""
    want = {'conv1.weight': (256, 64, 1, 1), 'conv2a.weight': (256, 256, 3, 3), 'conv2b.weight': (256, 256, 3, 3),
            'conv3a.weight': (256, 256, 3, 3), 'conv3b.weight': (256, 256, 3, 3), 'conv4a.weight': (256, 256, 1, 1),
            'conv4b.weight': (256, 256, 1, 1), 'conv4.weight': (3, 256, 1, 1), 'fc_z_cond.weight': (1024, None)}
    for k, shp in want.items():
        t = P.get(prefix + k)
        if t is None or len(t.shape) != len(shp) or any(a is not None and a != b for a, b in zip(shp, t.shape)):
            return False
    return True


class RenderCNNEngine:
    """P: dict with the reference's state-dict names `denoiser.*` (CUDA fp32).  Packs the weights once; keeps one
    workspace per frame size."""

    def __init__(self, P, precision=PRECISION_FP16X3, prefix='denoiser.'):
        if not supported(P, prefix):
            raise RuntimeError('RenderCNNEngine: unexpected denoiser shapes (expects conv1 64->256, 3x3 256->256, conv4 256->3)')
        self.P, self.prefix, self.precision = P, prefix, int(precision)
        self._pack = None
        self._ws = {}

    def invalidate(self):
        self._pack = None

    def pack(self):
        if self._pack is None:
            L = _lib.lib()
            ts = [self.P[self.prefix + n].detach().to(torch.float32).contiguous() for n in _NAMES]
            dev = ts[0].device
            pack = torch.empty(int(L.sdb_cnn_pack_bytes(self.precision)), dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev):
                _lib.check(L.sdb_cnn_pack(*[_ptr(t) for t in ts], self.precision, _ptr(pack), _stream(dev)), 'sdb_cnn_pack')
            self._pack = pack
        return self._pack

    def modulation(self, z):
        """fc_z_cond(z) -> [N, 4, 256]: (w, b) of the two modulated blocks (gancraft_base.py:208-209)."""
        p = self.prefix
        return F.linear(z, self.P[p + 'fc_z_cond.weight'], self.P[p + 'fc_z_cond.bias']).reshape(z.shape[0], 4, 256).contiguous()

    def forward(self, net_out, z, want_raw=True):
        """net_out [N,H,W,64] fp32 CUDA, z [N,256] (or [1,256]) -> (fake_images [N,3,H,W], fake_images_raw or None)."""
        if not net_out.is_cuda or net_out.dtype != torch.float32 or net_out.dim() != 4 or net_out.shape[-1] != 64:
            raise RuntimeError('net_out must be a float32 CUDA tensor [N,H,W,64]')
        L = _lib.lib()
        dev = net_out.device
        N, H, W = net_out.shape[:3]
        x = net_out.contiguous()
        mod = self.modulation(z.detach().to(dev, torch.float32))
        pack = self.pack()
        rgb = torch.empty(N, 3, H, W, dtype=torch.float32, device=dev)
        raw = torch.empty(N, 3, H, W, dtype=torch.float32, device=dev) if want_raw else None
        key = (H, W, self.precision, str(dev))
        ws = self._ws.get(key)
        ready = 1
        if ws is None:
            self._ws.clear()                                     # one frame size at a time (a few hundred MB to ~2 GB)
            ws = self._ws[key] = torch.empty(int(L.sdb_cnn_workspace_bytes(H, W, self.precision)), dtype=torch.uint8, device=dev)
            ready = 0
        with torch.cuda.device(dev):
            for i in range(N):
                m = mod[i if mod.shape[0] > 1 else 0]
                _lib.check(L.sdb_cnn_forward(_ptr(x[i]), H, W, _ptr(pack), _ptr(m), self.precision, _ptr(rgb[i]),
                                             _ptr(raw[i]) if want_raw else None, _ptr(ws), ready, _stream(dev)), 'sdb_cnn_forward')
                ready = 1
        return rgb, raw
