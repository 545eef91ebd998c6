"""CPU tests of host-side logic that needs no GPU: sizing entry points of the C ABI and the dispatch rules of the
reference-side hook (scenedreamer_b200.integration.fused_forward_perpix) with the compute calls stubbed out."""
import types

import pytest
import torch

from scenedreamer_b200 import _lib, integration, render


def test_sizing_entry_points_without_gpu():
    L = _lib.lib()
    assert L.sdb_render_train_record_bytes(1, 262, 262, 24) > L.sdb_render_train_record_bytes(1, 128, 128, 24) > 0
    assert L.sdb_render_train_record_bytes(0, 262, 262, 24) == 0 and L.sdb_render_train_record_bytes(1, 262, 262, 65) == 0
    assert L.sdb_render_backward_workspace_bytes(1, 262, 262, 24, 16, 19) > (16 << 19) * 32      # holds the table gradient
    assert L.sdb_sky_train_record_bytes(1, 262, 262) > 0 and L.sdb_sky_backward_workspace_bytes(1, 262, 262) > 0
    assert L.sdb_mlp_backward_pack_bytes() > 0 and L.sdb_sky_backward_pack_bytes() > 0
    import ctypes
    dims = (ctypes.c_int64 * 3)(103, 1024, 1024)
    assert L.sdb_height_bound_elems(dims, 4) == 64 * 64 and L.sdb_height_bound_elems(dims, 1) == 0
    # compute entry points refuse bad arguments before touching the device
    assert L.sdb_render_rays_train_forward(None, None, None) != 0
    assert L.sdb_render_rays_backward(None, None, None, None) != 0
    assert L.sdb_sky_backward(1, 8, 8, None, None, None, None, None, None, None, None) != 0


class _FakeTensor:
    """Just enough of a CUDA tensor for the dispatch rules (shape, is_cuda, slicing, == 0)."""

    def __init__(self, t):
        self.t = t
        self.shape = t.shape
        self.is_cuda = True
        self.device = 'cuda:0'
        self._base = None

    def __getitem__(self, idx):
        return _FakeTensor(self.t[idx])

    def __eq__(self, other):
        return self.t == other

    def contiguous(self):
        return self


def _fake_generator(n_views, monkeypatch, calls):
    class Gen:                                               # stands in for imaginaire.generators.scenedreamer.Generator
        def _forward_perpix(self, *a):
            calls.append('reference')
            return ('ref',)

        def forward(self, *a):
            return 'fwd'

    integration.install(Gen)
    assert Gen._forward_perpix is integration.fused_forward_perpix and Gen().forward() == 'fwd'
    gen = Gen()
    lin = torch.nn.Linear(4, 4)
    gen.render_net, gen.sky_net, gen.hash_encoder = lin, torch.nn.Linear(2, 2), torch.nn.Linear(2, 2)
    gen.hash_encoder.per_level_scale, gen.hash_encoder.base_resolution = 1.38, 16
    gen.hash_encoder.log2_hashmap_size, gen.hash_encoder.num_levels = 19, 16
    gen.voxel = types.SimpleNamespace(voxel_t=torch.zeros(4, 8, 8))
    gen.label_trans = types.SimpleNamespace(mcid2rdid_lut=torch.zeros(4, dtype=torch.long), ignore_id=0, dirt_id=3)
    gen.clip_feat_map, gen.keep_sky_out, gen.keep_sky_out_avgpool, gen.sky_global_avgpool = True, True, True, True
    gen.sample_use_box_boundaries, gen.raw_noise_std = False, 0.0
    gen.pe_params, gen.pe_params_sky = [0, 0, 0, False], [5, True]
    gen.coarse_deterministic_sampling, gen.num_samples, gen.sample_depth, gen.dists_scale = True, 4, 3, 0.25
    st = integration._state(gen)

    class R:
        def forward(self, *a, **k):
            calls.append('inference')
            n = a[0].shape[0]
            z = torch.zeros(n, 2, 2)
            return dict(net_out=torch.zeros(n, 2, 2, 64), total_weight=z, weights=z, rand_depth=z, sky=torch.zeros(n, 2, 2, 64))

        def invalidate(self):
            calls.append('invalidate')
    st.get_renderer = lambda g: R()

    def fake_train(P, vid, *a, **k):
        calls.append('train')
        z = torch.zeros(1, 2, 2)
        return dict(net_out=torch.zeros(1, 2, 2, 64), total_weight=z, weights=z, rand_depth=z, sky=torch.zeros(1, 2, 2, 64))
    monkeypatch.setattr(render, 'render_rays_train', fake_train)
    vid = _FakeTensor(torch.ones(n_views, 2, 2, 6, 1, dtype=torch.int32))
    dep = _FakeTensor(torch.zeros(n_views, 2, 2, 2, 6, 1))
    rd = _FakeTensor(torch.zeros(n_views, 2, 2, 1, 3))
    return gen, vid, dep, rd


def test_hook_dispatch_rules(monkeypatch):
    calls = []
    gen, vid, dep, rd = _fake_generator(1, monkeypatch, calls)
    ori, z, genc = torch.zeros(1, 3), torch.zeros(1, 4), torch.zeros(1, 2)
    f = integration.fused_forward_perpix
    with torch.no_grad():
        out = f(gen, None, vid, dep, rd, ori, z, genc)
    assert calls == ['inference'] and len(out) == 12
    calls.clear()
    out = f(gen, None, vid, dep, rd, ori, z, genc)                 # parameters require grad -> recording forward + fused backward
    assert calls == ['train'] and len(out) == 12
    calls.clear()
    gen2, vid2, dep2, rd2 = _fake_generator(3, monkeypatch, calls)
    out = f(gen2, None, vid2, dep2, rd2, torch.zeros(3, 3), torch.zeros(3, 4), genc)
    assert calls == ['train'] * 3 and out[0].shape[0] == 3           # a batch = one recorded pass per view
    calls.clear()
    gen.sky_avg = torch.zeros(1, 64)                                 # pre-set sky mean under autograd: reference composition
    f(gen, None, vid, dep, rd, ori, z, genc)
    assert calls == ['reference']
    calls.clear()
    del gen.sky_avg
    gen.raw_noise_std = 0.5                                          # option outside the fused path
    with torch.no_grad():
        f(gen, None, vid, dep, rd, ori, z, genc)
    assert calls == ['reference']
    calls.clear()
    gen.raw_noise_std = 0.0
    for q in list(gen.render_net.parameters()) + list(gen.hash_encoder.parameters()) + list(gen.sky_net.parameters()):
        q.requires_grad_(False)
    f(gen, None, vid, dep, rd, ori, z, genc)                        # nothing to differentiate: inference kernel even with grad mode on
    assert calls == ['inference']


def test_hook_gates_and_epochs(monkeypatch):
    calls = []
    gen, vid, dep, rd = _fake_generator(1, monkeypatch, calls)
    ori, z, genc = torch.zeros(1, 3), torch.zeros(1, 4), torch.zeros(1, 2)
    f = integration.fused_forward_perpix
    with torch.no_grad():
        f(gen, None, vid, dep, rd, ori, None, genc)                  # style_dims == 0: no style code -> reference composition
        assert calls == ['reference']
        calls.clear()
        gen.pe_params = [0, 0, 0, True]                              # view direction fed to the MLP: not covered
        f(gen, None, vid, dep, rd, ori, z, genc)
        assert calls == ['reference']
        calls.clear()
        gen.pe_params = [0, 0, 0, False]
        monkeypatch.setenv('SDB200_FUSED', '0')                      # switch: the reference's own composition
        f(gen, None, vid, dep, rd, ori, z, genc)
        assert calls == ['reference']
        monkeypatch.delenv('SDB200_FUSED')
    calls.clear()
    for q in list(gen.render_net.parameters()) + list(gen.hash_encoder.parameters()):
        q.requires_grad_(False)
    f(gen, None, vid, dep, rd, ori, z, genc)                         # only sky_net trains: still the differentiable path
    assert calls == ['train']
    # every public entry starts a new epoch
    st = integration._state(gen)
    e = st.epoch
    assert gen.forward() == 'fwd' and st.epoch == e + 1
    # wrappers: WrappedModel / DDP (.module), ModelAverage (.averaged_model), nested
    inner = types.SimpleNamespace(module=types.SimpleNamespace(averaged_model=types.SimpleNamespace(module=gen)))
    assert integration._unwrap(inner) is gen
    with pytest.raises(TypeError):
        integration.patch_generator(types.SimpleNamespace(module=object()))
    integration.invalidate(inner)
    assert st.epoch == e + 2


def test_renderer_caches_key_on_identity_not_address(monkeypatch):
    """a new style tensor that happens to reuse the freed tensor's address must repack."""
    built = []
    monkeypatch.setattr(render, 'pack_mlp', lambda P, z, prec: built.append(float(z[0, 0])) or ('pack', len(built)))
    r = render.FusedPerPixelRenderer({}, (4, 8, 8), torch.zeros(4, dtype=torch.int32), 1.38)
    ptrs = set()
    for k in range(4):
        z = torch.full((1, 256), float(k))
        ptrs.add(z.data_ptr())
        assert r.pack_for(z) == ('pack', k + 1)
        assert r.pack_for(z) == ('pack', k + 1)                      # same object, same version: cached
        del z
    assert built == [0.0, 1.0, 2.0, 3.0]
    z = torch.zeros(1, 256)
    p1 = r.pack_for(z)
    z.add_(1.0)                                                      # in-place edit bumps the version counter
    assert r.pack_for(z) != p1
    r.invalidate()
    assert r.pack_for(z) != ('pack', len(built) - 1)


def test_style_fold_is_modlinear_for_one_style_code():
    """The fold the fused renderer consumes (render.modulated_weights: W' = W * alpha(z), bias beta(z)) is ModLinear's forward
    (layers.py:247-260, restated in oracle.mod_linear) for one style code -- here the torch formulation on the CPU; the fused CUDA
    fold (csrc/modulate.cu) is compared with this formulation, values and all gradients, in tests/test_gpu_train.py."""
    import torch

    import oracle
    from scenedreamer_b200 import render
    P = oracle.make_params(seed=4, stress=True, table_entries=64)
    g = torch.Generator().manual_seed(1)
    z = oracle.style_mlp(torch.randn(1, 128, generator=g), P)            # [1, 256]
    x = torch.randn(1, 37, 256, generator=g)
    wh, bh = render.modulated_weights(P, z[0])                           # CPU tensors -> the torch formulation
    assert wh.shape == (5, 256, 256) and bh.shape == (5, 256)
    for l, k in enumerate((2, 3, 4, 5, 6)):
        ref = oracle.mod_linear(x, z, P, 'render_net.fc_%d' % k)
        got = x[0] @ wh[l].t() + bh[l]
        assert float((got - ref[0]).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_new_entry_points_refuse_bad_arguments_before_touching_the_device():
    """Round-2 entry points (banded raycast, float16 grid tables, ModLinear fold, sp_trilinear, Adam): argument checks come before
    any CUDA call, so they can be exercised without a GPU.  Pointers handed over here are never dereferenced."""
    import ctypes
    L = _lib.lib()
    EINVAL, EUNSUPPORTED = -1, -2
    i64x3 = ctypes.c_int64 * 3
    f3, f2, i2, i3 = ctypes.c_float * 3, ctypes.c_float * 2, ctypes.c_int32 * 2, ctypes.c_int32 * 3
    dummy = ctypes.c_void_p(0x1000)                                    # non-null, never read: every call below fails a check first
    dims, strides = i64x3(8, 8, 8), i64x3(64, 8, 1)
    o, d, u = f3(1, 1, 1), f3(0, 1, 0), f3(1, 0, 0)

    def bands(band, img=(16, 8)):
        return L.sdb_ray_voxel_intersection_perspective_bands(dummy, dims, strides, o, d, u, 10.0, f2(4, 4), i2(*img), 4, band, dummy,
                                                              dummy, dummy, None, 0, None)
    assert bands(None) == EINVAL                                       # no band description
    assert bands(i3(0, 16, 8)) == EINVAL                               # overlapping bands (stride < rows)
    assert bands(i3(-1, 8, 8)) == EINVAL and bands(i3(0, 0, 8)) == EINVAL
    assert bands(i3(0, 8, 8), img=(0, 8)) == EINVAL                    # empty image
    # float16 grid tables: odd C never reaches this path in the reference (grid.py:38)
    assert L.sdb_grid_encode_forward_f16(dummy, dummy, dummy, dummy, 16, 3, 1, 4, 1.0, 16, 0, None, 0, 0, None) == EUNSUPPORTED
    assert L.sdb_grid_encode_forward_f16(dummy, dummy, dummy, dummy, 16, 6, 2, 4, 1.0, 16, 0, None, 0, 0, None) == EUNSUPPORTED
    assert L.sdb_grid_encode_forward_f16(None, dummy, dummy, dummy, 16, 3, 2, 4, 1.0, 16, 0, None, 0, 0, None) == EINVAL
    assert L.sdb_grid_encode_forward_f16(dummy, dummy, dummy, dummy, 16, 3, 2, 4, 1.0, 16, 1, None, 0, 0, None) == EINVAL   # dy_dx asked, not given
    assert L.sdb_grid_encode_backward_f16(dummy, dummy, dummy, dummy, dummy, 16, 3, 1, 4, 1.0, 16, 0, None, None, 0, 0, None) == EUNSUPPORTED
    assert L.sdb_grid_encode_forward_f16(dummy, dummy, dummy, dummy, 0, 3, 2, 4, 1.0, 16, 0, None, 0, 0, None) == 0          # empty batch: nothing to do
    # ModLinear fold
    ptrs = (ctypes.c_void_p * 25)(*([0x1000] * 25))
    assert L.sdb_modulate_forward(None, dummy, 256, 256, 256, dummy, dummy, dummy, None) == EINVAL
    assert L.sdb_modulate_forward(ptrs, dummy, 0, 256, 256, dummy, dummy, dummy, None) == EINVAL
    holes = (ctypes.c_void_p * 25)(*([0x1000] * 24 + [None]))
    assert L.sdb_modulate_forward(holes, dummy, 256, 256, 256, dummy, dummy, dummy, None) == EINVAL
    assert L.sdb_modulate_backward(ptrs, holes, dummy, dummy, dummy, dummy, 256, 256, 256, dummy, dummy, None) == EINVAL
    assert L.sdb_modulate_backward(ptrs, ptrs, dummy, dummy, dummy, None, 256, 256, 256, dummy, dummy, None) == EINVAL
