"""GPU tests of the drop-in boundary: gridencoder.GridEncoder autograd (forward + both backward
kernels) and the reference-side integration hook that replaces Generator._forward_perpix."""
import os
import sys
import types

import numpy as np
import pytest
import torch

import oracle
from scenedreamer_b200 import integration, ops, render, synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
DEV = 'cuda:0'
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


@pytest.fixture(scope='module')
def gridencoder_pkg():
    """The REFERENCE's own `gridencoder` package (staged copy, unmodified) on top of dropin/_gridencoder.py: its grid.py
    does `import _gridencoder as _backend` (gridencoder/grid.py:9-12), so no copy of it lives in this repository."""
    from oracle import refgen
    ref = refgen.reference_python_root()
    if ref is None:
        pytest.skip('reference Python not staged (oracle/build_ref.py)')
    sys.path.insert(0, os.path.join(ROOT, 'dropin'))
    sys.path.insert(1, ref)
    import gridencoder
    assert gridencoder.grid._backend.__file__.startswith(os.path.join(ROOT, 'dropin'))
    yield gridencoder
    sys.path.remove(os.path.join(ROOT, 'dropin'))
    sys.path.remove(ref)


def device_level_scales(L, pls, base):
    S = torch.tensor(float(np.float32(np.log2(pls))), device=DEV)
    return (torch.exp2(torch.arange(L, device=DEV, dtype=torch.float32) * S) * float(base) - 1.0).cpu()


@pytest.mark.parametrize('D,L,C,base,log2T,desired', [(5, 16, 8, 16, 19, 2048), (3, 8, 2, 4, 12, 64)])
def test_gridencoder_module_forward_backward(gridencoder_pkg, D, L, C, base, log2T, desired):
    ge = gridencoder_pkg.GridEncoder(input_dim=D, num_levels=L, level_dim=C, base_resolution=base,
                                     log2_hashmap_size=log2T, desired_resolution=desired).to(DEV)
    g = torch.Generator().manual_seed(D)
    ge.embeddings.data = ((torch.rand(ge.embeddings.shape, generator=g) * 2 - 1) * 0.1).to(DEV)
    x = (torch.rand(2, 500, D, generator=g) * 2 - 1)
    xd = x.to(DEV).requires_grad_(True)
    y = ge(xd)
    assert y.shape == (2, 500, L * C)
    ls = device_level_scales(L, ge.per_level_scale, base)
    ref = oracle.grid_encoder_module_forward(x, ge.embeddings.detach().cpu(), ge.offsets.cpu(), ge.per_level_scale, base,
                                             level_scales=ls)
    np.testing.assert_allclose(y.detach().cpu().numpy(), ref.numpy(), rtol=1e-5, atol=2e-7)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy.to(DEV))
    x01 = ((x + 1) / 2).reshape(-1, D)
    _, dy_dx = oracle.grid_encode_forward(x01, ge.embeddings.detach().cpu(), ge.offsets.cpu(), ge.per_level_scale, base, True,
                                          level_scales=ls)
    gl = gy.reshape(-1, L, C).permute(1, 0, 2).contiguous()
    ege, egi = oracle.grid_encode_backward(gl, x01, ge.embeddings.detach().cpu(), ge.offsets.cpu(), ge.per_level_scale, base,
                                           dy_dx, level_scales=ls)
    np.testing.assert_allclose(ge.embeddings.grad.cpu().numpy(), ege.numpy(), rtol=1e-4, atol=1e-5)
    # d/dx of the [-1,1] -> [0,1] mapping is 1/2
    np.testing.assert_allclose(xd.grad.cpu().reshape(-1, D).numpy(), 0.5 * egi.numpy(), rtol=1e-3,
                               atol=1e-4 * float(egi.abs().max()))
    # no input grad requested -> only the embedding grad kernel runs
    ge.zero_grad()
    ge(x.to(DEV)).backward(gy.to(DEV))
    np.testing.assert_allclose(ge.embeddings.grad.cpu().numpy(), ege.numpy(), rtol=1e-4, atol=1e-5)


def test_gridencoder_module_under_autocast(gridencoder_pkg):
    """AMP on: the reference's grid.py hands the drop-in a float16 table (grid.py:38-39); outputs come back half, within half
    precision of the float32 module, and both gradients arrive (table gradient in float32 through autograd's cast)."""
    ge = gridencoder_pkg.GridEncoder(input_dim=5, num_levels=16, level_dim=8, base_resolution=16, log2_hashmap_size=19,
                                     desired_resolution=2048).to(DEV)
    g = torch.Generator().manual_seed(9)
    ge.embeddings.data = ((torch.rand(ge.embeddings.shape, generator=g) * 2 - 1) * 0.1).to(DEV)
    x = (torch.rand(3000, 5, generator=g) * 2 - 1).to(DEV).requires_grad_(True)
    y32 = ge(x).detach()
    with torch.autocast('cuda', dtype=torch.float16):
        y16 = ge(x)
    assert y16.dtype == torch.float16 and y16.shape == y32.shape
    assert float((y16.float() - y32).abs().max()) < 1e-2 * float(y32.abs().max()) + 1e-4
    gy = torch.randn(y16.shape, generator=g).to(DEV) * 0.1
    y16.backward(gy.half())
    assert ge.embeddings.grad is not None and ge.embeddings.grad.dtype == torch.float32 and x.grad is not None
    g16, gx16 = ge.embeddings.grad.clone(), x.grad.clone()
    ge.zero_grad()
    x.grad = None
    ge(x).backward(gy)
    assert float((g16 - ge.embeddings.grad).abs().max()) < 2e-2 * float(ge.embeddings.grad.abs().max())
    assert float((gx16 - x.grad).abs().max()) < 5e-2 * float(x.grad.abs().max())


def test_patched_forward_perpix_matches_oracle(golden_ops):
    """A duck-typed stand-in for the reference Generator (same attribute names as
    imaginaire/generators/scenedreamer.py / gancraft_base.py) through integration.patch_generator."""
    world = synth.SyntheticVoxelWorld(size=128, seed=7)
    pose = synth.eval_camera_poses(world, maxstep=8, pattern=0)[2]
    o, d, u, f, c, res = synth.frame_camera(world, pose, resolution_hw=(36, 52), pad=4)
    vid, dep, rd = ops.ray_voxel_intersection_perspective(world.voxel_t.to(DEV), o, d, u, f, c, res, 6)
    P = oracle.make_params(seed=44, stress=True)
    offsets, pls = oracle.grid_offsets()

    class Holder(torch.nn.Module):
        def __init__(self, sd):
            super().__init__()
            for k, v in sd.items():
                self.register_parameter(k.replace('.', '__'), torch.nn.Parameter(v.to(DEV)))

        def state_dict(self, *a, **k):
            return {n.replace('__', '.'): b.detach() for n, b in super().named_parameters()}

        def named_parameters(self, *a, **k):
            return [(n.replace('__', '.'), b) for n, b in super().named_parameters(*a, **k)]

    def sub(prefix):
        return Holder({k[len(prefix) + 1:]: v for k, v in P.items() if k.startswith(prefix + '.')})

    called = {}

    class Gen:                                                # duck-typed stand-in; the real class: tests/test_gpu_generator.py
        def _forward_perpix(self, *a):
            called['ref'] = True

    gen = Gen()
    gen.render_net, gen.sky_net, gen.hash_encoder = sub('render_net'), sub('sky_net'), sub('hash_encoder')
    gen.hash_encoder.per_level_scale, gen.hash_encoder.base_resolution = pls, 16
    gen.hash_encoder.log2_hashmap_size, gen.hash_encoder.num_levels = 19, 16
    gen.voxel = types.SimpleNamespace(voxel_t=world.voxel_t.to(DEV))
    gen.label_trans = types.SimpleNamespace(mcid2rdid_lut=torch.from_numpy(golden_ops['mc2reduced_lut']).long(),
                                            ignore_id=0, dirt_id=3)
    gen.clip_feat_map, gen.keep_sky_out, gen.keep_sky_out_avgpool, gen.sky_global_avgpool = True, True, True, True
    gen.sample_use_box_boundaries, gen.raw_noise_std = False, 0.0
    gen.pe_params, gen.pe_params_sky = [0, 0, 0, False], [5, True]
    gen.coarse_deterministic_sampling, gen.num_samples, gen.sample_depth, gen.dists_scale = True, 24, 3, 0.25
    integration.patch_generator(gen)
    g = torch.Generator().manual_seed(8888)
    z = oracle.style_mlp(torch.randn(1, 128, generator=g), P)
    genc = torch.tanh(torch.randn(1, 2, generator=g))
    with torch.no_grad():
        ret = gen._forward_perpix(None, vid.unsqueeze(0), dep.unsqueeze(0), rd.unsqueeze(0), o.unsqueeze(0).to(DEV),
                                  z.to(DEV), genc.to(DEV))
    torch.cuda.synchronize()
    assert len(ret) == 12 and 'ref' not in called
    S = torch.tensor(float(np.float32(np.log2(pls))), device=DEV)
    ls = (torch.exp2(torch.arange(16, device=DEV, dtype=torch.float32) * S) * 16.0 - 1.0).cpu()
    ref = oracle.forward_perpix(P, vid.unsqueeze(0).cpu(), dep.unsqueeze(0).cpu(), rd.unsqueeze(0).cpu(), o.unsqueeze(0), z, genc,
                                list(world.voxel_t.shape), torch.from_numpy(golden_ops['mc2reduced_lut']), offsets, pls,
                                level_scales=ls)
    assert float((ret[0].cpu() - ref['net_out']).abs().max()) <= 1e-3
    assert float((ret[2].cpu() - ref['weights']).abs().max()) <= 1e-3                      # weights [N,H,W,S,1]
    np.testing.assert_allclose(ret[4].cpu().numpy(), ref['rand_depth'].numpy(), rtol=1e-6, atol=1e-5)   # rand_depth
    assert float((ret[3].cpu() - ref['total_weights']).abs().max()) <= 1e-3
    assert torch.equal(ret[9].cpu(), ref['sky_mask']) and torch.equal(ret[10].cpu(), ref['sky_only_mask'])
    # depth variant of the reference: sum(weights * rand_depth) (scenedreamer.py:816)
    dmap = torch.sum(ret[2] * ret[4], dim=-2)
    assert float((dmap.cpu() - ref['depth_map']).abs().max()) <= 1e-3
    # with autograd enabled the hook runs the recording forward + fused backward: gradients land on the
    # module's own Parameters and on z / global_enc (gen_update of trainers/gancraft.py)
    zg, gg = z.clone().to(DEV).requires_grad_(True), genc.clone().to(DEV).requires_grad_(True)
    ret_t = gen._forward_perpix(None, vid.unsqueeze(0), dep.unsqueeze(0), rd.unsqueeze(0), o.unsqueeze(0).to(DEV), zg, gg)
    assert len(ret_t) == 12 and 'ref' not in called and ret_t[0].requires_grad
    assert float((ret_t[0].detach() - ret[0]).abs().max()) <= 1e-3
    G = torch.randn(ret_t[0].shape, generator=torch.Generator().manual_seed(2)).to(DEV)
    (ret_t[0] * G).sum().backward()
    Pc = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    zc, gc = z.clone().requires_grad_(True), genc.clone().requires_grad_(True)
    ref_t = oracle.forward_perpix_autograd(Pc, vid.unsqueeze(0).cpu(), dep.unsqueeze(0).cpu(), rd.unsqueeze(0).cpu(), o.unsqueeze(0),
                                           zc, gc, list(world.voxel_t.shape), torch.from_numpy(golden_ops['mc2reduced_lut']),
                                           offsets, pls, level_scales=ls)
    (ref_t * G.cpu()).sum().backward()
    got = dict(gen.render_net.named_parameters())
    for name, a, b in (('z', zg.grad, zc.grad), ('global_enc', gg.grad, gc.grad),
                       ('embeddings', dict(gen.hash_encoder.named_parameters())['embeddings'].grad, Pc['hash_encoder.embeddings'].grad),
                       ('fc_1.weight', got['fc_1.weight'].grad, Pc['render_net.fc_1.weight'].grad),
                       ('fc_4.weight_alpha', got['fc_4.weight_alpha'].grad, Pc['render_net.fc_4.weight_alpha'].grad),
                       ('sky fc3.weight', dict(gen.sky_net.named_parameters())['fc3.weight'].grad, Pc['sky_net.fc3.weight'].grad)):
        rel = float((a.cpu().double() - b.double()).norm() / b.double().norm())
        print('patched train grad %-20s rel-L2 %.3e' % (name, rel))
        assert rel <= 1e-2, name
    # a batch of two views under autograd = two recorded passes, still the fused path; the gradients add up
    for q in list(gen.render_net.parameters()) + list(gen.hash_encoder.parameters()) + list(gen.sky_net.parameters()):
        q.grad = None
    two = lambda t: torch.cat([t, t], 0)
    zg2 = two(z.clone().to(DEV)).requires_grad_(True)
    ret2 = gen._forward_perpix(None, two(vid.unsqueeze(0)), two(dep.unsqueeze(0)), two(rd.unsqueeze(0)), two(o.unsqueeze(0).to(DEV)),
                               zg2, gg.detach())
    assert 'ref' not in called and ret2[0].shape[0] == 2
    assert float((ret2[0][0] - ret2[0][1]).abs().max()) == 0.0
    (ret2[0] * two(G)).sum().backward()
    a, b = got['fc_1.weight'].grad.cpu().double(), 2.0 * Pc['render_net.fc_1.weight'].grad.double()
    assert float((a - b).norm() / b.norm()) <= 1e-2
    # options outside the fused path (here: a pre-set sky_avg under autograd) defer to the reference's own composition
    gen.sky_avg = torch.zeros(1, 1, 1, 1, 64, device=DEV)
    gen._forward_perpix(None, vid.unsqueeze(0), dep.unsqueeze(0), rd.unsqueeze(0), o.unsqueeze(0).to(DEV), zg, gg)
    assert called.get('ref')
