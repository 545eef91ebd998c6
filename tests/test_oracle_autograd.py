"""CPU tests of the oracle's differentiable restatement (test infrastructure for the training parity tests):
the hash-grid autograd.Function built on oracle.c is checked against finite differences."""
import numpy as np
import torch

import oracle
from oracle import ref_ops


def test_grid_encode_autograd_matches_finite_differences():
    torch.manual_seed(0)
    offsets, pls = oracle.grid_offsets(num_levels=4, log2_hashmap_size=12, desired_resolution=64)
    n_emb = int(offsets[-1])
    emb = ((torch.rand(n_emb, 8) * 2 - 1) * 0.1).requires_grad_(True)
    x = (torch.rand(64, 5) * 0.9 + 0.05).requires_grad_(True)
    G = torch.randn(64, 4 * 8)
    y = ref_ops._GridEncodeFn.apply(x, emb, offsets, pls, 16, None)
    (y * G).sum().backward()
    # embeddings: the encode is linear in the table -> directional derivative is exact
    d = torch.randn_like(emb)
    with torch.no_grad():
        y1 = ref_ops._GridEncodeFn.apply(x.detach(), emb.detach() + d, offsets, pls, 16, None)
        lin = float(((y1 - y.detach()) * G).sum())
    assert abs(lin - float((emb.grad * d).sum())) <= 1e-3 * max(1.0, abs(lin))
    # inputs: piecewise linear, central differences with a step far below the finest cell (1/64)
    eps = 1e-4
    for dim in (0, 3, 4):
        dx = torch.zeros_like(x)
        dx[:, dim] = eps
        with torch.no_grad():
            yp = ref_ops._GridEncodeFn.apply(x.detach() + dx, emb.detach(), offsets, pls, 16, None)
            ym = ref_ops._GridEncodeFn.apply(x.detach() - dx, emb.detach(), offsets, pls, 16, None)
        fd = ((yp - ym) * G).sum(1) / (2 * eps)
        ok = (fd - x.grad[:, dim]).abs() <= 2e-2 * (1.0 + fd.abs())
        assert float(ok.float().mean()) > 0.9, dim      # rows whose +-eps straddles a cell boundary may differ


def test_forward_perpix_autograd_equals_forward_perpix(golden_ops):
    """Same numbers as the non-differentiable oracle path (which the golden fixtures pin)."""
    from scenedreamer_b200 import synth
    world = synth.SyntheticVoxelWorld(size=64, seed=3)
    pose = synth.eval_camera_poses(world, maxstep=8, pattern=0)[1]
    o, d, u, f, c, res = synth.frame_camera(world, pose, resolution_hw=(10, 14), pad=2)
    vid, dep, rd = oracle.ray_voxel_intersection_perspective(world.voxel_t, o, d, u, f, c, res, 6)
    vid, dep, rd = vid.unsqueeze(0), dep.unsqueeze(0), rd.unsqueeze(0)
    P = oracle.make_params(seed=2, stress=True, table_entries=16 * (1 << 19))
    g = torch.Generator().manual_seed(4)
    z = oracle.style_mlp(torch.randn(1, 128, generator=g), P)
    genc = torch.tanh(torch.randn(1, 2, generator=g))
    offsets, pls = oracle.grid_offsets()
    lut = torch.from_numpy(golden_ops['mc2reduced_lut'])
    a = oracle.forward_perpix(P, vid, dep, rd, o.unsqueeze(0), z, genc, list(world.voxel_t.shape), lut, offsets, pls,
                              num_samples=6)['net_out']
    Pl = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    b = oracle.forward_perpix_autograd(Pl, vid, dep, rd, o.unsqueeze(0), z, genc.clone().requires_grad_(True),
                                       list(world.voxel_t.shape), lut, offsets, pls, num_samples=6)
    assert float((a - b.detach()).abs().max()) <= 1e-5
    b.sum().backward()
    assert float(Pl['hash_encoder.embeddings'].grad.abs().max()) > 0
