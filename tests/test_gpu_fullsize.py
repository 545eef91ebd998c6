"""Full-size (BASELINE C2: 570x990 rays, 24 spp, scene 1024) checks of the fused path: a window of the
frame against the oracle (1e-3), preblended vs raw-table agreement, finiteness and run-to-run determinism."""
import numpy as np
import pytest
import torch

import oracle
from scenedreamer_b200 import ops, render, synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
DEV = 'cuda:0'


def test_c2_frame_window_parity_and_properties(golden_ops):
    world = synth.SyntheticVoxelWorld(1024, 3407)
    pose = synth.eval_camera_poses(world, maxstep=40, pattern=0)[7]
    o, d, u, f, c, res = synth.frame_camera(world, pose, (540, 960), 30)
    vox = world.voxel_t.to(DEV)
    vid, dep, rd = ops.ray_voxel_intersection_perspective(vox, o, d, u, f, c, res, 6)
    P = oracle.make_params(seed=0, stress=True)
    Pd = {k: v.to(DEV) for k, v in P.items()}
    g = torch.Generator().manual_seed(8888)
    z = oracle.style_mlp(torch.randn(1, 128, generator=g), P)
    genc = torch.tanh(torch.randn(1, 2, generator=g))
    lut_raw = torch.from_numpy(golden_ops['mc2reduced_lut'])
    offsets, pls = oracle.grid_offsets()
    r = render.FusedPerPixelRenderer(Pd, world.voxel_t.shape, render.reduced_label_lut(lut_raw), pls)
    args = (vid.unsqueeze(0), dep.unsqueeze(0), rd.unsqueeze(0), o.unsqueeze(0), z.to(DEV), genc.to(DEV))
    out = r.forward(*args)
    torch.cuda.synchronize()
    net = out['net_out']
    assert net.shape == (1, res[0], res[1], 64) and bool(torch.isfinite(net).all())
    live = float((vid[..., 0, 0] != 0).float().mean())
    assert 0.3 < live < 0.95
    # determinism
    net2 = r.forward(*args)['net_out']
    assert torch.equal(net, net2)
    # pre-blended table vs the raw 5-D table (exact reference corner arithmetic): same result to fp32 rounding
    r_raw = render.FusedPerPixelRenderer(Pd, world.voxel_t.shape, render.reduced_label_lut(lut_raw), pls, preblend=False)
    net_raw = r_raw.forward(*args, sky=out['sky'], sky_avg=out['sky_avg'])['net_out']
    assert float((net_raw - net).abs().max()) <= 2e-4
    # a 48x64 window straddling the horizon vs the oracle (same frame-global sky mean)
    x0 = 400
    ys = int(torch.nonzero((vid[:, x0:x0 + 64, 0, 0] != 0).any(dim=1))[0])      # first row with a hit in these columns
    y0 = max(0, min(res[0] - 48, ys - 16))
    sl = (slice(y0, y0 + 48), slice(x0, x0 + 64))
    S = torch.tensor(float(np.float32(np.log2(pls))), device=DEV)
    ls = (torch.exp2(torch.arange(16, device=DEV, dtype=torch.float32) * S) * 16.0 - 1.0).cpu()
    ref = oracle.forward_perpix(P, vid[sl].unsqueeze(0).cpu(), dep[:, sl[0], sl[1]].unsqueeze(0).cpu(),
                                rd[sl].unsqueeze(0).cpu(), o.unsqueeze(0), z, genc, list(world.voxel_t.shape), lut_raw,
                                offsets, pls, sky_avg=out['sky_avg'].cpu().reshape(1, 1, 1, 1, 64), level_scales=ls)
    err = float((net[0][sl].cpu() - ref['net_out'][0]).abs().max())
    derr = float((out['depth'][0][sl].cpu() - ref['depth_map'][0].squeeze(-1)).abs().max())
    print('C2 window (%d:%d, %d:%d) max err net_out %.3e depth %.3e, live fraction %.2f' % (y0, y0 + 48, x0, x0 + 64, err, derr, live))
    # depth = sum_s w*t with t of several hundred voxels in this scene: ulp(512) = 6e-5, so 1e-3 ABSOLUTE is at
    # the fp32 rounding level (the reference's own fp32 evaluation is 1.2e-3 away from a float64 referee on this frame:
    # tests/test_gpu_generator.py); the bar is 1e-3 absolute or 1e-5 relative, whichever is larger
    dmax = float(ref['depth_map'].abs().max())
    assert err <= 1e-3 and derr <= max(1e-3, 1e-5 * dmax), (err, derr, dmax)
    wl = float((vid[sl][..., 0, 0] != 0).float().mean())
    assert 0.2 < wl < 1.0, wl                     # the window really mixes sky and geometry
    # full-frame DDA vs oracle, bit exact
    evid, edep, _ = oracle.ray_voxel_intersection_perspective(world.voxel_t, o, d, u, f, c, res, 6)
    assert torch.equal(vid.cpu(), evid)
    assert torch.equal(torch.nan_to_num(dep, nan=-1.0).cpu().view(torch.int32), torch.nan_to_num(edep, nan=-1.0).view(torch.int32))


def test_c4_frame_window_parity_and_properties(golden_ops):
    """BASELINE C4 (HBM-bound stress): one 2160x3840 frame, 40 samples/ray, scene_size 2048 (2190x3870 rays cast and
    shaded, 339 M samples): finiteness, determinism of a re-run, and a 48x64 window (DDA bit-exact, net_out / depth
    1e-3) against the oracle with the frame-global sky mean of the full frame."""
    world = synth.SyntheticVoxelWorld(2048, 3407)
    pose = synth.eval_camera_poses(world, maxstep=40, pattern=0)[3]
    o, d, u, f, c, res = synth.frame_camera(world, pose, (2160, 3840), 30)
    assert tuple(res) == (2190, 3870)
    vox = world.voxel_t.to(DEV)
    vid, dep, rd = ops.ray_voxel_intersection_perspective(vox, o, d, u, f, c, res, 6)
    P = oracle.make_params(seed=0, stress=True)
    Pd = {k: v.to(DEV) for k, v in P.items()}
    g = torch.Generator().manual_seed(8888)
    z = oracle.style_mlp(torch.randn(1, 128, generator=g), P)
    genc = torch.tanh(torch.randn(1, 2, generator=g))
    lut_raw = torch.from_numpy(golden_ops['mc2reduced_lut'])
    offsets, pls = oracle.grid_offsets()
    r = render.FusedPerPixelRenderer(Pd, world.voxel_t.shape, render.reduced_label_lut(lut_raw), pls)
    args = (vid.unsqueeze(0), dep.unsqueeze(0), rd.unsqueeze(0), o.unsqueeze(0), z.to(DEV), genc.to(DEV))
    out = r.forward(*args, num_samples=40)
    torch.cuda.synchronize()
    net = out['net_out']
    assert net.shape == (1, res[0], res[1], 64) and bool(torch.isfinite(net).all())
    assert bool(torch.isfinite(out['depth']).all()) and float(out['total_weight'].max()) <= 1.0 + 1e-5
    net2 = r.forward(*args, num_samples=40, sky=out['sky'], sky_avg=out['sky_avg'])['net_out']
    assert torch.equal(net, net2)
    x0 = 1900
    hit_rows = torch.nonzero((vid[:, x0:x0 + 64, 0, 0] != 0).any(dim=1))
    y0 = max(0, min(res[0] - 48, int(hit_rows[0]) - 16))
    sl = (slice(y0, y0 + 48), slice(x0, x0 + 64))
    # DDA of the window on the CPU: same camera with the principal point shifted to the window origin
    evid, edep, erd = oracle.ray_voxel_intersection_perspective(world.voxel_t, o, d, u, f, [c[0] - y0, c[1] - x0], [48, 64], 6)
    assert torch.equal(vid[sl].cpu(), evid)
    assert torch.equal(torch.nan_to_num(dep[:, sl[0], sl[1]], nan=-1.0).cpu().view(torch.int32),
                       torch.nan_to_num(edep, nan=-1.0).view(torch.int32))
    assert torch.equal(rd[sl].cpu(), erd)
    S = torch.tensor(float(np.float32(np.log2(pls))), device=DEV)
    ls = (torch.exp2(torch.arange(16, device=DEV, dtype=torch.float32) * S) * 16.0 - 1.0).cpu()
    ref = oracle.forward_perpix(P, evid.unsqueeze(0), edep.unsqueeze(0), erd.unsqueeze(0), o.unsqueeze(0), z, genc,
                                list(world.voxel_t.shape), lut_raw, offsets, pls, num_samples=40,
                                sky_avg=out['sky_avg'].cpu().reshape(1, 1, 1, 1, 64), level_scales=ls)
    err = float((net[0][sl].cpu() - ref['net_out'][0]).abs().max())
    derr = float((out['depth'][0][sl].cpu() - ref['depth_map'][0].squeeze(-1)).abs().max())
    dmax = float(ref['depth_map'].abs().max())
    wl = float((evid[..., 0, 0] != 0).float().mean())
    print('C4 window (%d:%d, %d:%d) max err net_out %.3e depth %.3e (max depth %.1f), window live fraction %.2f'
          % (y0, y0 + 48, x0, x0 + 64, err, derr, dmax, wl))
    assert err <= 1e-3 and derr <= max(1e-3, 1e-5 * dmax), (err, derr, dmax)
    assert 0.1 < wl <= 1.0, wl
