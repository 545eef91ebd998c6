"""GPU parity tests for the boundary ops (DDA, hash-grid encoder, positional encoding) and the
tcgen05 self test.  Everything is called through the C ABI (scenedreamer_b200.ops -> libsdb200.so)
and compared with (1) the CPU oracle and (2) the reference's own CUDA extension prebuilt into
oracle/_ref/ (when present)."""
import numpy as np
import pytest
import torch

import oracle
from scenedreamer_b200 import ops, synth

from _ref_ext import load as load_ref

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def bits(t):
    return t.detach().cpu().contiguous().view(torch.int32)


@pytest.fixture(scope='module')
def world():
    return synth.SyntheticVoxelWorld(size=256, seed=11)


def _frame(world, k, hw=(60, 100), pad=6, pattern=0):
    pose = synth.eval_camera_poses(world, maxstep=8, pattern=pattern)[k]
    return synth.frame_camera(world, pose, resolution_hw=hw, pad=pad)


@pytest.mark.parametrize('k,pattern', [(0, 0), (3, 0), (5, 4)])
def test_dda_bit_exact_vs_oracle(world, k, pattern):
    o, d, u, f, c, res = _frame(world, k, pattern=pattern)
    vox = world.voxel_t.to(DEV)
    vid, dep, rd = ops.ray_voxel_intersection_perspective(vox, o, d, u, f, c, res, 6)
    evid, edep, erd = oracle.ray_voxel_intersection_perspective(world.voxel_t, o, d, u, f, c, res, 6)
    assert vid.shape == (res[0], res[1], 6, 1) and dep.shape == (2, res[0], res[1], 6, 1) and rd.shape == (res[0], res[1], 1, 3)
    assert torch.equal(vid.cpu(), evid)                       # voxel ids / hit mask: bit exact
    assert torch.equal(bits(rd), bits(erd))                   # ray directions: bit exact
    assert torch.equal(torch.isnan(dep.cpu()), torch.isnan(edep))
    assert torch.equal(bits(torch.nan_to_num(dep, nan=-1.0)), bits(torch.nan_to_num(edep, nan=-1.0)))
    assert (evid[..., 0, 0] != 0).float().mean() > 0.2        # the frame really hits the scene


def test_dda_strided_volume_and_edge_cases(world):
    o, d, u, f, c, res = _frame(world, 2, hw=(17, 23), pad=0)
    base = torch.zeros(world.voxel_t.shape[0], world.voxel_t.shape[1], world.voxel_t.shape[2] * 2, dtype=torch.int32)
    base[:, :, ::2] = world.voxel_t
    strided = base[:, :, ::2]
    assert not strided.is_contiguous()
    vid, dep, rd = ops.ray_voxel_intersection_perspective(base.to(DEV)[:, :, ::2], o, d, u, f, c, res, 4)
    evid, edep, erd = oracle.ray_voxel_intersection_perspective(strided, o, d, u, f, c, res, 4)
    assert torch.equal(vid.cpu(), evid)
    assert torch.equal(bits(torch.nan_to_num(dep, nan=-1.0)), bits(torch.nan_to_num(edep, nan=-1.0)))
    # camera looking away from the volume: everything empty, NaN depths, id 0
    vid, dep, rd = ops.ray_voxel_intersection_perspective(world.voxel_t.to(DEV), [500., 128., 128.], [1., 0., 0.],
                                                           [0., 1., 0.], 30.0, [7.5, 9.5], [16, 20], 6)
    assert int(vid.abs().sum()) == 0 and bool(torch.isnan(dep).all())
    # axis-aligned ray (zero direction components -> HUGE_VALF axis times)
    vid, dep, rd = ops.ray_voxel_intersection_perspective(world.voxel_t.to(DEV), [200., 100.5, 77.5], [-1., 0., 0.],
                                                           [0., 1., 0.], 1.0, [0.0, 0.0], [1, 1], 6)
    evid, edep, _ = oracle.ray_voxel_intersection_perspective(world.voxel_t, [200., 100.5, 77.5], [-1., 0., 0.],
                                                              [0., 1., 0.], 1.0, [0.0, 0.0], [1, 1], 6)
    assert torch.equal(vid.cpu(), evid)
    assert torch.equal(bits(torch.nan_to_num(dep, nan=-1.0)), bits(torch.nan_to_num(edep, nan=-1.0)))
    with pytest.raises(RuntimeError):
        ops.ray_voxel_intersection_perspective(world.voxel_t, o, d, u, f, c, res, 4)      # CPU tensor
    with pytest.raises(RuntimeError):
        ops.ray_voxel_intersection_perspective(world.voxel_t.to(DEV).float(), o, d, u, f, c, res, 4)


def test_dda_bit_exact_vs_reference_cuda(world):
    ref = load_ref('ref_voxlib')
    if ref is None:
        pytest.skip('oracle/_ref/ref_voxlib not built')
    vox = world.voxel_t.to(DEV)
    for k, pattern in ((1, 0), (6, 4)):
        o, d, u, f, c, res = _frame(world, k, hw=(135, 240), pad=30, pattern=pattern)
        vid, dep, rd = ops.ray_voxel_intersection_perspective(vox, o, d, u, f, c, res, 6)
        rvid, rdep, rrd = ref.ray_voxel_intersection_perspective(vox, o, d, u, float(f), [float(c[0]), float(c[1])],
                                                                 [int(res[0]), int(res[1])], 6)
        assert torch.equal(vid, rvid)
        assert torch.equal(bits(rd), bits(rrd))
        assert torch.equal(bits(torch.nan_to_num(dep, nan=-1.0)), bits(torch.nan_to_num(rdep, nan=-1.0)))


def test_dda_row_bands_equal_the_rows_of_the_frame(world):
    """Single-frame sharding: the banded call returns exactly the rows the whole-frame call computes."""
    vox = world.voxel_t.to(DEV)
    o, d, u, f, c, res = _frame(world, 2, hw=(135, 240), pad=30, pattern=0)
    vid, dep, rd = ops.ray_voxel_intersection_perspective(vox, o, d, u, f, c, res, 6)
    for first, bh, stride in ((0, 16, 64), (48, 16, 64), (16, 16, 32), (0, res[0], res[0])):
        rows = [y for y0 in range(first, res[0], stride) for y in range(y0, min(res[0], y0 + bh))]
        bvid, bdep, brd = ops.ray_voxel_intersection_perspective(vox, o, d, u, f, c, [len(rows), res[1]], 6, band=(first, bh, stride))
        idx = torch.tensor(rows, device=DEV)
        assert torch.equal(bvid, vid[idx]) and torch.equal(bits(brd), bits(rd[idx]))
        assert torch.equal(bits(torch.nan_to_num(bdep, nan=-1.0)), bits(torch.nan_to_num(dep[:, idx], nan=-1.0)))
    with pytest.raises(RuntimeError):
        ops.ray_voxel_intersection_perspective(vox, o, d, u, f, c, [16, res[1]], 6, band=(0, 16, 8))       # overlapping bands


GE_CASES = [
    # D, C, L, base, log2T, desired, gridtype, B
    (5, 8, 16, 16, 19, 2048, 0, 4096),      # the SceneDreamer encoder
    (3, 2, 8, 4, 12, 64, 0, 3000),          # dense + hashed levels mixed
    (3, 4, 6, 4, 10, 48, 1, 1025),          # tiled
    (2, 1, 4, 8, 14, 64, 0, 777),
    (4, 2, 5, 4, 12, 40, 0, 513),
]


def device_level_scales(L, pls, base):
    """exp2f(level*S)*H - 1 evaluated by CUDA's exp2f (== what the reference kernel computes,
    gridencoder.cu:126); libm's exp2f can differ by 1 ulp, i.e. ~1e-4 cells at the finest level."""
    S = torch.tensor(float(np.float32(np.log2(pls))), device=DEV)
    lv = torch.arange(L, device=DEV, dtype=torch.float32)
    return (torch.exp2(lv * S) * float(base) - 1.0).cpu()


def _ge_setup(D, C, L, base, log2T, desired, seed=0, table_scale=0.1):
    offsets, pls = oracle.grid_offsets(D, L, None, base, log2T, desired)
    g = torch.Generator().manual_seed(seed)
    emb = (torch.rand(int(offsets[-1]), C, generator=g) * 2 - 1) * table_scale
    return offsets, pls, emb, g


@pytest.mark.parametrize('D,C,L,base,log2T,desired,gridtype,B', GE_CASES)
def test_grid_encode_forward_backward_vs_oracle(D, C, L, base, log2T, desired, gridtype, B):
    offsets, pls, emb, g = _ge_setup(D, C, L, base, log2T, desired)
    x = torch.rand(B, D, generator=g)
    x[::97] = 1.2        # out-of-range rows -> zero output, no gradient
    x[5] = 0.0
    x[6] = 1.0
    S = np.log2(pls)
    out = torch.empty(L, B, C, device=DEV)
    dy_dx = torch.empty(B, L * D * C, device=DEV)
    xe, ee, oe = x.to(DEV), emb.to(DEV), offsets.to(DEV)
    ops.grid_encode_forward(xe, ee, oe, out, B, D, C, L, S, base, True, dy_dx, gridtype, False)
    ls = device_level_scales(L, pls, base)
    eout, edy = oracle.grid_encode_forward(x, emb, offsets, pls, base, True, gridtype, False, level_scales=ls)
    np.testing.assert_allclose(out.cpu().numpy(), eout.numpy(), rtol=1e-5, atol=2e-7)
    np.testing.assert_allclose(dy_dx.cpu().numpy(), edy.numpy(), rtol=1e-4, atol=1e-4 * float(edy.abs().max()))
    assert float(out[:, ::97].abs().max()) == 0.0
    grad = torch.randn(L, B, C, generator=g)
    ge = torch.zeros_like(ee)
    gi = torch.zeros(B, D, device=DEV)
    ops.grid_encode_backward(grad.to(DEV), xe, ee, oe, ge, B, D, C, L, S, base, True, dy_dx, gi, gridtype, False)
    ege, egi = oracle.grid_encode_backward(grad, x, emb, offsets, pls, base, edy, gridtype, False, level_scales=ls)
    np.testing.assert_allclose(ge.cpu().numpy(), ege.numpy(), rtol=1e-4, atol=1e-5 * max(1.0, float(ege.abs().max())))
    np.testing.assert_allclose(gi.cpu().numpy(), egi.numpy(), rtol=1e-3, atol=1e-4 * max(1.0, float(egi.abs().max())))
    # size-independent properties: linearity of the gradient scatter and conservation of mass
    ge2 = torch.zeros_like(ee)
    ops.grid_encode_backward((2 * grad).to(DEV), xe, ee, oe, ge2, B, D, C, L, S, base, False, dy_dx, gi, gridtype, False)
    np.testing.assert_allclose(ge2.cpu().numpy(), 2 * ge.cpu().numpy(), rtol=1e-4, atol=1e-5)
    inb = ((x >= 0) & (x <= 1)).all(-1)
    np.testing.assert_allclose(float(ge.sum()), float(grad[:, inb].sum()), rtol=1e-3, atol=1e-2)


def test_grid_encode_unsupported_and_errors():
    offsets, pls, emb, g = _ge_setup(3, 2, 4, 4, 10, 32)
    x = torch.rand(16, 3, device=DEV)
    out = torch.empty(4, 16, 2, device=DEV)
    dd = torch.empty(1, device=DEV)
    with pytest.raises(RuntimeError):
        ops.grid_encode_forward(x.cpu(), emb.to(DEV), offsets.to(DEV), out, 16, 3, 2, 4, 1.0, 4, False, dd, 0, False)
    with pytest.raises(RuntimeError):
        ops.grid_encode_forward(x, emb.to(DEV), offsets.to(DEV), out, 16, 6, 2, 4, 1.0, 4, False, dd, 0, False)
    with pytest.raises(RuntimeError):
        ops.grid_encode_forward(x, emb.to(DEV), offsets.to(DEV), out, 16, 3, 3, 4, 1.0, 4, False, dd, 0, False)
    with pytest.raises(RuntimeError):
        ops.grid_encode_forward(x.t().contiguous().t(), emb.to(DEV), offsets.to(DEV), out, 16, 3, 2, 4, 1.0, 4, False, dd, 0, False)


def test_grid_encode_vs_reference_cuda():
    ref = load_ref('ref_gridencoder')
    if ref is None:
        pytest.skip('oracle/_ref/ref_gridencoder not built')
    for (D, C, L, base, log2T, desired, gridtype, B) in GE_CASES[:3]:
        offsets, pls, emb, g = _ge_setup(D, C, L, base, log2T, desired, seed=3)
        x = torch.rand(B, D, generator=g).to(DEV)
        S = np.log2(pls)
        ee, oe = emb.to(DEV), offsets.to(DEV)
        out, rout = torch.empty(L, B, C, device=DEV), torch.empty(L, B, C, device=DEV)
        dd, rdd = torch.empty(B, L * D * C, device=DEV), torch.empty(B, L * D * C, device=DEV)
        ops.grid_encode_forward(x, ee, oe, out, B, D, C, L, S, base, True, dd, gridtype, False)
        ref.grid_encode_forward(x, ee, oe, rout, B, D, C, L, S, base, True, rdd, gridtype, False)
        torch.cuda.synchronize()
        np.testing.assert_allclose(out.cpu().numpy(), rout.cpu().numpy(), rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(dd.cpu().numpy(), rdd.cpu().numpy(), rtol=1e-5, atol=1e-5 * float(rdd.abs().max()))
        grad = torch.randn(L, B, C, generator=g).to(DEV)
        ge, rge = torch.zeros_like(ee), torch.zeros_like(ee)
        gi, rgi = torch.zeros(B, D, device=DEV), torch.zeros(B, D, device=DEV)
        ops.grid_encode_backward(grad, x, ee, oe, ge, B, D, C, L, S, base, True, dd, gi, gridtype, False)
        ref.grid_encode_backward(grad, x, ee, oe, rge, B, D, C, L, S, base, True, rdd, rgi, gridtype, False)
        torch.cuda.synchronize()
        np.testing.assert_allclose(ge.cpu().numpy(), rge.cpu().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(gi.cpu().numpy(), rgi.cpu().numpy(), rtol=1e-4, atol=1e-4 * float(rgi.abs().max()))


def test_grid_encode_float16_table_vs_reference_cuda():
    """The reference's autocast path (grid.py:38-39: half table, half outputs / dy_dx / gradients, float32 coordinates):
    forward, dy_dx and the coordinate gradient are bit-identical to the reference's CUDA (c10::Half rounding after every
    operator, restated in gridenc.cu); the table gradient is accumulated with half2 atomics in both, so it is order-dependent."""
    ref = load_ref('ref_gridencoder')
    if ref is None:
        pytest.skip('oracle/_ref/ref_gridencoder not built')
    for (D, C, L, base, log2T, desired, gridtype, B) in ((5, 8, 16, 16, 19, 2048, 0, 4096), (3, 2, 8, 4, 12, 64, 0, 3000),
                                                         (3, 4, 6, 4, 10, 48, 1, 1025)):
        offsets, pls, emb, g = _ge_setup(D, C, L, base, log2T, desired, seed=5)
        x = torch.rand(B, D, generator=g).to(DEV)
        x[::97] = 1.5                                                     # out-of-range samples: zero rows
        S = np.log2(pls)
        ee, oe = emb.to(DEV).half(), offsets.to(DEV)
        mk = lambda *shape: torch.full(shape, float('nan'), device=DEV, dtype=torch.float16)
        out, rout, dd, rdd = mk(L, B, C), mk(L, B, C), mk(B, L * D * C), mk(B, L * D * C)
        ops.grid_encode_forward(x, ee, oe, out, B, D, C, L, S, base, True, dd, gridtype, False)
        ref.grid_encode_forward(x, ee, oe, rout, B, D, C, L, S, base, True, rdd, gridtype, False)
        torch.cuda.synchronize()
        assert torch.equal(out.view(torch.int16), rout.view(torch.int16))
        assert torch.equal(dd.view(torch.int16), rdd.view(torch.int16))
        assert float(out[:, ::97].abs().max()) == 0.0 and float(out.float().abs().max()) > 0
        grad = (torch.randn(L, B, C, generator=g) * 0.1).to(DEV).half()
        ge, rge = torch.zeros_like(ee), torch.zeros_like(ee)
        gi, rgi = torch.zeros(B, D, device=DEV, dtype=torch.float16), torch.zeros(B, D, device=DEV, dtype=torch.float16)
        ops.grid_encode_backward(grad, x, ee, oe, ge, B, D, C, L, S, base, True, dd, gi, gridtype, False)
        ref.grid_encode_backward(grad, x, ee, oe, rge, B, D, C, L, S, base, True, rdd, rgi, gridtype, False)
        torch.cuda.synchronize()
        assert torch.equal(gi.view(torch.int16), rgi.view(torch.int16))
        # float32 accumulation of the same addends as the referee of both half-atomic results
        g32, e32 = torch.zeros(ee.shape, device=DEV), ee.float()
        d32 = torch.empty(1, device=DEV)
        ops.grid_encode_backward(grad.float(), x, e32, oe, g32, B, D, C, L, S, base, False, d32, d32, gridtype, False)
        scale = float(g32.abs().max())
        err_ours, err_ref = float((ge.float() - g32).abs().max()), float((rge.float() - g32).abs().max())
        print('f16 table grad: max |ours - fp32| %.3e, |reference - fp32| %.3e (max |g| %.3e)' % (err_ours, err_ref, scale))
        assert err_ours <= max(2.0 * err_ref, 2e-3 * scale)
    with pytest.raises(RuntimeError):                                     # odd C stays float32 in the reference (grid.py:38)
        ops.grid_encode_forward(x, torch.zeros(64, 1, device=DEV, dtype=torch.float16), oe, mk(L, B, 1), B, D, 1, L, S, base,
                                False, mk(1), gridtype, False)
    with pytest.raises(RuntimeError):                                     # mixed dtypes
        ops.grid_encode_forward(x, ee, oe, torch.empty(L, B, C, device=DEV), B, D, C, L, S, base, False, mk(1), gridtype, False)


def test_positional_encoding_vs_oracle_and_reference():
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(37, 50, 1, 3, generator=g) * 2 - 1)
    y = ops.positional_encoding(x.to(DEV), 5, -1, True)
    assert y.shape == (37, 50, 1, 33)
    # the reference's own self-check tolerance (positional_encoding.py:63)
    np.testing.assert_allclose(y.cpu().numpy(), oracle.positional_encoding_pt(x, 5, -1, True).numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(y.cpu().numpy(), oracle.positional_encoding(x, 5, -1, True).numpy(), rtol=1e-5, atol=1e-5)
    x2 = torch.rand(6, 5, 7, generator=g) * 8
    y2 = ops.positional_encoding(x2.to(DEV), 4, 1, False)
    assert y2.shape == (6, 40, 7)
    np.testing.assert_allclose(y2.cpu().numpy(), oracle.positional_encoding_pt(x2, 4, 1, False).numpy(), rtol=1e-4, atol=1e-4)
    gy = torch.randn(y.shape, generator=g)
    gx = ops.positional_encoding_backward(gy.to(DEV), y, 5, -1, True)
    np.testing.assert_allclose(gx.cpu().numpy(), oracle.positional_encoding_backward(gy, y.cpu(), 5, -1, True).numpy(),
                               rtol=1e-4, atol=1e-4)
    ref = load_ref('ref_voxlib')
    if ref is not None:
        ry = ref.positional_encoding(x.to(DEV), 5, -1, True)
        np.testing.assert_allclose(y.cpu().numpy(), ry.cpu().numpy(), rtol=1e-6, atol=1e-6)
        rgx = ref.positional_encoding_backward(gy.to(DEV), ry, 5, -1, True)
        np.testing.assert_allclose(gx.cpu().numpy(), rgx.cpu().numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('N,K,bf16', [(256, 256, False), (256, 128, False), (64, 256, False), (256, 256, True), (32, 16, False)])
def test_tcgen05_selftest(N, K, bf16):
    g = torch.Generator().manual_seed(N + K)
    a = torch.randn(128, K, generator=g).to(DEV)
    b = torch.randn(N, K, generator=g).to(DEV)
    c = ops.tc_selftest(a, b, bf16=bf16, variant=0)
    torch.cuda.synchronize()
    lo = torch.bfloat16 if bf16 else torch.float16
    ref = a.to(lo).double() @ b.to(lo).double().t()
    err = float((c.double() - ref).abs().max())
    print('tcgen05 selftest N=%d K=%d bf16=%s max abs err %.3e (ref max %.2f)' % (N, K, bf16, err, float(ref.abs().max())))
    assert err < 1e-3 * max(1.0, float(ref.abs().max()))


def test_dda_origin_on_cell_faces_and_far_origin(world):
    """Exact-zero numerators (origin on integer coordinates -> signed zeros in IEEE division), origins far
    outside the grid and near-axis-parallel rays: still bit-exact (the kernel divides by a per-ray constant
    with the 3-FMA fast path of div.rn.f32 and falls back to the generic division outside its range)."""
    vox = world.voxel_t.to(DEV)
    X = world.voxel_t.shape[1]
    cases = [
        ([60.0, float(X), 128.0], [-0.2, -1.0, 0.01], [1.0, 0.0, 0.0]),       # origin exactly on the +x face, integer coords
        ([40.0, 64.0, 64.0], [-0.5, 0.7, 0.3], [1.0, 0.0, 0.0]),              # origin on a voxel corner inside the grid
        ([900.0, -700.5, 300.25], [-1.0, 1.0, -0.2], [1.0, 0.0, 0.0]),        # far outside
        ([55.5, 100.5, 100.5], [-1e-7, 1.0, 1e-9], [1.0, 0.0, 0.0]),          # nearly axis-parallel
    ]
    for o, d, u in cases:
        for f in (40.0, 400.0):
            args = (o, d, u, f, [31.5, 47.5], [64, 96], 6)
            vid, dep, rd = ops.ray_voxel_intersection_perspective(vox, *args)
            evid, edep, erd = oracle.ray_voxel_intersection_perspective(world.voxel_t, *args)
            assert torch.equal(vid.cpu(), evid)
            assert torch.equal(bits(rd), bits(erd))
            assert torch.equal(bits(torch.nan_to_num(dep, nan=-1.0)), bits(torch.nan_to_num(edep, nan=-1.0)))


@pytest.mark.parametrize('block_log2', [2, 3, 4, 6])
def test_dda_empty_space_flight_is_bit_identical(world, block_log2):
    """The exact flight across empty column blocks (sdb_build_height_bound + ..._ex) changes nothing but the step count:
    ids, depths and directions equal the plain cell-by-cell walk bit for bit, for every block size, for cameras above,
    inside and outside the volume, looking down, up and along the axes -- and a stale bound is rebuilt after an edit."""
    vox = world.voxel_t.to(DEV).clone()
    X = world.voxel_t.shape[1]
    old = ops.HEIGHT_BOUND_BLOCK_LOG2
    ops.HEIGHT_BOUND_BLOCK_LOG2 = block_log2
    try:
        cams = [_frame(world, k, hw=(90, 150), pad=10, pattern=pat) for k, pat in ((0, 0), (2, 0), (4, 4), (7, 4))]
        cams += [([200.0, 100.5, 77.5], [-1.0, 0.0, 0.0], [0.0, 1.0, 0.0], 60.0, [31.5, 47.5], [64, 96]),        # straight down
                 ([30.0, 64.2, 64.7], [1.0, 0.3, 0.2], [0.0, 1.0, 0.0], 40.0, [31.5, 47.5], [64, 96]),           # from inside, up and out
                 ([90.0, -50.5, 300.25], [-0.1, 1.0, -0.2], [1.0, 0.0, 0.0], 80.0, [31.5, 47.5], [64, 96]),      # from outside, grazing
                 ([60.0, float(X), 128.0], [-0.2, -1.0, 0.01], [1.0, 0.0, 0.0], 40.0, [31.5, 47.5], [64, 96])]   # origin on a face
        for cam in cams:
            a = ops.ray_voxel_intersection_perspective(vox, *cam[:6], 6)
            b = ops.ray_voxel_intersection_perspective(vox, *cam[:6], 6, empty_space_bound=False)
            assert torch.equal(a[0], b[0])
            assert torch.equal(bits(a[2]), bits(b[2]))
            assert torch.equal(bits(torch.nan_to_num(a[1], nan=-1.0)), bits(torch.nan_to_num(b[1], nan=-1.0)))
        # an in-place edit (a floating block high above the terrain) must invalidate the cached bound
        vox[vox.shape[0] - 3, 100:140, 100:140] = 7
        cam = cams[1]
        a = ops.ray_voxel_intersection_perspective(vox, *cam[:6], 6)
        b = ops.ray_voxel_intersection_perspective(vox, *cam[:6], 6, empty_space_bound=False)
        assert torch.equal(a[0], b[0]) and bool((a[0] == 7).any())
        assert torch.equal(bits(torch.nan_to_num(a[1], nan=-1.0)), bits(torch.nan_to_num(b[1], nan=-1.0)))
    finally:
        ops.HEIGHT_BOUND_BLOCK_LOG2 = old


@pytest.mark.parametrize('G', [64, 256])
def test_tcgen05_mn_major_operands_from_activation_tiles(G):
    """The activation tile layout of the fused kernels ([8-feature chunk][128 sample rows][16 B]) read as an MN-major
    tcgen05 operand with the samples as the reduction dimension (descriptor: LBO = 128 B along K, SBO = 2048 B along MN,
    instruction-descriptor major bits set): C = X^T Y, the shape of a weight-gradient GEMM."""
    import ctypes
    from scenedreamer_b200 import _lib
    g = torch.Generator().manual_seed(3)
    x = torch.randn(128, 128, generator=g).to(DEV)
    y = torch.randn(128, G, generator=g).to(DEV)
    c = torch.zeros(128, G, device=DEV)
    code = _lib.lib().sdb_tc_selftest_mn(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), ctypes.c_void_p(c.data_ptr()), G, 0,
                                         ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(code, 'sdb_tc_selftest_mn')
    torch.cuda.synchronize()
    ref = x.bfloat16().float().t() @ y.bfloat16().float()
    assert float((c - ref).abs().max()) <= 1e-3


def _sp_case(seed, ign_zero, strided_lut, C=40):
    g = torch.Generator().manual_seed(seed)
    X, Y, Z, M = 9, 12, 10, 300
    lut = torch.randint(0, M + (1 if ign_zero else 0), (X, Y, Z), generator=g, dtype=torch.int32)
    if strided_lut:
        lut = lut.permute(2, 0, 1).contiguous().permute(1, 2, 0)            # same values, non-contiguous strides
    feat = torch.randn(M, C, generator=g)
    wc = torch.rand(3, 50, 7, 3, generator=g) * torch.tensor([X + 2.0, Y + 2.0, Z + 2.0]) - 1.0      # some outside: clamped corners
    wc[0, 0, 0] = float('nan')
    wc[1, 3, 2, 1] = float('nan')
    wc[2, 5, 1] = torch.tensor([4.0, 7.0, 3.0])                            # exactly on a lattice point
    return lut, feat, wc


@pytest.mark.parametrize('ign_zero,strided,C', [(False, False, 40), (True, False, 40), (True, True, 40), (True, False, 64),
                                                 (False, True, 7), (True, False, 132)])
def test_sp_trilinear_worldcoord_vs_oracle_and_reference(ign_zero, strided, C):
    """voxlib.sp_trilinear_worldcoord[_backward] (surface parity): forward bit-exact against the CPU oracle and the
    reference's own CUDA extension, backward (atomics) to 1e-5.  C = 40 / 64 / 132: float4 lanes (16, 16, 32 per entry, the
    last with two chunks per lane); C = 7: the scalar kernel."""
    lut, feat, wc = _sp_case(3, ign_zero, strided, C)
    out = ops.sp_trilinear_worldcoord(feat.to(DEV), lut.to(DEV) if not strided else lut.to(DEV), wc.to(DEV), ign_zero, -1)
    ref = oracle.sp_trilinear_worldcoord(feat, lut, wc, ign_zero)
    assert out.shape == wc.shape[:-1] + (feat.shape[1],)
    assert torch.equal(out.cpu(), ref)
    assert float(out[0, 0, 0].abs().max()) == 0.0                           # NaN coordinate: nothing selected
    go = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    gf, = ops.sp_trilinear_worldcoord_backward(go.to(DEV), feat.to(DEV), lut.to(DEV), wc.to(DEV), ign_zero, False)
    gref = oracle.sp_trilinear_worldcoord_backward(go, feat, lut, wc, ign_zero)
    np.testing.assert_allclose(gf.cpu().numpy(), gref.numpy(), rtol=1e-5, atol=1e-5)
    # channel-first memory layout, channels still the last LOGICAL dim (reference :410-424)
    out_cf = ops.sp_trilinear_worldcoord(feat.to(DEV), lut.to(DEV), wc.to(DEV), ign_zero, -3)
    assert torch.equal(out_cf, out) and out_cf.stride(-1) == wc.shape[1] * wc.shape[2]
    rv = load_ref('ref_voxlib')
    if rv is not None:
        r_out = rv.sp_trilinear_worldcoord(feat.to(DEV), lut.to(DEV), wc.to(DEV), ign_zero, -1)
        assert torch.equal(r_out, out)
        r_cf = rv.sp_trilinear_worldcoord(feat.to(DEV), lut.to(DEV), wc.to(DEV), ign_zero, -3)
        assert torch.equal(r_cf, out_cf) and r_cf.stride() == out_cf.stride()
        r_g, = rv.sp_trilinear_worldcoord_backward(go.to(DEV), feat.to(DEV), lut.to(DEV), wc.to(DEV), ign_zero, False)
        np.testing.assert_allclose(gf.cpu().numpy(), r_g.cpu().numpy(), rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError):
        ops.sp_trilinear_worldcoord_backward(go.to(DEV), feat.to(DEV), lut.to(DEV), wc.to(DEV), ign_zero, True)


def test_tc_operand_with_shifted_start_row():
    """tcgen05 K-major operand read through a start address shifted by one 16-byte row inside a 130-row (haloed) buffer:
    what a 3x3 convolution tap of the RenderCNN kernel is.  Must equal the plain operand bit for bit."""
    g = torch.Generator().manual_seed(11)
    for N, K in ((256, 32), (64, 64)):
        a = torch.randn(128, K, generator=g).to(DEV)
        b = torch.randn(N, K, generator=g).to(DEV)
        c0 = ops.tc_selftest(a, b, variant=0)
        c2 = ops.tc_selftest(a, b, variant=2)
        torch.cuda.synchronize()
        assert torch.equal(c0, c2)
        assert float((c0 - a.half().float() @ b.half().float().t()).abs().max()) <= 1e-3


def test_pose_stats_match_torch(world):
    """f4: the camera sampler's rejection statistics (scenedreamer.py:127-142) from one device pass vs the reference's torch ops."""
    vox = world.voxel_t.to(DEV)
    for k in (0, 3):
        o, d, u, f, c, res = synth.frame_camera(world, synth.eval_camera_poses(world, maxstep=8, pattern=0)[k], (60, 90), 6)
        vid, dep, _ = ops.ray_voxel_intersection_perspective(vox, o, d, u, f, c, res, 6)
        st = ops.pose_stats(vid, dep).cpu()
        depth_map = dep[0, :, :, 0, :]
        avg = torch.mean(depth_map[~torch.isnan(depth_map)])
        cnt = torch.bincount(torch.flatten(vid[:, :, 0, 0]), weights=None, minlength=680).float() / (vid.size(0) * vid.size(1))
        ent = -torch.sum(cnt * torch.log(cnt + 1e-10))
        assert abs(float(st[0]) - float(avg)) <= 1e-5 * abs(float(avg)) and abs(float(st[1]) - float(ent)) <= 1e-5
    empty = torch.zeros(8, 8, 6, 1, dtype=torch.int32, device=DEV)
    nan = torch.full((2, 8, 8, 6, 1), float('nan'), device=DEV)
    st = ops.pose_stats(empty, nan).cpu()
    assert torch.isnan(st[0]) and abs(float(st[1])) <= 1e-6          # nothing hit: mean of nothing, one label -> zero entropy
