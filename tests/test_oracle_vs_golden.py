"""Pins the CPU oracle (oracle/) to outputs of the reference's own Python code
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import torch

import oracle


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_label_lut(golden_ops):
    g = golden_ops
    lut = T(g['mc2reduced_lut'])
    nlab, ign, dirt, water = [int(v) for v in g['label_meta']]
    assert (nlab, ign, dirt) == (12, 0, 3)
    out = lut.long()[T(g['mc2reduced_probe_in']).long()]
    out[out == ign] = dirt
    assert torch.equal(out.int(), T(g['mc2reduced_probe_out']))


def test_grid_offsets(golden_ops):
    off, pls = oracle.grid_offsets()
    assert torch.equal(off, T(golden_ops['ge5_offsets']))
    assert pls == float(golden_ops['ge5_per_level_scale'][0])
    assert off[1].item() == 1 << 19 and off[-1].item() == 16 << 19
    off3, pls3 = oracle.grid_offsets(input_dim=3, num_levels=8, base_resolution=4, log2_hashmap_size=12,
                                     desired_resolution=64)
    assert torch.equal(off3, T(golden_ops['ge3_offsets']))
    assert pls3 == float(golden_ops['ge3_per_level_scale'][0])


def test_sampling_deterministic(golden_ops):
    g = golden_ops
    d2 = T(g['samp_depth2'])
    for ns in (25, 5):
        rd, nd, idx = oracle.sample_depth_batched(d2, ns, deterministic=True, sample_depth=3)
        assert torch.equal(idx, T(g['samp_det%d_idx' % ns]))
        np.testing.assert_allclose(nd.numpy(), g['samp_det%d_new_dists' % ns], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(rd.numpy(), g['samp_det%d_rand_depth' % ns], rtol=1e-6, atol=1e-6,
                                   equal_nan=True)


def test_sampling_stratified(golden_ops):
    g = golden_ops
    rd, nd, idx = oracle.sample_depth_batched(T(g['samp_depth2']), 25, deterministic=False, sample_depth=3,
                                              uniforms=T(g['samp_uniforms']))
    assert torch.equal(idx, T(g['samp_rnd_idx']))
    np.testing.assert_allclose(nd.numpy(), g['samp_rnd_new_dists'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(rd.numpy(), g['samp_rnd_rand_depth'], rtol=1e-6, atol=1e-6, equal_nan=True)


def test_volume_rendering(golden_ops):
    g = golden_ops
    w = oracle.volum_rendering_relu(T(g['vr_sigma']), T(g['vr_dists']), dim=-2)
    np.testing.assert_allclose(w.numpy(), g['vr_weights'], rtol=1e-5, atol=1e-7)


def test_mlps(golden_ops):
    g = golden_ops
    for tag, stress in (('spec', False), ('stress', True)):
        P = oracle.make_params(seed=3, stress=stress, table_entries=8)
        z = oracle.style_mlp(T(g['mlp_%s_zin' % tag]), P)
        np.testing.assert_allclose(z.numpy(), g['mlp_%s_z' % tag], rtol=1e-5, atol=1e-6)
        x = T(g['mlp_%s_x' % tag])
        lab = T(g['mlp_%s_lab' % tag])
        s, c = oracle.render_mlp(x.reshape(2, -1, 128), z, lab.reshape(2, -1), P)
        scale = np.abs(g['mlp_%s_c' % tag]).max()
        np.testing.assert_allclose(s.reshape(2, 3, 4, 5, 1).numpy(), g['mlp_%s_sigma' % tag], rtol=1e-4,
                                   atol=1e-5 * max(1.0, np.abs(g['mlp_%s_sigma' % tag]).max()))
        np.testing.assert_allclose(c.reshape(2, 3, 4, 5, 64).numpy(), g['mlp_%s_c' % tag], rtol=1e-4,
                                   atol=1e-5 * max(1.0, scale))
        sky = oracle.sky_mlp(T(g['mlp_%s_pe' % tag]).reshape(2, -1, 33), z, P)
        np.testing.assert_allclose(sky.reshape(2, 3, 4, 1, 64).numpy(), g['mlp_%s_sky' % tag], rtol=1e-4,
                                   atol=1e-5)
    # the stress set must really exercise clamp / opacity (otherwise 1e-3 parity is vacuous)
    assert np.abs(g['mlp_stress_c']).max() > 1.0 and np.abs(g['mlp_stress_sigma']).max() > 10.0


def test_positional_encoding(golden_ops):
    g = golden_ops
    x = T(g['pe_in'])
    # the reference's own CUDA-vs-PyTorch self check uses rtol=atol=1e-5 (positional_encoding.py:63)
    np.testing.assert_allclose(oracle.positional_encoding(x, 5, -1, True).numpy(), g['pe_out_5_orig'],
                               rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oracle.positional_encoding(x, 4, 1, False).numpy(), g['pe_out_4_dim1'],
                               rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oracle.positional_encoding_pt(x, 5, -1, True).numpy(), g['pe_out_5_orig'],
                               rtol=1e-6, atol=1e-6)
    # backward against autograd of the pure-PyTorch statement
    xr = x.clone().requires_grad_(True)
    y = oracle.positional_encoding_pt(xr, 5, -1, True)
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(1))
    y.backward(gy)
    gx = oracle.positional_encoding_backward(gy, y.detach(), 5, -1, True)
    np.testing.assert_allclose(gx.numpy(), xr.grad.numpy(), rtol=1e-4, atol=1e-4)


def test_forward_perpix(golden_fpp, golden_ops):
    """oracle.forward_perpix == reference Generator._forward_perpix on the committed tiny frame."""
    g = golden_fpp
    lut = T(golden_ops['mc2reduced_lut'])
    offsets, pls = oracle.grid_offsets()
    vid, dep, rd = T(g['fpp_voxel_id']), T(g['fpp_depth2']), T(g['fpp_raydirs'])
    cam = g['fpp_cam']
    for tag, stress in (('spec', False), ('stress', True)):
        P = oracle.make_params(seed=9, stress=stress)
        r = oracle.forward_perpix(P, vid.unsqueeze(0), dep.unsqueeze(0), rd.unsqueeze(0),
                                  T(cam[0:3]).float().unsqueeze(0), T(g['fpp_%s_z' % tag]),
                                  T(g['fpp_%s_genc' % tag]), [int(v) for v in g['fpp_voxel_dims']], lut,
                                  offsets, pls, num_samples=24, deterministic=True)
        assert torch.equal(r['new_idx'], T(g['fpp_%s_new_idx' % tag]))
        np.testing.assert_allclose(r['rand_depth'].numpy(), g['fpp_%s_rand_depth' % tag], rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(r['new_dists'].numpy(), g['fpp_%s_new_dists' % tag], rtol=1e-5, atol=1e-7)
        tol = 2e-5 if not stress else 2e-4
        np.testing.assert_allclose(r['net_out_s'].numpy(), g['fpp_%s_net_out_s' % tag], rtol=1e-4,
                                   atol=tol * max(1.0, np.abs(g['fpp_%s_net_out_s' % tag]).max()))
        np.testing.assert_allclose(r['net_out_c'][:, 10:12].numpy(), g['fpp_%s_net_out_c' % tag], rtol=1e-4, atol=tol)
        np.testing.assert_allclose(r['weights'].numpy(), g['fpp_%s_weights' % tag], rtol=1e-4, atol=tol)
        np.testing.assert_allclose(r['sky_used'].numpy(), g['fpp_%s_skynet_out_c' % tag], rtol=1e-4, atol=tol)
        assert torch.equal(r['nosky_mask'], T(g['fpp_%s_nosky_mask' % tag]))
        np.testing.assert_allclose(r['net_out'].numpy(), g['fpp_%s_net_out' % tag], rtol=1e-4, atol=tol)
    # the stress frame must be non-trivial
    assert np.abs(g['fpp_stress_net_out']).max() > 0.5
    assert g['fpp_stress_weights'].max() > 0.5
