"""How much of the shaded work of the C2 bench frames is dead (diagnostics, GPU): per frame
  credited samples, rays cast, live 16x8 tiles, live rays, tile-steps executed (tile-level early termination),
  and what RAY-level compaction + termination would shade (sum over live rays of the steps until T < 1e-7).
Usage: python tools/ray_stats.py [--frames 6]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--frames', type=int, default=6)
    a = ap.parse_args()
    from scenedreamer_b200 import render, synth
    dev = torch.device('cuda', 0)
    world, poses, P, z, genc, lut = bench.build_workload(dev)
    fr = bench.FrameRenderer(world, P, z, genc, lut, dev, render.PRECISION_FP16X3, bench.SPP)
    S = bench.SPP
    for k in range(a.frames):
        cam = synth.frame_camera(world, poses[(k * 7) % 40], bench.OUT_HW, bench.PAD)
        o, d, u, f, c, res = cam
        vid, dep, rd = fr.ops.ray_voxel_intersection_perspective(fr.voxel, o, d, u, f, c, res, 6)
        vid, dep, rd = vid.unsqueeze(0), dep.unsqueeze(0), rd.unsqueeze(0)
        fr.r.early_stop = 0.0
        out = fr.r.forward(vid, dep, rd, o.unsqueeze(0), fr.z, fr.genc, num_samples=S, want_samples=True)
        w = out['weights'][0, ..., 0]                                  # [H,W,S]
        live = vid[0, :, :, 0, 0] != 0
        T_after = 1.0 - torch.cumsum(w.double(), -1)                   # transmittance after sample s (w = (1-e^-e) T_before)
        done = T_after < 1e-7
        first = torch.where(done.any(-1), done.float().argmax(-1) + 1, torch.full_like(done[..., 0], S, dtype=torch.int64).long())
        ray_steps = (first.clamp(max=S) * live).sum().item()
        fr.r.early_stop = None
        ws = fr.r.forward(vid, dep, rd, o.unsqueeze(0), fr.z, fr.genc, num_samples=S)['workspace'][:8].view(torch.int32).cpu()
        H, W = res
        tiles = ((H + 7) // 8) * ((W + 15) // 16)
        print('frame %2d: rays %d live %.3f | tiles %d live %d (%.3f) | tile-steps exec %d -> shaded %.2f M | live rays x S %.2f M | '
              'ray-level termination %.2f M | credited %.2f M'
              % (k, H * W, float(live.float().mean()), tiles, int(ws[0]), int(ws[0]) / tiles, int(ws[1]), int(ws[1]) * 128 / 1e6,
                 float(live.sum()) * S / 1e6, ray_steps / 1e6, bench.SAMPLES_PER_FRAME / 1e6))
        hist = torch.bincount(first[live].clamp(max=S), minlength=S + 1).cpu().numpy()
        print('          termination step histogram (live rays):', ' '.join(str(int(v)) for v in hist))


if __name__ == '__main__':
    main()
