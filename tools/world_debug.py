"""Diagnostics (GPU): on-device world builder vs a numpy emulation of the same algorithm, stage by stage."""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
from scenedreamer_b200 import synth, worldgen  # noqa: E402


def main():
    size = 192
    h, sem, tree = synth.make_bev(size, seed=11)
    tree[::5, ::7] = np.where(sem[::5, ::7] != 9, sem[::5, ::7], 255)
    models = synth.make_tree_models()
    random.seed(7)
    w = worldgen.build_world(h, sem, tree, models, 'cuda:0')
    torch.cuda.synchronize()
    hm = h.copy()
    hm[hm < 0] = 0
    hq = ((hm - hm.min()) / (1 - hm.min()) * 255).astype(np.int16)
    random.seed(7)
    inst = worldgen.tree_instances(hq, tree, models)
    X = Z = size
    world = np.zeros((256, X, Z), np.int64)
    lab = np.asarray(worldgen.BIOME2MC)[sem.astype(np.int64)]
    xi, zi = np.meshgrid(np.arange(X), np.arange(Z), indexing='ij')
    for k in range(17):
        world[np.clip(hq.astype(np.int64) + k, 0, 255), xi, zi] = lab
    base = world.copy()
    for t, (hh, x, z, m) in enumerate(inst.tolist()):
        for (a, b, c), v in np.ndenumerate(models[m]):
            if v == 0:
                continue
            y, xx, zz = hh + a, x + b, z + c
            if y >= 256 or xx >= X or zz >= Z:
                continue
            key = ((t + 1) << 10) | int(v)
            old = world[y, xx, zz]
            if old == 0 or (old >= 1024 and old > key):
                world[y, xx, zz] = key
    dec = np.where(world >= 1024, world & 1023, world)
    nz = dec != 0
    top = 255 - np.argmax(nz[::-1], axis=0)
    top[~nz.any(0)] = 0
    gnd, sky = int(top.min()), int(top.max()) + 1
    print('emulation gnd %d sky %d | device gnd %d sky %d shape %s' % (gnd, sky, w['gnd_level'], w['sky_level'], tuple(w['voxel_t'].shape)))
    hmap = w['heightmap'].numpy()
    print('heightmap mismatches', int((hmap != top).sum()), 'of', top.size)
    if w['gnd_level'] == gnd and w['sky_level'] == sky:
        v = w['voxel_t'].cpu().numpy()
        bad = np.argwhere(v != dec[gnd:sky])
        print('voxel mismatches', len(bad), 'of', v.size)
        for (y, x, z) in bad[:10]:
            print('   at', (y + gnd, x, z), 'device', v[y, x, z], 'emulation', dec[y + gnd, x, z], 'base', base[y + gnd, x, z], 'key', world[y + gnd, x, z])
    else:
        i = np.argwhere(hmap != top)[:5]
        for (x, z) in i:
            print('   column', (x, z), 'device top', hmap[x, z], 'emulation', top[x, z], 'hq', hq[x, z])


if __name__ == '__main__':
    main()
