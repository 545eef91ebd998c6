"""Host side of the on-device scene builder (libsdb200: sdb_world_build / sdb_world_truncate).

Mirrors PCGVoxelGenerator.next_world (imaginaire/model_utils/pcg_gen.py:83-174): bird's-eye-view maps (height, semantic,
tree) + voxel tree models -> the voxel volume `voxel_t[height, x, z]`, the height map used by the camera controllers, the
world-to-local offset and the two conditioning maps.  What stays on the host is O(X*Z) bookkeeping that must consume the
host RNG exactly like the reference (the quantisation of the height map in numpy, the list of tree instances with
`random.choice` per accepted tree); the O(256*X*Z) volume only ever exists in HBM.
"""
import ctypes
import random

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib

SAMPLE_HEIGHT = 256
PAD_NUM = 16
BOUNDARY = 50
BIOME2MC = [28, 9, 8, 1, 9, 8, 9, 8, 30, 26]                                  # pcg_gen.py:118
BIOME_TREES = [[], [5], [1, 7], [], [1, 2], [1, 2, 3], [4], [0, 3], [5, 6, 7], []]   # pcg_gen.py:105-116, in dict order


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def tree_instances(height_q, tree_map, tree_models, rng=random):
    """The reference's double loop over biomes and tree cells (pcg_gen.py:132-146) without the pasting: returns int32
    [n, 4] = (h + 16, x, z, model id) in iteration order; `rng.choice` is called exactly when the reference calls it."""
    X, Z = height_q.shape
    h16 = height_q.astype(np.int64) + PAD_NUM
    inst = []
    for biome_id in range(len(BIOME2MC)):
        selected = BIOME_TREES[biome_id]
        if len(selected) == 0:
            continue
        xs, zs = np.nonzero(tree_map == biome_id)                              # row-major, like the boolean mask indexing
        for x, z in zip(xs.tolist(), zs.tolist()):
            h = int(h16[x, z])
            if x < BOUNDARY or x > X - BOUNDARY or z < BOUNDARY or z > Z - BOUNDARY or h > SAMPLE_HEIGHT - BOUNDARY:
                continue
            inst.append((h, x, z, rng.choice(selected)))
    return np.asarray(inst, dtype=np.int32).reshape(-1, 4)


def build_world(height_map, semantic_map, tree_map, tree_models, device, rng=random):
    """height_map float [X, Z] (values < 0 are water), semantic_map uint8 [X, Z] in 0..9, tree_map uint8 [X, Z] (255 = none),
    tree_models: sequence of int32 [dh, dx, dz] voxel models (ckpt['assets']).
    -> dict(voxel_t int32 [sky-gnd, X, Z] on `device`, heightmap int64 [X, Z] (CPU, like the reference), gnd_level,
            current_height_map [1,1,X,Z], current_semantic_map [1,C,X,Z] on `device`, total_size)."""
    L = _lib.lib()
    dev = torch.device(device)
    hm = np.array(height_map, copy=True)
    hm[hm < 0] = 0                                                                                   # pcg_gen.py:94
    hq = ((hm - hm.min()) / (1 - hm.min()) * (SAMPLE_HEIGHT - 1)).astype(np.int16)                   # :95
    X, Z = hq.shape
    sem = np.asarray(semantic_map)
    trees = np.asarray(tree_map)
    inst = tree_instances(hq, trees, tree_models, rng)
    models = [np.ascontiguousarray(np.asarray(m.cpu() if torch.is_tensor(m) else m, dtype=np.int32)) for m in tree_models]
    mdim = np.asarray([m.shape for m in models], dtype=np.int32).reshape(-1, 3)
    moff = np.cumsum([0] + [m.size for m in models[:-1]]).astype(np.int64) if models else np.zeros(0, np.int64)
    flat = np.concatenate([m.reshape(-1) for m in models]) if models else np.zeros(1, np.int32)
    label = np.asarray(BIOME2MC, dtype=np.int32)[sem.astype(np.int64)]
    with torch.cuda.device(dev):
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        # C-contiguous uploads: the maps may arrive Fortran-ordered (np.load keeps the order of the saved array), and a torch
        # tensor made from such an array keeps transposed strides all the way to the device
        d_hq = torch.from_numpy(np.ascontiguousarray(hq, dtype=np.int32)).to(dev)
        d_label = torch.from_numpy(np.ascontiguousarray(label, dtype=np.int32)).to(dev)
        d_inst = torch.from_numpy(np.ascontiguousarray(inst)).to(dev) if len(inst) else None
        d_models, d_mdim, d_moff = torch.from_numpy(flat).to(dev), torch.from_numpy(mdim).to(dev), torch.from_numpy(moff).to(dev)
        world = torch.empty(SAMPLE_HEIGHT, X, Z, dtype=torch.int32, device=dev)                       # scratch: 1 GB at 1024^2, 4.3 GB at 2048^2
        heightmap = torch.empty(X, Z, dtype=torch.int64, device=dev)
        minmax = torch.empty(2, dtype=torch.int32, device=dev)
        _lib.check(L.sdb_world_build(_ptr(d_hq), _ptr(d_label), X, Z, SAMPLE_HEIGHT, _ptr(d_inst), int(len(inst)), _ptr(d_models),
                                     _ptr(d_mdim), _ptr(d_moff), _ptr(world), _ptr(heightmap), _ptr(minmax), st), 'sdb_world_build')
        gnd, top = [int(v) for v in minmax.cpu()]                                                     # the output shape is data-dependent
        sky = top + 1
        voxel_t = torch.empty(sky - gnd, X, Z, dtype=torch.int32, device=dev)
        _lib.check(L.sdb_world_truncate(_ptr(world), X, Z, gnd, sky, _ptr(voxel_t), st), 'sdb_world_truncate')
        del world
        # O(X*Z) host arithmetic, uploaded: torch's CPU int64 / int division is what the reference's value is bit for bit
        h16 = torch.from_numpy(np.ascontiguousarray(hq, dtype=np.int64)) + PAD_NUM
        current_height_map = (h16 / (SAMPLE_HEIGHT - 1))[None, None].to(dev)                          # :167
        org_sem = torch.from_numpy(np.ascontiguousarray(sem)).to(dev)
        org_sem[torch.from_numpy(np.ascontiguousarray(trees != 255)).to(dev)] = 10                                          # :100-101
        current_semantic_map = F.one_hot(org_sem.to(torch.int64)).to(torch.float).permute(2, 0, 1)[None]   # :168
    return dict(voxel_t=voxel_t, heightmap=heightmap.cpu(), gnd_level=gnd, sky_level=sky, current_height_map=current_height_map,
                current_semantic_map=current_semantic_map, total_size=(X, Z))


def fused_next_world(self, device, world_dir, pcg_asset):
    """Replacement body of PCGVoxelGenerator.next_world (same arguments, same attributes set)."""
    import os
    import cv2
    if torch.device(device).type != 'cuda':
        return type(self)._sdb200_reference_next_world(self, device, world_dir, pcg_asset)
    height_map = np.load(os.path.join(world_dir, 'heightmap.npy'))                                    # :87-92
    semantic_map = cv2.imread(os.path.join(world_dir, 'semanticmap.png'), 0)
    tree_map = cv2.imread(os.path.join(world_dir, 'treemap.png'), 0)
    w = build_world(height_map, semantic_map, tree_map, pcg_asset['assets'], device)
    self.total_size = w['total_size']
    self.trans_mat = torch.eye(4)
    self.current_height_map = w['current_height_map']
    self.current_semantic_map = w['current_semantic_map']
    self.heightmap = w['heightmap']
    self.voxel_t = w['voxel_t']
    self.trans_mat[0, 3] += w['gnd_level']


def install(pcg_cls):
    if '_sdb200_reference_next_world' not in pcg_cls.__dict__ and 'next_world' in pcg_cls.__dict__:
        pcg_cls._sdb200_reference_next_world = pcg_cls.next_world
        pcg_cls.next_world = fused_next_world
    return pcg_cls
