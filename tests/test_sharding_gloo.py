"""N>1 host logic on CPU: world_size-2 gloo run of the frame sharding + the single all-gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scenedreamer_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = sharding.frames_for_rank(n_frames, rank, world)
    # a "rendered frame" is a tensor filled with its global frame index
    local = torch.stack([torch.full((2, 3, 4), float(f)) for f in mine])
    allf = sharding.gather_frames(local)
    ok = allf.shape == (n_frames, 2, 3, 4) and all(float(allf[f].mean()) == float(f) for f in range(n_frames))
    q.put((rank, mine, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_and_all_gather_world2():
    world, n_frames = 2, 6
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3, 5]
    assert res[0][2] and res[1][2]


def _mean_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    H, W, C = 37, 5, 64
    g = torch.Generator().manual_seed(3)
    sky = torch.randn(H, W, C, generator=g)                                  # the same "frame" on every rank
    bands = [b for b in sharding.cyclic_bands(H, rank, world, band_tiles=1, tile_h=4) if b[1] > b[0]]
    parts = [(sky[y0:y1].reshape(-1, C).mean(0), (y1 - y0) * W) for (y0, y1) in bands]
    got = sharding.global_mean(parts)
    ref = sky.reshape(-1, C).mean(0)
    q.put((rank, float((got.reshape(C) - ref).abs().max())))
    dist.barrier()
    dist.destroy_process_group()


def test_global_mean_over_bands_world2():
    """The frame-global sky mean of the single-frame mode: band means on two ranks -> one small all-gather -> the mean of the frame."""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mean_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(err < 1e-6 for _, err in res)
    # one process, no process group: the plain weighted mean
    a, b = torch.ones(64), torch.zeros(64)
    assert torch.allclose(sharding.global_mean([(a, 30), (b, 10)]), torch.full((1, 64), 0.75))


def test_partitions_are_disjoint_and_complete():
    for world in (1, 2, 4, 8):
        frames = sorted(f for r in range(world) for f in sharding.frames_for_rank(40, r, world))
        assert frames == list(range(40))
        assert all(sharding.frame_owner(f, world) == f % world for f in range(40))


def test_band_spec_of_cyclic_bands():
    for world in (1, 2, 4, 8):
        for H in (570, 2190, 64):
            for r in range(world):
                bands = sharding.cyclic_bands(H, r, world)
                if not any(b[1] > b[0] for b in bands):
                    continue
                first, rows, stride, total = sharding.band_spec(bands)
                assert first == bands[0][0] and rows == (16 if world > 1 else H) or H < 16
                assert stride == (16 * world if world > 1 and len([b for b in bands if b[1] > b[0]]) > 1 else rows)
                # what the banded raycast computes: virtual row v -> frame row first + (v // rows) * stride + v % rows, v < total
                covered = [first + (v // rows) * stride + v % rows for v in range(total)]
                assert covered == [y for (y0, y1) in bands for y in range(y0, y1)]
    import pytest
    with pytest.raises(ValueError):
        sharding.band_spec([(0, 16), (20, 36), (48, 64)])                  # not equally spaced
    with pytest.raises(ValueError):
        sharding.band_spec([(0, 16), (32, 40), (64, 80)])                  # a short band in the middle
    with pytest.raises(ValueError):
        sharding.band_spec([(5, 5)])


def test_assemble_bands_restores_the_frame():
    """cyclic_bands -> per-rank concatenation (as bench.py --mode strong renders it) -> all-gather layout -> assemble_bands == frame."""
    for world in (1, 2, 4, 8):
        for H in (570, 64, 9):
            W = 7
            frame = torch.arange(2 * H * W, dtype=torch.float32).reshape(2, H, W)
            per = [sharding.cyclic_bands(H, r, world) for r in range(world)]
            n_slots = len(per[0])
            band_rows = 16 if world > 1 else H
            gathered = torch.zeros(world, 2, n_slots * band_rows, W)
            for r, bands in enumerate(per):
                rows = torch.cat([frame[:, y0:y1] for (y0, y1) in bands if y1 > y0], 1) if any(b[1] > b[0] for b in bands) else frame[:, :0]
                gathered[r, :, :rows.shape[1]] = rows                         # only a rank's LAST band can be short or missing
            out = sharding.assemble_bands(gathered, world, n_slots, band_rows)
            assert torch.equal(out[:, :H], frame)


def test_cyclic_bands_cover_the_frame_once():
    for world in (1, 2, 4, 8):
        for H in (570, 2190, 64, 9):
            per = [sharding.cyclic_bands(H, r, world) for r in range(world)]
            assert len({len(p) for p in per}) == 1                      # same number of slots on every rank (equal-size gather)
            bands = sorted(b for p in per for b in p if b[1] > b[0])
            assert bands[0][0] == 0 and bands[-1][1] == H
            assert all(bands[i][1] == bands[i + 1][0] for i in range(len(bands) - 1))
            assert all(b[0] % 8 == 0 and (b[1] - b[0] <= 16 or world == 1) for b in bands)      # one GPU: the frame is one band
