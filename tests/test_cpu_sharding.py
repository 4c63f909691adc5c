"""CPU tests of the patch-sharded sliding-window driver's HOST logic (prediction.py:80-110 sharded by window):
window partition, batch sizing, the planes each rank needs, and - world_size 2 and 3, gloo - that the point-to-point
exchange of the ranks' fixed-point partial planes (exchange_partials + gather_planes, the code predict_device runs over
NCCL) reproduces the single-process accumulation bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from se_unet_airseg_b200.inference import (SlidingWindowPredictor, coverage_counts, exchange_partials, gather_planes,
                                           owner_ranges, shard_range, split_batches, window_starts)


def test_shard_range_partitions_the_window_list():
    for n in (1, 7, 37, 294):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_split_batches_is_even_and_bounded():
    assert split_batches(294, 7) == [7] * 42
    assert split_batches(37, 7) == [7, 6, 6, 6, 6, 6]
    assert split_batches(0, 7) == []
    for n in range(1, 60):
        s = split_batches(n, 7)
        assert sum(s) == n and max(s) <= 7 and len(set(s)) <= 2 and len(s) == -(-n // 7)


def test_shard_planes_cover_exactly_the_windows_of_the_rank():
    sw = SlidingWindowPredictor(model=None)
    shape = (512, 512, 400)
    sx, sy, sz = (window_starts(n) for n in shape)
    wins = [(a, b, c) for a in sx for b in sy for c in sz]
    assert len(wins) == 294
    for world in (2, 4, 8):
        covered = np.zeros(512, dtype=bool)
        for r in range(world):
            lo, hi = shard_range(len(wins), r, world)
            xa, xb = sw.shard_planes(shape, r, world)
            assert xa == min(w[0] for w in wins[lo:hi]) and xb == max(w[0] for w in wins[lo:hi]) + 128
            covered[xa:xb] = True
            assert xb - xa <= 512 // world + 192        # a rank copies roughly its share of the volume, not all of it
        assert covered.all()


def _accumulate(shape, wins, probs, acc_log2):
    acc = np.zeros(shape, dtype=np.int64)
    for (a, b, c), p in zip(wins, probs):
        acc[a:a + 16, b:b + 16, c:c + 16] += np.rint(p.astype(np.float32) * np.float32(2.0 ** acc_log2)).astype(np.int64)
    return acc


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shape, wins, probs = _case()
    lo, hi = shard_range(len(wins), rank, world)
    part = torch.from_numpy(_accumulate(shape, wins[lo:hi], probs[lo:hi], 26).astype(np.int32))
    slabs = []
    for r in range(world):
        a, b = shard_range(len(wins), r, world)
        slabs.append((min(w[0] for w in wins[a:b]), max(w[0] for w in wins[a:b]) + 16) if b > a else (0, 0))
    owners = owner_ranges(shape[0], world)
    # planes outside the rank's slab were never written by its windows: poison them to prove they are not used
    poison = torch.ones_like(part, dtype=torch.bool)
    poison[slabs[rank][0]:slabs[rank][1]] = False
    poison[owners[rank][0]:owners[rank][1]] = False
    part[poison] = 123456789
    exchange_partials(part, slabs, owners, rank, world)          # the exchange step of predict_device(_shard=...)
    x0, x1 = owners[rank]
    own = part[x0:x1].clone()
    part.fill_(-1)                                               # only the owned planes travel to rank 0
    part[x0:x1] = own
    gather_planes(part, owners, rank, world)
    if rank == 0:
        torch.save(part, out)
    dist.destroy_process_group()


def _case():
    shape = (40, 24, 28)
    sx, sy, sz = (window_starts(n, 16, 8) for n in shape)
    wins = [(a, b, c) for a in sx for b in sy for c in sz]
    rng = np.random.RandomState(3)
    probs = [rng.rand(16, 16, 16).astype(np.float32) for _ in wins]
    return shape, wins, probs


@pytest.mark.parametrize("world", [2, 3, 8])
def test_partial_plane_exchange_is_bit_exact(tmp_path, world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "acc.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    got = torch.load(out).numpy().astype(np.int64)
    shape, wins, probs = _case()
    want = _accumulate(shape, wins, probs, 26)
    assert np.array_equal(got, want)
    # the analytic count volume (product of per-axis coverages) equals the number of windows that touched each voxel
    cnt = np.zeros(shape, dtype=np.int64)
    for a, b, c in wins:
        cnt[a:a + 16, b:b + 16, c:c + 16] += 1
    cx, cy, cz = (coverage_counts(n, window_starts(n, 16, 8), 16) for n in shape)
    assert np.array_equal(cnt, cx[:, None, None] * cy[None, :, None] * cz[None, None, :])
    assert cnt.max() * 2 ** 26 < 2 ** 31


def test_owner_ranges_partition_the_first_axis():
    for n in (40, 400, 512):
        for world in (1, 2, 3, 4, 8):
            o = owner_ranges(n, world)
            assert o[0][0] == 0 and o[-1][1] == n and all(a[1] == b[0] for a, b in zip(o, o[1:]))
            assert max(b - a for a, b in o) - min(b - a for a, b in o) <= 1
