"""CPU tests of the patch-sharded sliding-window driver's HOST logic (prediction.py:80-110 sharded by window):
window partition, batch sizing, the planes each rank needs, and - world_size 2, gloo - that summing the ranks' fixed-point
partial volumes with one integer reduce reproduces the single-process accumulation bit for bit."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from se_unet_airseg_b200.inference import (SlidingWindowPredictor, coverage_counts, shard_range, split_batches,
                                           window_starts)


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
    dist.reduce(part, dst=0, op=dist.ReduceOp.SUM)                 # the exchange step of predict_device(_shard=...)
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


def test_two_rank_integer_partial_volume_reduce_is_bit_exact(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "acc.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
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
