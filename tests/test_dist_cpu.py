"""CPU: host-side sharding logic and the final gather of match lists, world_size 2 over gloo."""

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sslam_b200 import dist as sdist


def test_shard_frames_covers_every_pair_once():
    for T in (2, 3, 10, 599, 600, 601):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, e, n = sdist.shard_frames(T, world, r)
                assert n == max(0, e - s - 1) or n == 0
                seen += [(t, t + 1) for t in range(s, s + n)]
                if n:
                    assert e - s == n + 1          # exactly one halo frame
            assert seen == [(t, t + 1) for t in range(T - 1)]


def test_shard_pairs_and_all_pairs_index():
    idx = sdist.all_pairs_index(256)
    assert idx.shape == (32640, 2) and idx.dtype == torch.int32
    assert bool((idx[:, 0] < idx[:, 1]).all())
    owned = [sdist.deal_pairs_block_cyclic(idx, 8, r) for r in range(8)]
    allpos = torch.cat(owned).sort().values
    assert torch.equal(allpos, torch.arange(32640))
    sizes = [o.numel() for o in owned]
    assert max(sizes) - min(sizes) <= 2 * 16 * 16   # balanced to within a couple of tiles
    parts = [sdist.shard_pairs(64, 4, r) for r in range(4)]
    assert torch.equal(torch.cat(parts).sort().values, torch.arange(64))


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, N = 5, 16
        g = torch.Generator().manual_seed(100 + rank)
        counts = torch.randint(0, N + 1, (P,), generator=g, dtype=torch.int32)
        pairs = torch.full((P, N, 2), -1, dtype=torch.int32)
        scores = torch.zeros(P, N)
        for p in range(P):
            n = int(counts[p])
            pairs[p, :n, 0] = torch.arange(n, dtype=torch.int32)
            pairs[p, :n, 1] = torch.arange(n, dtype=torch.int32) + rank
            scores[p, :n] = rank + 0.5
        out = sdist.gather_match_lists(pairs, scores, counts, dst=0)
        bank = torch.full((2, 3, 4), float(rank))
        full = sdist.all_gather_bank(bank)
        assert full.shape == (2 * world, 3, 4)
        for r in range(world):
            assert bool((full[2 * r:2 * r + 2] == r).all())
        if rank == 0:
            gp, gs, gc = out
            assert gp.shape == (world * P, N, 2) and gc.shape == (world * P,)
            lists = sdist.compact_match_lists(gp, gs, gc)
            for r in range(world):
                for p in range(P):
                    m, s = lists[r * P + p]
                    assert m.dtype == np.int64 and m.shape[0] == int(gc[r * P + p])
                    if m.shape[0]:
                        assert (m[:, 1] - m[:, 0] == r).all() and (s == r + 0.5).all()
            open(os.path.join(tmp, "ok"), "w").write("1")
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_gather_match_lists_gloo_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
