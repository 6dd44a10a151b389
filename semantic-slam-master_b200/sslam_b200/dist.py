"""Multi-GPU sharding of the hot path: one process per GPU, no collective between kernels.

Frames and frame pairs are independent units, so ranks work on disjoint shards
and the only communication is the final gather of match lists (``torch.distributed``: NCCL over
NVLink on GPUs, gloo in the CPU tests) — plus, for all-pairs (loop-closure) matching, one
all_gather of the normalised descriptor banks before the pair list is dealt out.
"""

import torch
import torch.distributed as dist


def shard_frames(num_frames, world, rank, halo=1):
    """Contiguous frame range [start, stop) of `rank` for consecutive-pair matching.  Each rank
    also extracts `halo` frame(s) of its right neighbour so that every pair (t, t+1) is local —
    one redundant extract per rank instead of any exchange.  Returns (start, stop, num_pairs)."""
    base = [(num_frames - 1) * r // world for r in range(world + 1)]      # pair boundaries
    p0, p1 = base[rank], base[rank + 1]
    if p1 <= p0:
        return p0, p0, 0
    return p0, min(num_frames, p1 + halo), p1 - p0


def shard_pairs(num_pairs, world, rank):
    """Independent pairs striped p mod world (c3-style).  Returns a LongTensor of pair ids."""
    return torch.arange(rank, num_pairs, world)


def all_pairs_index(num_keyframes):
    """Upper-triangular (a < b) pair list, int32 (P, 2), in row-major order."""
    iu = torch.triu_indices(num_keyframes, num_keyframes, offset=1)
    return iu.t().contiguous().to(torch.int32)


def deal_pairs_block_cyclic(pair_index, world, rank, block=16):
    """Deal an (a, b) pair list to ranks in tiles of `block` keyframes on each axis so that a rank
    re-uses resident descriptor tiles.  Returns the positions (LongTensor) owned by `rank`."""
    a = pair_index[:, 0].long() // block
    b = pair_index[:, 1].long() // block
    nb = int(b.max().item()) + 1 if pair_index.numel() else 1
    tile = a * nb + b
    uniq, inv = torch.unique(tile, sorted=True, return_inverse=True)
    owner = torch.arange(uniq.numel()) % world
    return torch.nonzero(owner[inv] == rank).squeeze(1)


def all_gather_bank(bank, group=None):
    """all_gather of equally-shaped per-rank descriptor banks (F_r, N, D) -> (world*F_r, N, D)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * bank.shape[0],) + tuple(bank.shape[1:]), dtype=bank.dtype,
                      device=bank.device)
    dist.all_gather_into_tensor(out, bank.contiguous(), group=group)
    return out


_gather_buffers = {}


def gather_match_lists(pairs, pair_scores, counts, dst=0, group=None):
    """Final collective: ONE gather of fixed-size records — per pair an int32 row
    [count | N (i, j) pairs, -1 padded | N fp32 score bits] (evaluation.pack_match_records) — into a
    preallocated buffer on `dst`.  Every rank must pass the same P, N.
    Returns (pairs (world*P,N,2), scores (world*P,N), counts (world*P)) on dst, None elsewhere."""
    from .evaluation import pack_match_records, unpack_match_records
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    rec = pack_match_records(pairs, pair_scores, counts)
    if rank != dst:
        dist.gather(rec, gather_list=None, dst=dst, group=group)
        return None
    key = (world, tuple(rec.shape), rec.device)
    buf = _gather_buffers.get(key)
    if buf is None:
        buf = _gather_buffers[key] = torch.empty((world,) + tuple(rec.shape), dtype=rec.dtype, device=rec.device)
    dist.gather(rec, gather_list=list(buf.unbind(0)), dst=dst, group=group)
    return unpack_match_records(buf.reshape(world * rec.shape[0], rec.shape[1]))


def compact_match_lists(pairs, pair_scores, counts):
    """Host-side view of gathered lists: list of ((K',2) int64 ndarray, (K',) fp32 ndarray)."""
    pairs, pair_scores, counts = pairs.cpu(), pair_scores.cpu(), counts.cpu()
    out = []
    for p in range(pairs.shape[0]):
        n = int(counts[p])
        out.append((pairs[p, :n].numpy().astype("int64"), pair_scores[p, :n].numpy()))
    return out
