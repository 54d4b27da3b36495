"""Multi-GPU plumbing: one process per GPU, `torch.distributed` for the only collective the path has.

Barcode groups (single-sample runs) or samples (VCF runs) shard across ranks with no data-path
exchange: every rank plans the same schedule, emits its own contiguous molecule range, and writes
its own FASTQ chunk.  What ranks do need from each other is where their chunk sits in the whole
job — per-rank (bytes_R1, bytes_R2, n_pairs, first_read_id) — a tiny all-gather (NCCL over NVLink on
GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence


@dataclass
class RankCounts:
    r1_bytes: int
    r2_bytes: int
    n_pairs: int
    first_pair: int


@dataclass
class ShardLayout:
    """Where every rank's output sits in the concatenated (rank-order) files."""
    counts: List[RankCounts]
    r1_offset: List[int]
    r2_offset: List[int]
    total_pairs: int
    total_r1_bytes: int
    total_r2_bytes: int


def layout_from_counts(counts: Sequence[RankCounts]) -> ShardLayout:
    o1, o2, a1, a2 = [], [], 0, 0
    for c in counts:
        o1.append(a1)
        o2.append(a2)
        a1 += c.r1_bytes
        a2 += c.r2_bytes
    pairs = sum(c.n_pairs for c in counts)
    # shards are contiguous molecule ranges in rank order: read ids must chain
    nxt = counts[0].first_pair if counts else 0
    for c in counts:
        if c.first_pair != nxt:
            raise ValueError(f"shards do not chain: expected first read id {nxt}, rank reports {c.first_pair}")
        nxt += c.n_pairs
    return ShardLayout(list(counts), o1, o2, pairs, a1, a2)


def exchange_counts(mine: RankCounts, device=None) -> ShardLayout:
    """All-gather of the four per-rank integers. Works on any initialised process group; with a single
    process (no group) it is the identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return layout_from_counts([mine])
    world = dist.get_world_size()
    t = torch.tensor([mine.r1_bytes, mine.r2_bytes, mine.n_pairs, mine.first_pair], dtype=torch.int64, device=device)
    if t.is_cuda:   # one collective into one tensor, one device-to-host read
        out = torch.empty(world * 4, dtype=torch.int64, device=t.device)
        dist.all_gather_into_tensor(out, t)
        rows = out.cpu().view(world, 4).tolist()
    else:           # gloo
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        rows = [p.tolist() for p in parts]
    return layout_from_counts([RankCounts(*[int(x) for x in r]) for r in rows])


def samples_for_rank(n_samples: int, rank: int, world: int) -> List[int]:
    """VCF mode: the reference runs one task per sample (Main.jl:104); here samples go round-robin to GPUs."""
    return list(range(rank, n_samples, world))
