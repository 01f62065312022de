"""Multi-GPU sharding of a batch of independent channels/files.

The path shards trivially: a channel (GC-ADPCM / ADX) or a stream (HCA) never exchanges data with another one during
compute, so there is NO data-path collective.  One process per GPU (torch.distributed, NCCL on GPUs / gloo in the CPU
tests); this module only decides who encodes what and brings the variable-size bitstreams back to one rank:

  partition(lengths, world)        greedy longest-first bin packing by sample count (lengths vary 10x in config C5)
  gather_bitstreams(local, ...)    one gather of the encoded bytes to `dst` (padded to the longest shard), restoring
                                   the caller's original order

The reference's counterpart is `Parallel.ForEach(files)` in src/VGAudio.Cli/Batch.cs:24-25 (a thread pool on one host).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np


def partition(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Assign item indices to `world` ranks so that the per-rank sample totals are balanced.
    Longest-first greedy: sort by length descending, always give the next item to the lightest rank; inside a rank the
    indices are then sorted by length so that warps of similar length sit together (they finish together)."""
    if world <= 0:
        raise ValueError("world must be positive")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += int(lengths[i])
    for r in range(world):
        shards[r].sort(key=lambda i: (-int(lengths[i]), i))
    return shards


def imbalance(lengths: Sequence[int], shards: List[List[int]]) -> float:
    """max rank load / mean rank load (1.0 = perfect)."""
    loads = [sum(int(lengths[i]) for i in s) for s in shards]
    mean = sum(loads) / max(len(loads), 1)
    return max(loads) / mean if mean > 0 else 1.0


def gather_bitstreams(local_items: Dict[int, np.ndarray], n_items: int, dst: int = 0, group=None,
                      device: Optional[str] = None) -> Optional[List[np.ndarray]]:
    """Gather {global index -> uint8 bitstream} from every rank to `dst`; returns the list in global order on `dst`,
    None elsewhere.  One collective for the payload (all shards padded to the longest) plus one tiny one for the
    index/length tables.  Works with the NCCL backend (tensors on the current CUDA device) and gloo (CPU)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if device is None:
        device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    idx = sorted(local_items)
    lens = [int(local_items[i].size) for i in idx]
    # tables: every rank learns every shard's (count, bytes)
    mine = torch.tensor([len(idx), sum(lens)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, mine, group=group)
    max_items = max(int(s[0]) for s in sizes)
    max_bytes = max(int(s[1]) for s in sizes)
    table = torch.full((max_items, 2), -1, dtype=torch.int64, device=device)
    if idx:
        table[: len(idx), 0] = torch.tensor(idx, dtype=torch.int64, device=device)
        table[: len(idx), 1] = torch.tensor(lens, dtype=torch.int64, device=device)
    payload = torch.zeros(max(max_bytes, 1), dtype=torch.uint8, device=device)
    if idx:
        flat = np.concatenate([np.ascontiguousarray(local_items[i], dtype=np.uint8).ravel() for i in idx])
        payload[: flat.size] = torch.from_numpy(flat).to(device)
    if rank == dst:
        tables = [torch.empty_like(table) for _ in range(world)]
        payloads = [torch.empty_like(payload) for _ in range(world)]
    else:
        tables = payloads = None
    dist.gather(table, tables, dst=dst, group=group)
    dist.gather(payload, payloads, dst=dst, group=group)
    if rank != dst:
        return None
    out: List[Optional[np.ndarray]] = [None] * n_items
    for r in range(world):
        t = tables[r].cpu().numpy()
        p = payloads[r].cpu().numpy()
        pos = 0
        for gi, ln in t:
            if gi < 0:
                continue
            out[int(gi)] = p[pos: pos + int(ln)].copy()
            pos += int(ln)
    missing = [i for i, v in enumerate(out) if v is None]
    if missing:
        raise RuntimeError(f"items never encoded by any rank: {missing[:8]}")
    return out  # type: ignore[return-value]
