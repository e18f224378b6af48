"""Multi-GPU sharding of a window batch: windows are independent, so the path
shards with NO data-path collective — every rank polishes its own windows — and a single gather of
the consensus bytes to rank 0 puts the contig back together.  One process per GPU
(torch.distributed: NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

from .batch import WindowBatch


def window_cost(batch: WindowBatch) -> np.ndarray:
    """Estimated DP cells per window: sum over arms of (nodes_before + 1) * (len + 1), with the node
    count approximated by the running maximum arm length times a growth factor."""
    n_per = (batch.win["n_internal"] + batch.win["n_pre"] + batch.win["n_suf"]).astype(np.int64)
    first = batch.win["first_arm"].astype(np.int64)
    lens = batch.arms["len"].astype(np.int64)
    csum = np.concatenate([[0], np.cumsum(lens)])
    tot = csum[first + n_per] - csum[first]
    mean = np.where(n_per > 0, tot / np.maximum(n_per, 1), 0.0)
    return (1.5 * mean * tot + 64).astype(np.int64)


def shard_indices(batch: WindowBatch, world: int, rank: int) -> np.ndarray:
    """Cost-balanced deal: windows sorted by estimated cost (descending, stable) and dealt
    round-robin, so every rank gets the same mix of cheap and expensive windows.  Returned indices
    are ascending (the rank keeps the contig's window order)."""
    order = np.argsort(-window_cost(batch), kind="stable")
    mine = order[rank::world]
    return np.sort(mine)


def gather_consensus(local: Sequence[str], idx: np.ndarray, n_total: int, world: int, rank: int,
                     device: Optional[str] = None) -> Optional[List[str]]:
    """Final consensus gather: rank 0 receives every rank's (window index, consensus) pairs and
    returns the consensus strings in original window order; other ranks return None."""
    import torch
    import torch.distributed as dist

    blob = "".join(local).encode()
    lens = np.fromiter((len(s) for s in local), dtype=np.int64, count=len(local))
    if world == 1:
        out: List[Optional[str]] = [None] * n_total
        for i, s in zip(idx, local):
            out[int(i)] = s
        return out  # type: ignore[return-value]
    dev = torch.device(device) if device else torch.device("cpu")
    meta = torch.tensor([len(local), len(blob)], dtype=torch.int64, device=dev)
    metas = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(metas, meta)
    max_n = max(int(m[0]) for m in metas)
    max_b = max(int(m[1]) for m in metas)
    t_idx = torch.zeros(max_n, dtype=torch.int64, device=dev)
    t_len = torch.zeros(max_n, dtype=torch.int64, device=dev)
    t_blob = torch.zeros(max(max_b, 1), dtype=torch.uint8, device=dev)
    t_idx[: len(local)] = torch.from_numpy(np.asarray(idx, dtype=np.int64)).to(dev)
    t_len[: len(local)] = torch.from_numpy(lens).to(dev)
    if blob:
        t_blob[: len(blob)] = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    lists = None
    if rank == 0:
        lists = [[torch.zeros_like(t) for _ in range(world)] for t in (t_idx, t_len, t_blob)]
    for k, t in enumerate((t_idx, t_len, t_blob)):
        dist.gather(t, lists[k] if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    out = [None] * n_total
    for r in range(world):
        n_r = int(metas[r][0])
        ii = lists[0][r][:n_r].cpu().numpy()
        ll = lists[1][r][:n_r].cpu().numpy()
        raw = lists[2][r].cpu().numpy().tobytes()
        pos = 0
        for i, l in zip(ii, ll):
            out[int(i)] = raw[pos:pos + int(l)].decode()
            pos += int(l)
    return out  # type: ignore[return-value]


def polish_sharded(batch: WindowBatch, world: int, rank: int, run: Callable[[WindowBatch], List[str]],
                   device: Optional[str] = None) -> Optional[List[str]]:
    """Shard -> polish own windows with `run` (native.consensus on a GPU) -> gather to rank 0."""
    idx = shard_indices(batch, world, rank)
    local = run(batch.select(idx)) if len(idx) else []
    return gather_consensus(local, idx, batch.n_win, world, rank, device)
