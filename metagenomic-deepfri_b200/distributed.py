"""One process per GPU: shard a protein set by length-balanced bins, run each bin locally, gather
the score matrix on rank 0.  `torch.distributed` is used for the rendezvous and the single final
gather only - there is no collective on the compute path."""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np

from .sharding import chunks_by_residues, lpt_bins


def gather_scores(local_idx: np.ndarray, local_scores: np.ndarray, n_total: int, n_terms: int,
                  dst: int = 0) -> Optional[np.ndarray]:
    """Collect every rank's (indices, scores) on `dst` and scatter them into [n_total, C]."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        out = np.zeros((n_total, n_terms), np.float32)
        out[local_idx] = local_scores
        return out
    rank, world = dist.get_rank(), dist.get_world_size()
    payload = (np.asarray(local_idx, np.int64), np.ascontiguousarray(local_scores, np.float32))
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if rank != dst:
        return None
    out = np.zeros((n_total, n_terms), np.float32)
    for idx, sc in gathered:
        out[idx] = sc
    return out


def predict_sharded(forward: Callable[[np.ndarray], np.ndarray], lengths: Sequence[int], n_terms: int,
                    rank: int, world: int, max_residues: int = 400_000) -> Optional[np.ndarray]:
    """`forward(indices) -> scores[len(indices), C]` is run on this rank's LPT bin in chunks."""
    bins = lpt_bins(lengths, world)
    mine = bins[rank]
    parts, idxs = [], []
    for ch in chunks_by_residues(mine, lengths, max_residues):
        parts.append(forward(ch))
        idxs.append(ch)
    local_idx = np.concatenate(idxs) if idxs else np.zeros(0, np.int64)
    local_sc = np.concatenate(parts) if parts else np.zeros((0, n_terms), np.float32)
    return gather_scores(local_idx, local_sc, len(lengths), n_terms)
