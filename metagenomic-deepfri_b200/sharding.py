"""Length-balanced sharding of independent proteins across GPUs.

Proteins carry no cross-protein state (`pipeline.py:301-319` is a plain loop), so the path
shards with no data-path collective: greedy longest-processing-time assignment on the cost
model c(L) = alpha L^2 + beta L, then each rank runs its bin in chunks and only the final
score matrix is gathered.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

# seconds per protein = ALPHA L^2 + BETA L, fitted to the per-stage CUDA-event profile of one 16,384-protein configs[4] chunk on a
# B200 (`python tools/fit_cost.py profiles/r02b_bench_config4.json`: T = 4,813,437 residues, sum L^2 = 1.97e9; quadratic stages =
# contact maps + tile scan + adjacency product 13.05 ms, linear stages = LSTM-LM + embedding + X.W + head 47.22 ms; before the
# round-2 LSTM work the linear stages took 54.74 ms: BETA 1.117e-8)
ALPHA, BETA = 6.62e-12, 9.81e-9


def cost(lengths: Sequence[int], alpha: float = ALPHA, beta: float = BETA) -> np.ndarray:
    L = np.asarray(lengths, dtype=np.float64)
    return alpha * L * L + beta * L


def lpt_bins(lengths: Sequence[int], n_bins: int, alpha: float = ALPHA, beta: float = BETA) -> List[np.ndarray]:
    """Greedy LPT: heaviest protein first onto the currently lightest bin (heap: a million proteins in about a second).
    Returns, per bin, the protein indices in ascending index order - a random mix of lengths, which is what a chunk of the
    bin should look like (the LSTM kernel sorts each chunk itself)."""
    import heapq
    c = cost(lengths, alpha, beta)
    order = np.argsort(-c, kind="stable")
    heap = [(0.0, b) for b in range(n_bins)]
    owner = np.empty(len(c), np.int32)
    for i, ci in zip(order.tolist(), c[order].tolist()):
        load, b = heap[0]
        owner[i] = b
        heapq.heapreplace(heap, (load + ci, b))
    return [np.flatnonzero(owner == b).astype(np.int64) for b in range(n_bins)]


def chunks_by_residues(indices: np.ndarray, lengths: Sequence[int], max_residues: int,
                       max_proteins: int = 1 << 30) -> List[np.ndarray]:
    """Split one bin into launch-sized chunks bounded by total residues (HBM workspace)."""
    out, cur, tot = [], [], 0
    for i in indices:
        L = int(lengths[int(i)])
        if cur and (tot + L > max_residues or len(cur) >= max_proteins):
            out.append(np.asarray(cur, dtype=np.int64))
            cur, tot = [], 0
        cur.append(int(i))
        tot += L
    if cur:
        out.append(np.asarray(cur, dtype=np.int64))
    return out


def imbalance(bins: List[np.ndarray], lengths: Sequence[int]) -> float:
    c = cost(lengths)
    loads = np.array([c[b].sum() if len(b) else 0.0 for b in bins])
    return float(loads.max() / max(loads.mean(), 1e-30))
