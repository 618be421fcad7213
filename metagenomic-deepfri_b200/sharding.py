"""Length-balanced sharding of independent proteins across GPUs.

Proteins carry no cross-protein state (`pipeline.py:301-319` is a plain loop), so the path
shards with no data-path collective: greedy longest-processing-time assignment on the cost
model c(L) = alpha L^2 + beta L, then each rank runs its bin in chunks and only the final
score matrix is gathered.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

# fitted on B200 (tensor-core engine): the linear term (LSTM + dense) dominates below L ~ 1000
ALPHA, BETA = 3.0e-4, 1.0


def cost(lengths: Sequence[int], alpha: float = ALPHA, beta: float = BETA) -> np.ndarray:
    L = np.asarray(lengths, dtype=np.float64)
    return alpha * L * L + beta * L


def lpt_bins(lengths: Sequence[int], n_bins: int, alpha: float = ALPHA, beta: float = BETA) -> List[np.ndarray]:
    """Greedy LPT: heaviest protein first onto the currently lightest bin.  Returns, per bin, the
    protein indices sorted by descending length (the order the LSTM kernel wants)."""
    c = cost(lengths, alpha, beta)
    order = np.argsort(-c, kind="stable")
    loads = np.zeros(n_bins)
    bins: List[List[int]] = [[] for _ in range(n_bins)]
    for i in order:
        b = int(np.argmin(loads))
        bins[b].append(int(i))
        loads[b] += c[i]
    return [np.asarray(b, dtype=np.int64) for b in bins]


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
