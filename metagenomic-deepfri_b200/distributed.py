"""One process per GPU: shard a protein set by length-balanced bins, stream each bin through the path in chunks, gather the
score matrix on rank 0.  `torch.distributed` is used for the rendezvous and the single final gather only - there is no
collective on the compute path (the reference's loop over proteins is embarrassingly parallel,
`pipeline.py:301-319`)."""
from __future__ import annotations

from collections import deque
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .sharding import chunks_by_residues, lpt_bins

#: one chunk = one launch of the whole path; bounded by proteins and residues (HBM workspace ~ 6 KB per residue)
CHUNK_PROTEINS, CHUNK_RESIDUES = 16384, 5_200_000


def shard_job(lengths: Sequence[int], world: int, max_proteins: int = CHUNK_PROTEINS,
              max_residues: int = CHUNK_RESIDUES) -> List[List[np.ndarray]]:
    """LPT bins over `world` ranks, each cut into launch-sized chunks: `result[rank]` = list of protein-index arrays."""
    return [chunks_by_residues(b, lengths, max_residues, max_proteins) for b in lpt_bins(lengths, world)]


def stream_chunks(submit: Callable[[object, np.ndarray], object], chunks: Iterable[Tuple[object, int]], out: np.ndarray,
                  depth: int = 2) -> None:
    """Software pipeline over chunks: `submit(chunk, out_rows)` returns a job with `.wait()`; at most `depth` jobs are in flight,
    so chunk k + 1 is packed and copied while chunk k computes (`Predictor.submit_structures`).  `chunks` yields
    `(chunk, n_proteins)`; the scores of the chunks land in consecutive rows of `out`."""
    jobs = deque()
    row = 0
    for chunk, n in chunks:
        jobs.append(submit(chunk, out[row:row + n]))
        row += n
        if len(jobs) >= depth:
            jobs.popleft().wait()
    while jobs:
        jobs.popleft().wait()


def gather_scores(local_idx: np.ndarray, local_scores: np.ndarray, n_total: int, n_terms: int,
                  dst: int = 0) -> Optional[np.ndarray]:
    """Collect every rank's (indices, scores) on `dst` as one [n_total, C] float32 matrix in original protein order.
    Tensors, not pickles: one `gather` of the padded score blocks and one of the index vectors (NCCL: over NVLink, with the
    scatter into protein order done on the GPU and one device->host copy into pinned memory; gloo: on the host)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        if len(local_idx) == n_total and np.array_equal(local_idx, np.arange(n_total)):
            return local_scores
        out = np.zeros((n_total, n_terms), np.float32)
        out[local_idx] = local_scores
        return out
    rank, world = dist.get_rank(), dist.get_world_size()
    cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
    n_local = torch.tensor([len(local_idx)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    nmax = max(max(counts), 1)
    sc = torch.zeros((nmax, n_terms), dtype=torch.float32, device=dev)
    ix = torch.zeros(nmax, dtype=torch.int64, device=dev)
    if len(local_idx):
        sc[:len(local_idx)].copy_(torch.from_numpy(np.ascontiguousarray(local_scores, np.float32)), non_blocking=True)
        ix[:len(local_idx)].copy_(torch.from_numpy(np.ascontiguousarray(local_idx, np.int64)), non_blocking=True)
    sc_all = [torch.empty_like(sc) for _ in range(world)] if rank == dst else None
    ix_all = [torch.empty_like(ix) for _ in range(world)] if rank == dst else None
    dist.gather(sc, sc_all, dst=dst)
    dist.gather(ix, ix_all, dst=dst)
    if rank != dst:
        return None
    final = torch.zeros((n_total, n_terms), dtype=torch.float32, device=dev)
    for r in range(world):
        if counts[r]:
            final.index_copy_(0, ix_all[r][:counts[r]], sc_all[r][:counts[r]])
    if cuda:
        host = torch.empty((n_total, n_terms), dtype=torch.float32, pin_memory=True)
        host.copy_(final, non_blocking=True)
        torch.cuda.synchronize()
        return host.numpy()
    return final.numpy()


def predict_sharded(forward: Callable[[np.ndarray], np.ndarray], lengths: Sequence[int], n_terms: int,
                    rank: int, world: int, max_residues: int = CHUNK_RESIDUES,
                    max_proteins: int = CHUNK_PROTEINS) -> Optional[np.ndarray]:
    """`forward(indices) -> scores[len(indices), C]` is run on this rank's LPT bin chunk by chunk; rank 0 returns the
    [len(lengths), C] matrix in input order, the other ranks None."""
    mine = shard_job(lengths, world, max_proteins, max_residues)[rank]
    parts = [forward(ch) for ch in mine]
    local_idx = np.concatenate(mine) if mine else np.zeros(0, np.int64)
    local_sc = np.concatenate(parts) if parts else np.zeros((0, n_terms), np.float32)
    return gather_scores(local_idx, local_sc, len(lengths), n_terms)
