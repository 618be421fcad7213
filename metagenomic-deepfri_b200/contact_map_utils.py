"""Drop-in for `mDeepFRI.contact_map_utils` (Cython, `contact_map_utils.pyx`) on B200.

Same names, argument meaning and dtype strictness as the reference; the arithmetic runs in the
CUDA kernels of `csrc/cmap_kernels.cu` through the C ABI.  `threads` is accepted and ignored.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _require_buffer(a, dtype, what):
    # the reference's typed memoryviews reject anything but exact-dtype C-contiguous ndarrays
    if not isinstance(a, np.ndarray):
        raise TypeError(f"{what}: expected a numpy.ndarray, got {type(a).__name__}")
    if a.dtype != dtype:
        raise ValueError(f"Buffer dtype mismatch, expected '{np.dtype(dtype).name}' but got '{a.dtype.name}'")
    if a.ndim != 2:
        raise ValueError(f"Buffer has wrong number of dimensions (expected 2, got {a.ndim})")
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("ndarray is not C-contiguous")


def pairwise_sqeuclidean(X: np.ndarray, threads: int = 1) -> np.ndarray:
    """`contact_map_utils.pyx:17-37`: D[i,j] = sum_k (X[i,k]-X[j,k])^2 in unfused fp32."""
    _require_buffer(X, np.float32, "X")
    n, m = X.shape
    D = np.empty((n, n), np.float32)
    ctx = _lib.default_context()
    _lib.check(_lib.lib().mdf_pairwise_sqeuclidean(ctx.handle, _lib.fp(X), n, m, _lib.fp(D)))
    return D


def align_contact_map(query_alignment: str, target_alignment: str, sparse_target_contact_map: np.ndarray,
                      generated_contacts: int = 2, threads: int = 1) -> np.ndarray:
    """`contact_map_utils.pyx:44-117`: transfer a sparse target contact map onto the query."""
    q = query_alignment.encode("ascii")
    t = target_alignment.encode("ascii")
    if len(q) != len(t):
        raise ValueError("query and target alignments must have the same number of columns")
    _require_buffer(sparse_target_contact_map, np.int32, "sparse_target_contact_map")
    sp = sparse_target_contact_map
    if sp.shape[0] and sp.shape[1] != 2:
        raise ValueError("sparse_target_contact_map must have shape (nnz, 2)")
    ctx = _lib.default_context()
    L = _lib.lib()
    lq = C.c_int(0)
    _lib.check(L.mdf_align_contact_map(ctx.handle, q, t, len(q), _lib.ip(sp), sp.shape[0], int(generated_contacts),
                                       None, C.byref(lq)))
    out = np.empty((lq.value, lq.value), np.int32)
    if lq.value:
        _lib.check(L.mdf_align_contact_map(ctx.handle, q, t, len(q), _lib.ip(sp), sp.shape[0],
                                           int(generated_contacts), _lib.ip(out), C.byref(lq)))
    return out
