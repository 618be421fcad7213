"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement of the reference's contact-map half of the hot path:
  * `libcmap_oracle.so` (oracle/cmap_oracle.c) through ctypes — pairwise_sqeuclidean,
    threshold, sparsify, align_contact_map;
  * NumPy glue restating `mDeepFRI/bio_utils.py:196-227` (calculate_contact_map) and
    `:348-385` (build_align_contact_map), `alignment.py:38-62` (insert_gaps) and
    `predict.pyx:17-48` (seq2onehot).

Parity is PINNED against the reference itself: `ref_module()` returns the unmodified
`contact_map_utils.pyx` compiled into oracle/_ref/ (oracle/Makefile), and
tests/test_oracle_cmap.py asserts bit equality between it, this port and tests/golden/.
"""
from __future__ import annotations

import ctypes
import importlib
import os
import subprocess
import sys
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None


def build(ref: bool = True) -> None:
    """Compile the C port (always) and the reference .pyx (only if /root/reference exists)."""
    subprocess.run(["make", "-C", _HERE, "port"], check=True, capture_output=True)
    if ref and os.path.exists("/root/reference/mDeepFRI/contact_map_utils.pyx"):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcmap_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = ctypes.CDLL(path)
        c_f = ctypes.POINTER(ctypes.c_float)
        c_i = ctypes.POINTER(ctypes.c_int32)
        L.mdf_oracle_pairwise_sqeuclidean.argtypes = [c_f, ctypes.c_int, ctypes.c_int, c_f, ctypes.c_int]
        L.mdf_oracle_pairwise_sqeuclidean.restype = None
        L.mdf_oracle_threshold.argtypes = [c_f, ctypes.c_int64, ctypes.c_float, c_i]
        L.mdf_oracle_threshold.restype = None
        L.mdf_oracle_sparsify.argtypes = [c_i, ctypes.c_int, c_i]
        L.mdf_oracle_sparsify.restype = ctypes.c_int64
        L.mdf_oracle_query_length.argtypes = [ctypes.c_char_p, ctypes.c_int]
        L.mdf_oracle_query_length.restype = ctypes.c_int
        L.mdf_oracle_align_contact_map.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, c_i,
                                                   ctypes.c_int64, ctypes.c_int, c_i, ctypes.c_int]
        L.mdf_oracle_align_contact_map.restype = ctypes.c_int
        _LIB = L
    return _LIB


def ref_module():
    """The reference's own compiled contact_map_utils (oracle/_ref), or None if not built."""
    global _REF
    if _REF is None:
        d = os.path.join(_HERE, "_ref")
        if not os.path.isdir(d):
            return None
        sys.path.insert(0, d)
        try:
            _REF = importlib.import_module("contact_map_utils")
        except ImportError:
            return None
        finally:
            sys.path.remove(d)
    return _REF


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


# ---- contact_map_utils.pyx:17-37
def pairwise_sqeuclidean(X: np.ndarray, threads: int = 1) -> np.ndarray:
    X = np.ascontiguousarray(X, dtype=np.float32)
    n, m = X.shape
    D = np.empty((n, n), np.float32)
    lib().mdf_oracle_pairwise_sqeuclidean(_fp(X), n, m, _fp(D), threads)
    return D


def pairwise_sqeuclidean_np(X: np.ndarray) -> np.ndarray:
    """Same arithmetic in NumPy fp32: ((0+dx^2)+dy^2)+dz^2, each op rounded to fp32."""
    X = np.asarray(X, dtype=np.float32)
    d = np.zeros((X.shape[0], X.shape[0]), np.float32)
    for k in range(X.shape[1]):
        diff = X[:, None, k] - X[None, :, k]
        d = d + diff * diff
    np.fill_diagonal(d, 0.0)
    return d


# ---- bio_utils.py:196-227
def threshold_sq(threshold) -> np.float32:
    """`threshold**2` is a Python scalar; NumPy 2 compares it as float32 (weak scalar)."""
    return np.float32(threshold ** 2)


def calculate_contact_map(coordinates: np.ndarray, threshold=6.0, mode: str = "matrix",
                          threads: int = 1) -> np.ndarray:
    D = pairwise_sqeuclidean(coordinates, threads)
    cmap = np.empty(D.shape, np.int32)
    lib().mdf_oracle_threshold(_fp(D), D.size, float(threshold_sq(threshold)), _ip(cmap))
    if mode == "sparse":
        n = cmap.shape[0]
        nnz = lib().mdf_oracle_sparsify(_ip(cmap), n, None)
        pairs = np.empty((nnz, 2), np.int32)
        lib().mdf_oracle_sparsify(_ip(cmap), n, _ip(pairs))
        return pairs
    return cmap


# ---- contact_map_utils.pyx:44-117
def align_contact_map(query_alignment: str, target_alignment: str, sparse_target_contact_map: np.ndarray,
                      generated_contacts: int = 2, threads: int = 1) -> np.ndarray:
    q = query_alignment.encode("ascii")
    t = target_alignment.encode("ascii")
    sp = np.ascontiguousarray(sparse_target_contact_map, dtype=np.int32).reshape(-1, 2)
    Lq = lib().mdf_oracle_query_length(q, len(q))
    out = np.empty((Lq, Lq), np.int32)
    r = lib().mdf_oracle_align_contact_map(q, t, len(q), _ip(sp), sp.shape[0], generated_contacts,
                                           _ip(out), threads)
    if r != Lq:
        raise MemoryError("oracle align_contact_map failed")
    return out


# ---- bio_utils.py:348-385 (without the AlignmentResult object / logging)
def build_align_contact_map(gapped_query: str, gapped_target: str, coords: Optional[np.ndarray],
                            threshold: float = 6, generated_contacts: int = 2) -> Optional[np.ndarray]:
    if coords is None:
        return None
    sparse = calculate_contact_map(coords, threshold=threshold, mode="sparse")
    return align_contact_map(gapped_query, gapped_target, sparse, generated_contacts)


# ---- alignment.py:38-62
def insert_gaps(sequence: str, reference: str, alignment_string: str) -> Tuple[str, str]:
    s, r = list(sequence), list(reference)
    for i, a in enumerate(alignment_string):
        if a == "I":
            s.insert(i, "-")
        elif a == "D":
            r.insert(i, "-")
    return "".join(s), "".join(r)


# ---- predict.pyx:17-48
ALPHABET = "-DGULNTKHYWCPVSOIEFXQABZRM"


def seq2onehot(seq: str) -> np.ndarray:
    b = seq.encode("ascii")
    out = np.zeros((len(b), 26), np.float32)
    for i, ch in enumerate(b):
        k = ALPHABET.find(chr(ch))
        if k < 0:
            raise ValueError(f"Invalid character in sequence: {seq[i]}")
        out[i, k] = 1.0
    return out
