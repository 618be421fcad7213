"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

Python face of oracle/nw_oracle.c (Gotoh global alignment as the reference requests it from PyOpal, alignment.py:163-221), plus
an independent pure-Python three-state DP used to check the C restatement's scores on small cases, `insert_gaps`
(alignment.py:38-62) and the BLOSUM62 table used by the tests (VTML80, the reference's default, lives in the absent
`scoring_matrices` package; its values are not reproduced here)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

BLOSUM62_ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*"
_B62 = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0 -4
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1 -4
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1 -4
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1 -4
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1 -4
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1 -4
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1 -4
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1 -4
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1 -4
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1 -4
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1 -4
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1 -4
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2 -4
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0 -4
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0 -4
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2 -4
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1 -4
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1 -4
-2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1 -4
-1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1 -4
-4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1
"""
BLOSUM62 = np.array(_B62.split(), dtype=np.int8).reshape(24, 24)


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libnw_oracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", _HERE, "port"], check=True, capture_output=True)
        L = ctypes.CDLL(path)
        L.mdf_oracle_nw.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
        L.mdf_oracle_nw.restype = ctypes.c_int32
        _LIB = L
    return _LIB


def encode(seq: str, alphabet: str) -> np.ndarray:
    lut = np.full(256, 255, np.uint8)
    lut[np.frombuffer(alphabet.encode(), np.uint8)] = np.arange(len(alphabet))
    codes = lut[np.frombuffer(seq.encode("ascii"), np.uint8)]
    if (codes == 255).any():
        raise ValueError(f"sequence holds a residue outside the scoring matrix alphabet: {seq[int(np.flatnonzero(codes == 255)[0])]}")
    return codes


def align(query: str, target: str, matrix: np.ndarray = BLOSUM62, alphabet: str = BLOSUM62_ALPHABET, gap_open: int = 10,
          gap_extend: int = 1, full: bool = True) -> Tuple[int, str]:
    """(score, alignment string over M/X/I/D) - align_pairwise's `alignment[0].alignment` (alignment.py:212-214)."""
    q, t = encode(query, alphabet), encode(target, alphabet)
    S = np.ascontiguousarray(matrix, np.int8)
    ops = ctypes.create_string_buffer(len(q) + len(t) + 1)
    n = ctypes.c_int(0)
    score = lib().mdf_oracle_nw(q.ctypes.data, len(q), t.ctypes.data, len(t), S.ctypes.data, S.shape[0], gap_open, gap_extend,
                                ops if full else None, ctypes.byref(n))
    return int(score), ops.raw[:n.value].decode() if full else ""


def score_python(query: str, target: str, matrix=BLOSUM62, alphabet=BLOSUM62_ALPHABET, gap_open=10, gap_extend=1) -> int:
    """Independent check of the optimum: three-state DP over alignment COLUMNS (state = kind of the last column), pure Python."""
    q, t = encode(query, alphabet), encode(target, alphabet)
    NEG = -10 ** 9
    lq, lt = len(q), len(t)
    best = {(0, 0, "S"): 0}
    M = [[NEG] * (lt + 1) for _ in range(lq + 1)]
    D = [[NEG] * (lt + 1) for _ in range(lq + 1)]
    I = [[NEG] * (lt + 1) for _ in range(lq + 1)]
    M[0][0] = 0
    for i in range(lq + 1):
        for j in range(lt + 1):
            if i and j:
                M[i][j] = max(M[i - 1][j - 1], D[i - 1][j - 1], I[i - 1][j - 1]) + int(matrix[q[i - 1], t[j - 1]])
            if i:
                D[i][j] = max(max(M[i - 1][j], I[i - 1][j]) - gap_open, D[i - 1][j] - gap_extend)
            if j:
                I[i][j] = max(max(M[i][j - 1], D[i][j - 1]) - gap_open, I[i][j - 1] - gap_extend)
    return max(M[lq][lt], D[lq][lt], I[lq][lt])


def score_of_ops(query: str, target: str, ops: str, matrix=BLOSUM62, alphabet=BLOSUM62_ALPHABET, gap_open=10, gap_extend=1) -> int:
    """Score an alignment string column by column (and check that it spells both sequences)."""
    q, t = encode(query, alphabet), encode(target, alphabet)
    i = j = s = 0
    prev = ""
    for o in ops:
        if o in "MX":
            assert (query[i] == target[j]) == (o == "M")
            s += int(matrix[q[i], t[j]]); i += 1; j += 1
        elif o == "D":
            s -= gap_extend if prev == "D" else gap_open; i += 1
        elif o == "I":
            s -= gap_extend if prev == "I" else gap_open; j += 1
        else:
            raise AssertionError(o)
        prev = o
    assert i == len(q) and j == len(t)
    return s


def insert_gaps(sequence: str, reference: str, alignment_string: str) -> Tuple[str, str]:
    """alignment.py:38-62."""
    s, r = list(sequence), list(reference)
    for i, a in enumerate(alignment_string):
        if a == "I":
            s.insert(i, "-")
        elif a == "D":
            r.insert(i, "-")
    return "".join(s), "".join(r)
