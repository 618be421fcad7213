"""Drop-in for the alignment step of `mDeepFRI.alignment` (`alignment.py:38-250`): `insert_gaps`, `AlignmentResult`,
`best_hit_database`, `align_pairwise`, `pairwise_against_database` - and the batched forms `align_pairs` / `best_hits` that
`align_mmseqs_results` (`alignment.py:266-322`) should call instead of a `multiprocessing.Pool` of PyOpal calls: every
(query, target) pair of a batch is aligned by one launch of the GPU Needleman-Wunsch kernels (`csrc/nw_align.cu`).

Scoring matrices: the reference's default is VTML80 from the `scoring_matrices` package (`alignment.py:166`).  That package is
resolved at run time when a matrix is given by name; BLOSUM62 is built in.  A matrix can always be passed explicitly as
`(alphabet, int8 [A, A] array)`.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib

Matrix = Union[str, Tuple[str, np.ndarray]]

BLOSUM62_ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*"
_B62 = (
    "4 -1 -2 -2 0 -1 -1 0 -2 -1 -1 -1 -1 -2 -1 1 0 -3 -2 0 -2 -1 0 -4 -1 5 0 -2 -3 1 0 -2 0 -3 -2 2 -1 -3 -2 -1 -1 -3 -2 -3 -1 0 -1 -4 "
    "-2 0 6 1 -3 0 0 0 1 -3 -3 0 -2 -3 -2 1 0 -4 -2 -3 3 0 -1 -4 -2 -2 1 6 -3 0 2 -1 -1 -3 -4 -1 -3 -3 -1 0 -1 -4 -3 -3 4 1 -1 -4 "
    "0 -3 -3 -3 9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4 -1 1 0 0 -3 5 2 -2 0 -3 -2 1 0 -3 -1 0 -1 -2 -1 -2 0 3 -1 -4 "
    "-1 0 0 2 -4 2 5 -2 0 -3 -3 1 -2 -3 -1 0 -1 -3 -2 -2 1 4 -1 -4 0 -2 0 -1 -3 -2 -2 6 -2 -4 -4 -2 -3 -3 -2 0 -2 -2 -3 -3 -1 -2 -1 -4 "
    "-2 0 1 -1 -3 0 0 -2 8 -3 -3 -1 -2 -1 -2 -1 -2 -2 2 -3 0 0 -1 -4 -1 -3 -3 -3 -1 -3 -3 -4 -3 4 2 -3 1 0 -3 -2 -1 -3 -1 3 -3 -3 -1 -4 "
    "-1 -2 -3 -4 -1 -2 -3 -4 -3 2 4 -2 2 0 -3 -2 -1 -2 -1 1 -4 -3 -1 -4 -1 2 0 -1 -3 1 1 -2 -1 -3 -2 5 -1 -3 -1 0 -1 -3 -2 -2 0 1 -1 -4 "
    "-1 -1 -2 -3 -1 0 -2 -3 -2 1 2 -1 5 0 -2 -1 -1 -1 -1 1 -3 -1 -1 -4 -2 -3 -3 -3 -2 -3 -3 -3 -1 0 0 -3 0 6 -4 -2 -2 1 3 -1 -3 -3 -1 -4 "
    "-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4 7 -1 -1 -4 -3 -2 -2 -1 -2 -4 1 -1 1 0 -1 0 0 0 -1 -2 -2 0 -1 -2 -1 4 1 -3 -2 -2 0 0 0 -4 "
    "0 -1 0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1 1 5 -2 -2 0 -1 -1 0 -4 -3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1 1 -4 -3 -2 11 2 -3 -4 -3 -2 -4 "
    "-2 -2 -2 -3 -2 -1 -2 -3 2 -1 -1 -2 -1 3 -3 -2 -2 2 7 -1 -3 -2 -1 -4 0 -3 -3 -3 -1 -2 -2 -3 -3 3 1 -2 1 -1 -2 -2 0 -3 -1 4 -3 -2 -1 -4 "
    "-2 -1 3 4 -3 0 1 -1 0 -3 -4 0 -3 -3 -2 0 -1 -4 -3 -3 4 1 -1 -4 -1 0 0 1 -3 3 4 -2 0 -3 -3 1 -1 -3 -1 0 -1 -3 -2 -2 1 4 -1 -4 "
    "0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2 0 0 -2 -1 -1 -1 -1 -1 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 1")
BLOSUM62 = np.array(_B62.split(), dtype=np.int8).reshape(24, 24)


def resolve_matrix(scoring_matrix: Matrix) -> Tuple[str, np.ndarray]:
    """-> (alphabet, int8 [A, A]).  Names other than BLOSUM62 come from `scoring_matrices.ScoringMatrix.from_name`, exactly as the
    reference resolves them (`alignment.py:177-178`)."""
    if not isinstance(scoring_matrix, str):
        alphabet, m = scoring_matrix
        m = np.ascontiguousarray(m, np.int8)
        if m.shape != (len(alphabet), len(alphabet)) or len(alphabet) > 32:
            raise ValueError("scoring matrix must be [A, A] with A = len(alphabet) <= 32")
        return str(alphabet), m
    if scoring_matrix.upper() == "BLOSUM62":
        return BLOSUM62_ALPHABET, BLOSUM62
    try:
        from scoring_matrices import ScoringMatrix
    except ImportError as e:
        raise RuntimeError(f"scoring matrix {scoring_matrix!r} is resolved through the `scoring_matrices` package (as in "
                           "mDeepFRI.alignment), which is not installed; pass (alphabet, matrix) explicitly") from e
    sm = ScoringMatrix.from_name(scoring_matrix)
    alphabet = str(sm.alphabet)
    m = np.array([[sm[a][b] if not hasattr(sm, "matrix") else sm.matrix[i][j] for j, b in enumerate(alphabet)]
                  for i, a in enumerate(alphabet)])
    if not np.array_equal(m, np.round(m)) or np.abs(m).max() > 127:
        raise ValueError(f"scoring matrix {scoring_matrix!r} is not integral in int8 range")
    return alphabet, m.astype(np.int8)


def insert_gaps(sequence: str, reference: str, alignment_string: str) -> Tuple[str, str]:
    """`alignment.py:38-62`: 'I' puts '-' into the query at that column, 'D' into the target."""
    s: List[str] = list(sequence)
    r: List[str] = list(reference)
    for i, a in enumerate(alignment_string):
        if a == "I":
            s.insert(i, "-")
        elif a == "D":
            r.insert(i, "-")
    return "".join(s), "".join(r)


class AlignmentResult:
    """`alignment.py:65-150`: same constructor, attributes and gapped strings."""

    def __init__(self, query_name: str = "", query_sequence: str = "", target_name: str = "", target_sequence: str = "",
                 alignment: str = "", query_identity: Optional[float] = None, query_coverage: Optional[float] = None,
                 target_coverage: Optional[float] = None, db_name: Optional[str] = None, coords: Optional[np.ndarray] = None):
        self.query_name = query_name
        self.query_sequence = query_sequence
        self.target_name = target_name
        self.target_sequence = target_sequence
        self.alignment = alignment
        self.query_identity = query_identity
        self.query_coverage = query_coverage
        self.target_coverage = target_coverage
        self.insert_gaps()
        self.db_name = db_name
        self.coords = coords
        self.target_coords = None
        self.cmap = None
        self.aligned_cmap = None

    def __repr__(self):
        return f"AlignmentResult(query_name={self.query_name}, target_name={self.target_name}, " \
               f"query_identity={self.query_identity}, query_coverage={self.query_coverage})"

    __str__ = __repr__

    def insert_gaps(self):
        self.gapped_sequence, self.gapped_target = insert_gaps(self.query_sequence, self.target_sequence, self.alignment)


def nw_align(queries: Sequence[str], targets: Sequence[str], gap_open: int = 10, gap_extend: int = 1,
             scoring_matrix: Matrix = "VTML80", full: bool = True) -> Tuple[np.ndarray, List[str]]:
    """n global alignments in one launch: (int32 scores [n], alignment strings over M/X/I/D - empty strings when `full` is
    False).  Sequences are uppercased like `alignment.py:152-160`."""
    n = len(queries)
    if len(targets) != n:
        raise ValueError("queries and targets differ in length")
    scores = np.empty(n, np.int32)
    if n == 0:
        return scores, []
    alphabet, m = resolve_matrix(scoring_matrix)
    host = _lib.pyhost()
    qs = [q.upper() for q in queries]
    ts = [t.upper() for t in targets]
    qp, ql = host.pointers(qs)
    tp, tl = host.pointers(ts)
    ctx = _lib.default_context()
    L = _lib.lib()
    if not full:
        _lib.check(L.mdf_nw_align(ctx.handle, n, qp, ql, tp, tl, m.ctypes.data, alphabet.encode("ascii"), len(alphabet), int(gap_open),
                                  int(gap_extend), scores.ctypes.data, None, None, None))
        return scores, [""] * n
    lens = np.frombuffer(ql, np.int32).astype(np.int64) + np.frombuffer(tl, np.int32)
    off = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    ops = np.empty(int(off[-1]) + 1, np.uint8)
    ops_len = np.empty(n, np.int32)
    _lib.check(L.mdf_nw_align(ctx.handle, n, qp, ql, tp, tl, m.ctypes.data, alphabet.encode("ascii"), len(alphabet), int(gap_open),
                              int(gap_extend), scores.ctypes.data, ops.ctypes.data, _lib.lp(off), ops_len.ctypes.data))
    buf = ops.tobytes()
    return scores, [buf[off[p]:off[p] + ops_len[p]].decode("ascii") for p in range(n)]


def _stats(alignment: str, query: str, target: str) -> Tuple[float, float, float]:
    """PyOpal's `identity()` (matching columns / alignment length) and `coverage()` (aligned span / sequence length; a global
    alignment spans both sequences)."""
    ident = alignment.count("M") / len(alignment) if alignment else 0.0
    return ident, 1.0 if query else 0.0, 1.0 if target else 0.0


def align_pairwise(query: str, target: str, gap_open: int = 10, gap_extend: int = 1, scoring_matrix: Matrix = "VTML80"):
    """`alignment.py:196-221` -> (alignment_string, identity, query_coverage, target_coverage)."""
    _, ops = nw_align([query], [target], gap_open, gap_extend, scoring_matrix)
    return (ops[0],) + _stats(ops[0], query, target)


def best_hits(queries: Sequence[str], target_dicts: Sequence[Dict[str, str]], gap_open: int = 10, gap_extend: int = 1,
              scoring_matrix: Matrix = "VTML80") -> List[Tuple[str, str]]:
    """`best_hit_database` (`alignment.py:163-194`) for many queries at once: every (query, candidate) pair is scored in ONE
    launch; per query the first candidate with the maximal score wins (Python's `max` over results in database order)."""
    flat_q, flat_t, owner = [], [], []
    for k, (q, d) in enumerate(zip(queries, target_dicts)):
        for t in d.values():
            flat_q.append(q); flat_t.append(t); owner.append(k)
    scores, _ = nw_align(flat_q, flat_t, gap_open, gap_extend, scoring_matrix, full=False)
    out, at = [], 0
    for q, d in zip(queries, target_dicts):
        keys = list(d.keys())
        if not keys:
            raise ValueError("best_hits: a query has no candidate targets")
        s = scores[at:at + len(keys)]
        best = keys[int(np.argmax(s))]                  # argmax returns the first maximum
        out.append((best, d[best].upper()))
        at += len(keys)
    return out


def best_hit_database(query: str, target_sequences: Dict[str, str], gap_open: int = 10, gap_extend: int = 1,
                      scoring_matrix: Matrix = "VTML80"):
    return best_hits([query], [target_sequences], gap_open, gap_extend, scoring_matrix)[0]


def align_pairs(query_ids: Sequence[str], query_sequences: Sequence[str], target_dicts: Sequence[Dict[str, str]], gap_open: int = 10,
                gap_extend: int = 1, scoring_matrix: Matrix = "VTML80") -> List[AlignmentResult]:
    """The `Pool.starmap(pairwise_against_database, ...)` of `alignment.py:314-320` as two launches: score every candidate,
    keep the best per query, align those pairs in full."""
    qs = [q.upper() for q in query_sequences]
    best = best_hits(qs, target_dicts, gap_open, gap_extend, scoring_matrix)
    _, ops = nw_align(qs, [b[1] for b in best], gap_open, gap_extend, scoring_matrix)
    out = []
    for qid, q, (tid, t), a in zip(query_ids, qs, best, ops):
        ident, qc, tc = _stats(a, q, t)
        out.append(AlignmentResult(qid, q, tid, t, a, ident, query_coverage=qc, target_coverage=tc))
    return out


def pairwise_against_database(query_id: str, query_sequence: str, target_sequences: Dict[str, str], gap_open: int = 10,
                              gap_extend: int = 1, scoring_matrix: Matrix = "VTML80") -> AlignmentResult:
    """`alignment.py:223-250`."""
    return align_pairs([query_id], [query_sequence], [target_sequences], gap_open, gap_extend, scoring_matrix)[0]
