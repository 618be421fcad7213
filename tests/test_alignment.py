"""CPU tests of the alignment oracle (oracle/nw_oracle.c: Gotoh global alignment as the reference requests it from PyOpal) against
the reference's own known-answer tests (mDeepFRI/tests/test_alignment.py:22-45) and an independent DP, and of the host-side
drop-in pieces that need no GPU (insert_gaps, AlignmentResult)."""
import numpy as np
import pytest

import nw_oracle as nw
from metagenomic_deepfri_b200 import alignment

QUERY = "MAGFLKVVQLLAKYGSKAVQWAWANKGKILDWLNAGQAIDWVVS"
TARGETS = dict(seq1="MESILDLQELETSEEESALMAASTVSNNC", seq2="MKKAVIVENKGCATCSIGAACLVDGPIPDFEIAGATGLFGLWG",
               seq3="MAGFLKVVQILAKYGSKAVQWAWANKGKILDWINAGQAIDWVVE", seq4="MAGFLKVVQILAKYGSKAVQWAWANKGKILDWINAGQAIDWVVE")


def test_reference_known_answers_oracle():
    # tests/test_alignment.py:22-26: the best hit is seq3 (seq4 is identical: the FIRST maximum wins)
    scores = {k: nw.align(QUERY, t, full=False)[0] for k, t in TARGETS.items()}
    assert max(scores, key=lambda k: scores[k]) == "seq3" and scores["seq3"] == scores["seq4"]
    # :28-34
    _, ops = nw.align(QUERY, TARGETS["seq3"])
    assert ops == "MMMMMMMMMXMMMMMMMMMMMMMMMMMMMMMMXMMMMMMMMMMX"
    assert round(ops.count("M") / len(ops), 2) == 0.93


@pytest.mark.parametrize("fn", [nw.insert_gaps, alignment.insert_gaps])
def test_insert_gaps_dialect(fn):
    # tests/test_alignment.py:38-45
    assert fn('AACT', 'AAT', 'MMDM') == ('AACT', 'AA-T')
    assert fn('AAT', 'AATC', 'MMMI') == ('AAT-', 'AATC')
    assert fn('AAT', 'FGTC', 'XXMI') == ('AAT-', 'FGTC')


def test_blosum62_tables_agree_and_are_symmetric():
    assert np.array_equal(nw.BLOSUM62, alignment.BLOSUM62) and nw.BLOSUM62_ALPHABET == alignment.BLOSUM62_ALPHABET
    assert np.array_equal(nw.BLOSUM62, nw.BLOSUM62.T)
    assert nw.BLOSUM62[nw.BLOSUM62_ALPHABET.index("W"), nw.BLOSUM62_ALPHABET.index("W")] == 11


def test_oracle_is_optimal_and_consistent():
    """Score == an independent three-state DP; the alignment string spells both sequences and re-scores to the optimum."""
    rng = np.random.default_rng(0)
    letters = list("ARNDCQEGHILKMFPSTWYV")
    for trial in range(300):
        a = "".join(rng.choice(letters, rng.integers(0, 30)))
        b = a if trial % 7 == 0 else "".join(rng.choice(letters, rng.integers(0, 30)))
        if trial % 5 == 0 and len(a) > 4:                     # related sequences with an indel: gaps really occur
            cut = int(rng.integers(1, len(a) - 1))
            b = a[:cut] + a[cut + int(rng.integers(1, 3)):]
        go, ge = (10, 1) if trial % 3 else (int(rng.integers(0, 6)), int(rng.integers(0, 3)))
        go = max(go, ge)            # the Gotoh recurrences equal the affine-gap model when opening costs at least an extension
        s, ops = nw.align(a, b, gap_open=go, gap_extend=ge)
        assert s == nw.score_python(a, b, gap_open=go, gap_extend=ge)
        assert s == nw.score_of_ops(a, b, ops, gap_open=go, gap_extend=ge)
        gq, gt = nw.insert_gaps(a, b, ops)
        assert len(gq) == len(gt) == len(ops) and gq.replace("-", "") == a and gt.replace("-", "") == b
    assert nw.align("", "") == (0, "") and nw.align("ACD", "") == (-12, "DDD") and nw.align("", "AC") == (-11, "II")
    with pytest.raises(ValueError, match="alphabet"):
        nw.align("ACJ", "ACD")


def test_alignment_result_mirror():
    r = alignment.AlignmentResult("q", "AACT", "t", "AAT", "MMDM", 0.75, query_coverage=1.0, target_coverage=1.0)
    assert (r.gapped_sequence, r.gapped_target) == ("AACT", "AA-T")
    assert r.coords is None and r.cmap is None and r.aligned_cmap is None and r.db_name is None
    assert "query_name=q" in repr(r)
    with pytest.raises(RuntimeError, match="scoring_matrices"):
        alignment.resolve_matrix("VTML80")                     # resolved through the absent package, never guessed
    alpha, m = alignment.resolve_matrix(("AC", np.array([[1, -1], [-1, 1]])))
    assert alpha == "AC" and m.dtype == np.int8
