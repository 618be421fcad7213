"""GPU parity tests of the batched Needleman-Wunsch kernels (`csrc/nw_align.cu`) through the C ABI against
oracle/nw_oracle.c: scores AND alignment strings bit for bit, for every strip width, several panels, empty sequences, other gap
penalties and matrices; the reference's known-answer tests through the drop-in functions."""
import numpy as np
import pytest

import nw_oracle as nw
from metagenomic_deepfri_b200 import alignment, synth
from test_alignment import QUERY, TARGETS

pytestmark = pytest.mark.gpu
B62 = (alignment.BLOSUM62_ALPHABET, alignment.BLOSUM62)


def related(rng, q, p_sub=0.3, p_indel=0.03):
    out = []
    for ch in q:
        u = rng.random()
        if u < p_indel:
            continue
        out.append(rng.choice(list(synth.AA20)) if u < p_indel + p_sub else ch)
        if rng.random() < p_indel:
            out.extend(rng.choice(list(synth.AA20), rng.integers(1, 6)))
    return "".join(out)


def check(queries, targets, go=10, ge=1, matrix=B62):
    scores, ops = alignment.nw_align(queries, targets, go, ge, matrix)
    only, _ = alignment.nw_align(queries, targets, go, ge, matrix, full=False)
    assert np.array_equal(scores, only)
    for q, t, s, o in zip(queries, targets, scores, ops):
        ws, wo = nw.align(q, t, matrix[1], matrix[0], go, ge)
        assert int(s) == ws, (len(q), len(t), int(s), ws)
        assert o == wo, (len(q), len(t))


def test_reference_known_answers():
    assert alignment.best_hit_database(QUERY, TARGETS, scoring_matrix=B62)[0] == "seq3"         # tests/test_alignment.py:22-26
    a, iden, qc, tc = alignment.align_pairwise(QUERY, TARGETS["seq3"], scoring_matrix=B62)         # :28-34
    assert a == "MMMMMMMMMXMMMMMMMMMMMMMMMMMMMMMMXMMMMMMMMMMX" and round(iden, 2) == 0.93 and qc == 1.0 and tc == 1.0
    r = alignment.pairwise_against_database("q1", QUERY.lower(), TARGETS, scoring_matrix=B62)
    assert r.target_name == "seq3" and r.alignment == a and r.gapped_sequence == QUERY and r.query_sequence == QUERY


def test_every_strip_width_and_panel_count():
    rng = np.random.default_rng(1)
    qs, ts = [], []
    for lt in (1, 2, 3, 4, 5, 31, 32, 33, 127, 128, 129, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 1500, 2100):
        for lq in (1, 7, min(lt + 3, 300), lt):
            if lq > 1200 and lt > 1200:
                lq = 1200
            q = "".join(rng.choice(list(synth.AA20), lq))
            qs.append(q)
            ts.append(related(rng, q)[:lt].ljust(lt, "A") if rng.random() < 0.7 else "".join(rng.choice(list(synth.AA20), lt)))
    check(qs, ts)


def test_edge_cases_and_parameters():
    rng = np.random.default_rng(2)
    qs = ["", "ACD", "", "A", "W" * 40, "ACDEFGHIKLMNPQRSTVWY" * 3]
    ts = ["", "", "ACDE", "A", "W" * 300, "ACDEFGHIKLMNPQRSTVWY"]
    check(qs, ts)
    qs = ["".join(rng.choice(list(synth.AA20), rng.integers(1, 200))) for _ in range(200)]
    ts = [related(rng, q) for q in qs]
    check(qs, ts)
    check(qs[:40], ts[:40], 0, 0)                  # free gaps: every tie rule is exercised
    check(qs[:40], ts[:40], 3, 3)
    check(qs[:40], ts[:40], 25, 0)
    alpha = "ACDEFGHIKLMNPQRSTVWYX"
    m = rng.integers(-6, 12, (len(alpha), len(alpha))).astype(np.int8)
    check(qs[:40], ts[:40], 7, 2, (alpha, m))      # an arbitrary (even asymmetric) matrix
    with pytest.raises(ValueError, match="alphabet"):
        alignment.nw_align(["ACJ"], ["ACD"], scoring_matrix=B62)
    assert alignment.nw_align([], [], scoring_matrix=B62)[1] == []


def test_metagenomic_batch_feeds_the_transfer_kernel():
    """2,048 keyed query / target pairs: the GPU alignment strings, through insert_gaps, are what the contact-map transfer
    consumes (gapped strings of equal length that spell the two sequences)."""
    wl = synth.keyed_workload(np.arange(2048), 5)
    targets = [t.replace("-", "") for t in wl.gapped_target]
    scores, ops = alignment.nw_align(wl.query_seqs, targets, scoring_matrix=B62)
    idx = np.random.default_rng(3).choice(2048, 96, replace=False)
    for i in idx:
        ws, wo = nw.align(wl.query_seqs[i], targets[i])
        assert int(scores[i]) == ws and ops[i] == wo
    for q, t, o in zip(wl.query_seqs, targets, ops):
        gq, gt = alignment.insert_gaps(q, t, o)
        assert len(gq) == len(gt) == len(o) and gq.replace("-", "") == q and gt.replace("-", "") == t
