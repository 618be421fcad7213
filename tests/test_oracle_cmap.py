"""CPU tests: the contact-map oracle against the reference's own known-answer tests, the compiled
reference (oracle/_ref, when built) and the committed golden vectors."""
import numpy as np
import pytest

import cmap_oracle as co

REF = co.ref_module()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref not built (needs /root/reference)")


def impls():
    out = [("port", co.pairwise_sqeuclidean, co.align_contact_map)]
    if REF is not None:
        out.append(("reference", REF.pairwise_sqeuclidean, REF.align_contact_map))
    return out


@pytest.mark.parametrize("name,pw,al", impls())
def test_reference_known_answers(name, pw, al):
    # mDeepFRI/tests/test_contact_map_utils.py:16-25 (values listed there; the reference forgot the assert)
    np.random.seed(42)
    m = np.random.rand(3, 3).astype(np.float32)
    exp = np.array([[0, 1.01354558, 0.12442072], [1.01354558, 0, 0.99467713], [0.12442072, 0.99467713, 0]], np.float32)
    assert np.allclose(pw(m), exp)
    # :99-110 stress case
    N = 100
    tc = np.array([[i, i + 1] for i in range(N - 1)], dtype=np.int32)
    r = al("A" * N, "A" * N, tc)
    assert r.shape == (N, N) and r[0, 1] == 1
    # :29-97 - the reference's expected matrices assume a symmetric write; the current code writes one
    # direction only (contact_map_utils.pyx:105-115).  With the symmetric sparse input the
    # pipeline really produces (np.argwhere on a symmetric map) the documented expectations hold:
    sym = lambda a: np.concatenate([a, a[:, ::-1]]).astype(np.int32)
    assert np.array_equal(al("AB", "AB", sym(np.array([[0, 1]]))), np.ones((2, 2), np.int32))
    assert np.array_equal(al("A-C", "ABC", sym(np.array([[0, 1], [1, 2], [0, 2]]))), np.ones((2, 2), np.int32))
    assert np.array_equal(al("ABC", "A-C", sym(np.array([[0, 1]])), generated_contacts=1), np.ones((3, 3), np.int32))
    # and the one-directional behaviour of the code as shipped:
    assert np.array_equal(al("AB", "AB", np.array([[0, 1]], np.int32)), np.array([[1, 1], [0, 1]], np.int32))


def test_three_point_map():
    # mDeepFRI/tests/test_conctact_map.py:36-41
    coords = np.array([[0, 0, 0], [5, 0, 0], [10, 0, 0]], np.float32)
    assert np.array_equal(co.calculate_contact_map(coords, 6.0), [[1, 1, 0], [1, 1, 1], [0, 1, 1]])
    # :20-28
    assert np.allclose(np.sqrt(co.pairwise_sqeuclidean(np.array([[0, 0, 0], [1, 1, 1]], np.float32))),
                       [[0, np.sqrt(3)], [np.sqrt(3), 0]])


def test_threshold_is_strict_and_float32():
    # bio_utils.py:214-220 under NumPy-2 weak-scalar promotion
    d = np.array([36.0, 35.999996, 36.000004], np.float32)
    assert list(d < co.threshold_sq(6)) == [False, True, False]
    x = np.array([[0, 0, 0], [6, 0, 0], [0, 5.9999995, 0]], np.float32)
    cm = co.calculate_contact_map(x, 6)
    assert cm[0, 1] == 0 and cm[0, 2] == 1


def test_port_matches_golden(cmap_golden):
    g = cmap_golden
    for k in range(int(g["n_cases"])):
        q, t = g[f"c{k}_q"].tobytes().decode(), g[f"c{k}_t"].tobytes().decode()
        coords = g[f"c{k}_coords"]
        thr, gen = g[f"c{k}_thr_gen"]
        thr = int(thr) if thr == int(thr) and k % 3 == 2 else float(thr)
        assert np.array_equal(co.pairwise_sqeuclidean(coords), g[f"c{k}_D"])
        sp = co.calculate_contact_map(coords, thr, "sparse") if len(coords) else np.zeros((0, 2), np.int32)
        assert np.array_equal(sp, g[f"c{k}_sparse"])
        assert np.array_equal(co.align_contact_map(q, t, sp, int(gen)), g[f"c{k}_aligned"])
        assert np.array_equal(co.pairwise_sqeuclidean_np(coords), g[f"c{k}_D"])


@needs_ref
def test_port_matches_reference_random():
    from metagenomic_deepfri_b200 import synth
    rng = np.random.default_rng(7)
    wl = synth.make_workload(60, 1, 300, seed=9, threshold=6.0)
    for i in range(len(wl)):
        c = wl.coords[i]
        assert np.array_equal(co.pairwise_sqeuclidean(c, threads=3), REF.pairwise_sqeuclidean(c, 2))
        sp = co.calculate_contact_map(c, (6.0, 10.0)[i % 2], "sparse")
        if i % 5 == 0:   # arbitrary (non-symmetric, out-of-range, negative) sparse input
            sp = rng.integers(-3, len(c) + 5, size=(50, 2)).astype(np.int32)
        gen = i % 4
        assert np.array_equal(co.align_contact_map(wl.gapped_query[i], wl.gapped_target[i], sp, gen, threads=2),
                              REF.align_contact_map(wl.gapped_query[i], wl.gapped_target[i], sp, gen, 2))


def test_insert_gaps_dialect():
    # mDeepFRI/tests/test_alignment.py:38-45
    assert co.insert_gaps("AACT", "AAT", "MMDM") == ("AACT", "AA-T")
    assert co.insert_gaps("AAT", "AATC", "MMMI") == ("AAT-", "AATC")
    assert co.insert_gaps("AAT", "FGTC", "XXMI") == ("AAT-", "FGTC")


def test_seq2onehot_oracle_and_product():
    # mDeepFRI/tests/test_predict.py:9-33, for the oracle twin and the product's host function
    from metagenomic_deepfri_b200 import predict
    for f in (co.seq2onehot, predict.seq2onehot):
        r = f("")
        assert r.shape == (0, 26) and r.dtype == np.float32
        assert np.all(f("D") == np.array([[0, 1] + [0] * 24]))
        r = f("-DGU")
        assert r.shape == (4, 26) and np.array_equal(r.argmax(1), [0, 1, 2, 3]) and r.sum() == 4
        with pytest.raises(ValueError):
            f("J")
    with pytest.raises(UnicodeEncodeError):
        predict.seq2onehot("ACDé")


def test_synthetic_alignments_are_consistent():
    from metagenomic_deepfri_b200 import synth
    wl = synth.make_workload(50, 5, 200, seed=3)
    for q, gq, gt, c in zip(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords):
        assert len(gq) == len(gt) and gq.replace("-", "") == q
        assert len(gt.replace("-", "")) == len(c) and c.dtype == np.float32
        assert co.align_contact_map(gq, gt, np.zeros((0, 2), np.int32)).shape == (len(q), len(q))
