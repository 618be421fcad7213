"""GPU parity tests (through the C ABI) of the contact-map half: bit-exact against the oracle,
the compiled reference when present, and the committed golden vectors."""
import numpy as np
import pytest

import cmap_oracle as co
from conftest import Aln
from metagenomic_deepfri_b200 import batching, bio_utils, contact_map, contact_map_utils, synth

pytestmark = pytest.mark.gpu


def test_golden_vectors(cmap_golden):
    g = cmap_golden
    for k in range(int(g["n_cases"])):
        q, t = g[f"c{k}_q"].tobytes().decode(), g[f"c{k}_t"].tobytes().decode()
        coords = np.ascontiguousarray(g[f"c{k}_coords"])
        thr, gen = g[f"c{k}_thr_gen"]
        thr = int(thr) if k % 3 == 2 else float(thr)
        assert np.array_equal(contact_map_utils.pairwise_sqeuclidean(coords), g[f"c{k}_D"]), k
        sp = bio_utils.calculate_contact_map(coords, thr, mode="sparse")
        assert sp.dtype == np.int32 and np.array_equal(sp, g[f"c{k}_sparse"].reshape(-1, 2)), k
        want = g[f"c{k}_aligned"].astype(np.int32)
        assert np.array_equal(contact_map_utils.align_contact_map(q, t, sp, int(gen)), want), k
        _, fused = bio_utils.build_align_contact_map(Aln(q, t, coords), thr, int(gen))
        assert fused.dtype == np.int32 and np.array_equal(fused, want), k


def test_reference_known_answer_cases():
    # mDeepFRI/tests/test_conctact_map.py:20-41 through the drop-in classes
    ca = contact_map.CAlphaCoordinates("test", np.array([[0, 0, 0], [5, 0, 0], [10, 0, 0]]))
    assert np.array_equal(ca.calculate_contact_map(threshold=6.0).cmap, [[1, 1, 0], [1, 1, 1], [0, 1, 1]])
    dm = contact_map.CAlphaCoordinates("t", np.array([[0, 0, 0], [1, 1, 1]])).calculate_distance_map()
    assert np.allclose(np.sqrt(dm.distance_map), [[0, np.sqrt(3)], [np.sqrt(3), 0]])
    with pytest.raises(ValueError, match="Coordinates are not 3D."):
        contact_map.CAlphaCoordinates("t", np.array([[1, 2], [3, 4]]))
    with pytest.raises(NotImplementedError):
        ca.calculate_distance_map(distance="euclidean")
    assert np.array_equal(ca.calculate_contact_map(6.0).sparsify(), np.argwhere(ca.calculate_contact_map(6.0).cmap == 1))
    # mDeepFRI/tests/test_contact_map_utils.py:16-25 and :99-110
    np.random.seed(42)
    m = np.random.rand(3, 3).astype(np.float32)
    assert np.allclose(contact_map_utils.pairwise_sqeuclidean(m),
                       [[0, 1.01354558, 0.12442072], [1.01354558, 0, 0.99467713], [0.12442072, 0.99467713, 0]])
    N = 100
    tc = np.array([[i, i + 1] for i in range(N - 1)], dtype=np.int32)
    r = contact_map_utils.align_contact_map("A" * N, "A" * N, tc)
    assert r.shape == (N, N) and r[0, 1] == 1 and r[1, 0] == 0      # one-directional, like the .pyx
    assert np.array_equal(contact_map_utils.align_contact_map("AB", "AB", np.array([[0, 1]], np.int32)), [[1, 1], [0, 1]])


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 127, 128, 129, 257, 1000, 2500])
def test_contact_map_sizes(n):
    rng = np.random.default_rng(n)
    X = synth.random_walk_coords(rng, [n])[0]
    assert np.array_equal(contact_map_utils.pairwise_sqeuclidean(X), co.pairwise_sqeuclidean(X))
    for thr in (6.0, 10.0, 6):
        assert np.array_equal(bio_utils.calculate_contact_map(X, thr), co.calculate_contact_map(X, thr))
        sp = bio_utils.calculate_contact_map(X, thr, mode="sparse")
        assert np.array_equal(sp, co.calculate_contact_map(X, thr, mode="sparse"))


def test_contact_map_edge_values():
    # exactly-on-threshold distances, NaN / inf coordinates, zero and negative thresholds
    X = np.array([[0, 0, 0], [6, 0, 0], [0, 5.9999995, 0], [np.nan, 0, 0], [np.inf, 0, 0], [0, 0, 6.0000005]], np.float32)
    for thr in (6, 6.0, 0, 0.0, -1.0, 1e30):
        assert np.array_equal(bio_utils.calculate_contact_map(X, thr), co.calculate_contact_map(X, thr)), thr
        assert np.array_equal(bio_utils.calculate_contact_map(X, thr, mode="sparse"),
                              co.calculate_contact_map(X, thr, mode="sparse")), thr
    D = contact_map_utils.pairwise_sqeuclidean(X)
    assert np.array_equal(D, co.pairwise_sqeuclidean(X), equal_nan=True)
    X5 = np.random.default_rng(0).random((40, 5)).astype(np.float32)     # generic column count
    assert np.array_equal(contact_map_utils.pairwise_sqeuclidean(X5), co.pairwise_sqeuclidean(X5))
    assert contact_map_utils.pairwise_sqeuclidean(np.zeros((0, 3), np.float32)).shape == (0, 0)
    assert bio_utils.calculate_contact_map(np.zeros((0, 3), np.float32), 6.0, mode="sparse").shape == (0, 2)


def test_dtype_strictness_like_typed_memoryviews():
    with pytest.raises(ValueError):
        contact_map_utils.pairwise_sqeuclidean(np.zeros((3, 3), np.float64))
    with pytest.raises(ValueError):
        contact_map_utils.pairwise_sqeuclidean(np.zeros((3, 6), np.float32)[:, ::2])
    with pytest.raises(ValueError):
        contact_map_utils.align_contact_map("AB", "AB", np.array([[0, 1]], np.int64))
    with pytest.raises(KeyError):
        bio_utils.calculate_contact_map(np.zeros((3, 3), np.float32), distance="euclidean")


def test_align_scatter_arbitrary_sparse_inputs():
    rng = np.random.default_rng(3)
    wl = synth.make_workload(30, 1, 300, seed=21, threshold=6.0)
    for i in range(len(wl)):
        nt = len(wl.coords[i])
        sp = rng.integers(-4, nt + 6, size=(int(rng.integers(0, 400)), 2)).astype(np.int32)   # incl. negative / OOR
        for gen in (0, 2, 7):
            got = contact_map_utils.align_contact_map(wl.gapped_query[i], wl.gapped_target[i], sp, gen)
            assert np.array_equal(got, co.align_contact_map(wl.gapped_query[i], wl.gapped_target[i], sp, gen))
    assert contact_map_utils.align_contact_map("---", "ABC", np.zeros((0, 2), np.int32)).shape == (0, 0)
    assert contact_map_utils.align_contact_map("", "", np.zeros((0, 2), np.int32)).shape == (0, 0)


@pytest.mark.parametrize("thr,gen", [(6, 2), (10.0, 2), (10.0, 0), (8.5, 5)])
def test_fused_build_transfer_batch(thr, gen):
    wl = synth.make_workload(120, 1, 700, seed=5, threshold=thr)
    alns = [Aln(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], i) for i in range(len(wl))]
    alns[3].coords = None                                               # bio_utils.py:381-383
    alns[5].coords = alns[5].coords[: len(alns[5].coords) // 2]         # structure shorter than the target sequence
    alns[6].coords = np.concatenate([alns[6].coords, alns[6].coords[:9]])
    alns[7].coords = alns[7].coords.astype(np.float64)                  # pipeline hands float64 sometimes
    dense = bio_utils.build_align_contact_maps(alns, thr, gen)
    packed = bio_utils.build_align_contact_maps(alns, thr, gen, packed=True)
    for i, a in enumerate(alns):
        c = None if a.coords is None else np.ascontiguousarray(a.coords, np.float32)
        want = co.build_align_contact_map(a.gapped_sequence, a.gapped_target, c, thr, gen)
        if want is None:
            assert dense[i] is None and packed[i] is None
            continue
        assert dense[i].dtype == np.int32 and np.array_equal(dense[i], want), i
        assert np.array_equal(batching.unpack_bits(packed[i], want.shape[0]), want), i
    al, cm = bio_utils.build_align_contact_map(alns[10], thr, gen)
    assert al is alns[10] and np.array_equal(cm, dense[10])
    assert bio_utils.build_align_contact_map(alns[3], thr, gen) == (alns[3], None)


@pytest.mark.parametrize("gen", [0, 1, 31, 32, 33, 70, 600])
def test_generated_contact_band_widths(gen):
    """The triangular kernel ORs the diagonal and the generated contacts (contact_map_utils.pyx:82-97) into the 32 x 32 blocks
    next to the diagonal instead of running a second pass: bands narrower than, equal to and wider than a block, wider than the
    whole map, query residues facing target gaps at block borders, lengths around the 32 / 128 boundaries."""
    rng = np.random.default_rng(gen)
    alns = []
    for n, L in enumerate([1, 2, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 160, 257, 300, 511]):
        coords = (rng.normal(size=(L, 3)) * 6).astype(np.float32).round(3)
        q, t, keep = [], [], []
        for i in range(L):
            # target gaps (-> generated residues) in runs, so that generated residues sit on both sides of block borders
            gap = (i // 7 + n) % 5 == 0 or i % 32 in (0, 31) and (i // 32 + n) % 2 == 0
            q.append("A")
            t.append("-" if gap else "G")
            keep.append(not gap)
        alns.append(Aln("".join(q), "".join(t), np.ascontiguousarray(coords[np.array(keep)]) if any(keep) else np.zeros((0, 3), np.float32), n))
    for thr in (6.0, 12):
        packed = bio_utils.build_align_contact_maps(alns, thr, gen, packed=True)
        for a, pk in zip(alns, packed):
            want = co.build_align_contact_map(a.gapped_sequence, a.gapped_target, a.coords, thr, gen)
            assert np.array_equal(batching.unpack_bits(pk, want.shape[0]), want), (a.target_name, thr, gen)


def test_full_size_properties():
    """Config-1 sized inputs (L up to 1000): checked through size-independent properties -
    symmetry (symmetric inputs), unit diagonal, idempotence across calls, and a checksum against
    the oracle on a bounded sample."""
    wl = synth.config_workload(1, 0.004)          # 400 pairs, L 50-1000
    alns = [Aln(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], i) for i in range(len(wl))]
    a = bio_utils.build_align_contact_maps(alns, 6, 2, packed=True)
    b = bio_utils.build_align_contact_maps(alns, 6, 2, packed=True)
    for i, al in enumerate(alns):
        assert np.array_equal(a[i], b[i])
        L = len(al.query_sequence)
        d = batching.unpack_bits(a[i], L)
        assert np.array_equal(d, d.T) and np.all(np.diag(d) == 1)
        assert not a[i].view(np.uint8).reshape(L, -1)[:, (L + 7) // 8 + 1:].any()      # padding bits stay 0
        if i % 16 == 0:
            assert np.array_equal(d, co.build_align_contact_map(al.gapped_sequence, al.gapped_target, al.coords, 6, 2))


def test_build_align_contact_maps_into_caller_arena():
    """`build_align_contact_maps(packed=True, out=arena)` (lists -> `mdf_cmap_build_transfer_ragged`): maps are views of the
    caller's buffer, identical to the oracle; an arena that is too small is refused before any GPU work."""
    import cmap_oracle as co
    from metagenomic_deepfri_b200 import batching, bio_utils, synth

    class Aln:
        pass
    wl = synth.make_workload(60, 1, 420, seed=5, threshold=6.0)
    alns = []
    for i in range(len(wl)):
        a = Aln()
        a.target_name = f"t{i}"
        a.coords, a.gapped_sequence, a.gapped_target = wl.coords[i], wl.gapped_query[i], wl.gapped_target[i]
        alns.append(a)
    alns[4].coords = None
    alns[7].coords = alns[7].coords.astype(np.float64)                 # converted like the reference's astype
    alns[9].coords = alns[9].coords[: max(1, len(alns[9].coords) // 2)]
    words = sum(len(q) * ((len(q) + 127) // 128 * 4) for q in wl.query_seqs)
    arena = np.zeros(words + 7, np.uint32)
    for thr, gen in ((6.0, 2), (10.0, 0)):
        maps = bio_utils.build_align_contact_maps(alns, thr, gen, packed=True, out=arena)
        plain = bio_utils.build_align_contact_maps(alns, thr, gen, packed=True)
        for i, a in enumerate(alns):
            want = co.build_align_contact_map(a.gapped_sequence, a.gapped_target, a.coords, thr, gen)
            if want is None:
                assert maps[i] is None and plain[i] is None
                continue
            assert np.shares_memory(maps[i], arena)
            assert np.array_equal(batching.unpack_bits(maps[i], want.shape[0]), want), i
            assert np.array_equal(maps[i], plain[i])
    with pytest.raises(ValueError):
        bio_utils.build_align_contact_maps(alns, 6.0, 2, packed=True, out=np.zeros(8, np.uint32))
