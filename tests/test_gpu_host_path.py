"""GPU tests of the host-facing half of the drop-in (round 2): the `.onnx` loader behind the C ABI, the asynchronous
submit / wait jobs with ragged Python-list inputs, the reference-layout dense map output, and the reference's own
numerical harness (weight_convert/random_100_protein_prediction.ipynb cell 1: random NON-symmetric 0/1 maps, L in [60, 1000))
against the torch-CPU executor."""
import ctypes as C

import numpy as np
import pytest

import cmap_oracle as co
import gcn_oracle as go
import spec
import torch_ref
from conftest import golden_workload
from metagenomic_deepfri_b200 import _lib, batching, onnx_lite, predict, synth

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def mf(model_dir):
    p = predict.Predictor(model_dir["mf"])
    yield p
    p.close()


def test_loader_styles_give_identical_scores(tmp_path):
    """Same weights lowered two ways (shared broadcast-Mul normalisation vs per-layer matmul(matmul(D, A), D), LSTM nodes with
    zero initial states): the C loader recognises both and the pipeline computes the same scores (up to the order of the
    floating-point atomics of the fused sum-pool, which differs from run to run)."""
    cfg = synth.GCNConfig(**spec.GCN_CASES["tc_small"][0])
    w = synth.make_weights(cfg, seed=8)
    pa, pb = str(tmp_path / "a.onnx"), str(tmp_path / "b.onnx")
    onnx_lite.save(synth.build_gcn_model(cfg, w, style="compact"), pa)
    onnx_lite.save(synth.build_gcn_model(cfg, w, style="tf2onnx"), pb)
    wl = golden_workload("tc_small")
    ya = predict.Predictor(pa).forward_structures(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords, 10.0, 2)
    yb = predict.Predictor(pb).forward_structures(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords, 10.0, 2)
    assert np.abs(ya - yb).max() < 1e-5
    want = np.stack([go.Predictor(pb).forward_pass(s, co.build_align_contact_map(q, t, c, 10.0, 2))
                     for s, q, t, c in zip(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)])
    assert np.abs(yb - want).max() <= TOL


def test_model_info_and_errors(mf, tmp_path):
    n_terms, H, E, n_gc, F = (C.c_int() for _ in range(5))
    dims = (C.c_int * 8)()
    _lib.check(_lib.lib().mdf_model_info(mf._handle, C.byref(n_terms), C.byref(H), C.byref(E), C.byref(n_gc), dims, C.byref(F)))
    assert (n_terms.value, H.value, E.value, n_gc.value, F.value, list(dims[:3])) == (489, 512, 1024, 3, 1024, [512, 512, 512])
    with pytest.raises(FileNotFoundError):
        predict.Predictor(str(tmp_path / "missing.onnx"))
    bad = tmp_path / "bad.onnx"
    bad.write_bytes(b"\x08\x08not an onnx file")
    with pytest.raises(RuntimeError):
        predict.Predictor(str(bad))
    m = synth.build_gcn_model(synth.GCNConfig(**spec.SMALL))
    for n in m.graph.nodes:                                   # A_hat = A + I: not the function the pipeline computes
        if n.op_type == "Add" and n.outputs[0].endswith("A_hat"):
            n.inputs[0] = "cmap"
    wrong = tmp_path / "wrong_norm.onnx"
    onnx_lite.save(m, str(wrong))
    with pytest.raises(_lib.UnsupportedModelError):
        predict.Predictor(str(wrong))


def test_submit_wait_jobs_overlap_and_match(mf):
    """Three chunks through submit / wait with two jobs in flight == the same chunks through the resident upload / run path;
    inputs are plain Python lists (float64 and list-of-lists structures are converted like the reference's astype)."""
    wl = synth.keyed_workload(np.arange(900), 5)
    parts = [range(0, 300), range(300, 600), range(600, 900)]
    sub = lambda xs, r: [xs[i] for i in r]
    want = []
    for r in parts:
        b = mf.upload(sub(wl.query_seqs, r), sub(wl.gapped_query, r), sub(wl.gapped_target, r), sub(wl.coords, r))
        mf.run(b, 10.0, 2)
        want.append(mf.fetch_scores(b))
        b.close()
    out = np.zeros((900, mf.n_terms), np.float32)
    jobs = []
    for k, r in enumerate(parts):
        coords = sub(wl.coords, r)
        if k == 1:
            coords = [c.astype(np.float64) for c in coords]          # not float32: converted on the retry path
        if k == 2:
            coords[0] = coords[0].tolist()
        jobs.append(mf.submit_structures(sub(wl.query_seqs, r), sub(wl.gapped_query, r), sub(wl.gapped_target, r), coords, 10.0, 2,
                                         out=out[r.start:r.stop]))
        if len(jobs) == 2:
            with pytest.raises(ValueError, match="both job slots"):
                mf.submit_structures(["ACD"], ["ACD"], ["ACD"], [np.zeros((3, 3), np.float32)])
            jobs.pop(0).wait()
    for j in jobs:
        j.wait()
    for r, w in zip(parts, want):
        assert np.abs(out[r.start:r.stop] - w).max() < 1e-5
    # device-side input errors surface at wait(), named like the reference names them
    with pytest.raises(ValueError, match="Invalid character in sequence: J"):
        mf.forward_structures(["ACJE"], ["ACJE"], ["ACDE"], [np.zeros((4, 3), np.float32)])
    with pytest.raises(ValueError, match="do not match"):
        mf.forward_structures(["ACDE"], ["ACD-"], ["ACDE"], [np.zeros((4, 3), np.float32)])
    with pytest.raises(UnicodeEncodeError):
        mf.forward_structures(["ACé"], ["ACé"], ["ACD"], [np.zeros((3, 3), np.float32)])
    with pytest.raises(ValueError, match="differ in length"):
        mf.forward_structures(["ACD"], ["ACD"], ["AC"], [np.zeros((2, 3), np.float32)])
    assert mf.forward_structures([], [], [], []).shape == (0, mf.n_terms)
    # the context is still healthy after the failed calls
    r = parts[0]
    y = mf.forward_structures(sub(wl.query_seqs, r), sub(wl.gapped_query, r), sub(wl.gapped_target, r), sub(wl.coords, r), 10.0, 2)
    assert np.abs(y - want[0]).max() < 1e-5


def test_flat_submit_through_the_c_abi(mf):
    """mdf_path_submit with flat host buffers (what a non-Python host passes), two jobs in flight."""
    wl = synth.make_workload(64, 30, 200, seed=3)
    L = _lib.lib()
    outs, jobs, keep = [], [], []
    for half in (range(0, 32), range(32, 64)):
        seqs = [wl.query_seqs[i] for i in half]
        ps = batching.pack_structures([wl.gapped_query[i] for i in half], [wl.gapped_target[i] for i in half], [wl.coords[i] for i in half])
        sb, so = predict._pack_checked(seqs)
        out = np.empty((32, mf.n_terms), np.float32)
        job = C.c_void_p()
        _lib.check(L.mdf_path_submit(mf._handle, 32, sb, _lib.lp(so), ps.coords.ctypes.data, _lib.lp(ps.coord_off), ps.q_aln, ps.t_aln,
                                     _lib.lp(ps.aln_off), float(np.float32(100.0)), 2, out.ctypes.data, C.byref(job)))
        outs.append(out); jobs.append(job); keep.append((ps, sb, so))
    for j in jobs:
        _lib.check(L.mdf_path_wait(j))
    with pytest.raises(ValueError, match="already been waited"):
        _lib.check(L.mdf_path_wait(jobs[0]))
    want = mf.forward_structures(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords, 10.0, 2)
    assert np.abs(np.concatenate(outs) - want).max() < 1e-5


def test_dense_reference_layout_output(mf):
    """mdf_batch_unpack_dense writes what build_align_contact_map returns (int32 [Lq, Lq] per protein), bit for bit."""
    import torch
    wl = synth.make_workload(40, 1, 300, seed=12, threshold=6.0)
    b = mf.upload(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)
    mf.run(b, 6.0, 2, upto=1)
    lens = np.array([len(s) for s in wl.query_seqs], np.int64)
    off = np.concatenate([[0], np.cumsum(lens * lens)])
    dev = torch.full((int(off[-1]) + 3,), -7, dtype=torch.int32, device="cuda")
    cells = C.c_int64()
    for shift in (0, 1, 3):                                    # every alignment of the block start against the 16-byte stores
        dev.fill_(-7)
        _lib.check(_lib.lib().mdf_batch_unpack_dense(b.handle, C.c_void_p(dev.data_ptr() + 4 * shift), C.byref(cells)))
        torch.cuda.synchronize()
        host = dev.cpu().numpy()
        assert cells.value == off[-1] and np.all(host[:shift] == -7) and np.all(host[shift + off[-1]:] == -7)
        for i in range(len(wl)):
            want = co.build_align_contact_map(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], 6.0, 2)
            got = host[shift + off[i]:shift + off[i + 1]].reshape(lens[i], lens[i])
            assert np.array_equal(got, want), (shift, i)
    assert mf.unpack_dense(b) == off[-1]
    b.close()


def test_notebook_harness_nonsymmetric_maps(mf, model_dir):
    """weight_convert/random_100_protein_prediction.ipynb cell 1 restated: random length in [60, 1000), random sequence, random
    NON-symmetric 0/1 map; the CUDA path against the torch-CPU executor of the same file.  Dense random maps (half of all
    entries set) are far denser than any protein's: the degree normalisation and the adjacency product at their extreme."""
    rng = np.random.default_rng(7)
    cpu = torch_ref.Predictor(model_dir["mf"])
    worst, flips, in_band = 0.0, 0, 0
    for k in range(10):
        L = int(rng.integers(60, 1000))
        seq = "".join(rng.choice(list(synth.AA20), L))
        cm = rng.integers(0, 2, (L, L)).astype(np.int32)
        want = cpu.forward_pass(seq, cm)
        got = mf.forward_pass(seq, cm)
        assert np.isfinite(got).all()
        worst = max(worst, float(np.abs(got - want).max()))
        band = np.abs(want - 0.1) <= TOL
        in_band += int(band.sum())
        flips += int(((got >= 0.1) != (want >= 0.1))[band].sum())
        assert np.array_equal((got >= 0.1)[~band], (want >= 0.1)[~band])
    print(f"notebook harness: max |score - torch-CPU| = {worst:.2e}; {in_band} scores within 1e-3 of the 0.1 threshold, {flips} of them flip")
    assert worst <= TOL


def test_calpha_cache_feeds_the_path(mf, tmp_path):
    """§8f row 3: structures parsed from PDB text once, cached, and handed to the batched path as views into the mapping give the
    same scores as the coordinate arrays themselves."""
    from metagenomic_deepfri_b200 import ingest
    wl = synth.keyed_workload(np.arange(200, 328), 5)
    rng = np.random.default_rng(1)
    targets = [t.replace("-", "") for t in wl.gapped_target]
    texts = [synth.pdb_text(t, c, rng, hetatm=(i % 3 == 0), extra_chain=(i % 5 == 0)) for i, (t, c) in enumerate(zip(targets, wl.coords))]
    ids = [f"target_{i}" for i in range(len(wl))]
    path = str(tmp_path / "targets.mdfca")
    assert ingest.write_cache_from_pdb_texts(path, ids, texts, "A", threads=4) == []
    cache = ingest.CoordsCache(path)
    views = cache.get(ids)
    for v, c in zip(views, wl.coords):
        assert np.array_equal(v, c)                       # keyed coordinates carry 3 decimals: the PDB text is lossless for them
    want = mf.forward_structures(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords, 10.0, 2)
    got = mf.forward_structures(wl.query_seqs, wl.gapped_query, wl.gapped_target, views, 10.0, 2)
    assert np.abs(got - want).max() < 1e-5
    cache.close()
