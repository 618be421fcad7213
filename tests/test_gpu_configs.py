"""GPU parity tests at the shapes of BASELINE.json configs[2..4]: four heads on a metagenomic length
distribution, long proteins (L 1000-2500), and batch-scale properties of the sharded 1M-protein job.

Oracle comparisons run on sizes the NumPy interpreter finishes in seconds; at the full batch size the
checks are size-independent properties (batch-composition invariance, agreement with the exact-fp32
SIMT engine on a sample, finiteness) - the SIMT engine itself is pinned to the oracle at 2e-5 in
test_gpu_gcn.py."""
import numpy as np
import pytest

import cmap_oracle as co
import gcn_oracle as go
from metagenomic_deepfri_b200 import predict, synth

pytestmark = pytest.mark.gpu
TOL = 1e-3                      # north-star tolerance on GO-term scores
HEADS = {"mf": 489, "bp": 1943, "cc": 320, "ec": 538}


def assert_scores(got, want, tol=TOL):
    err = float(np.abs(got - want).max())
    assert err <= tol, f"max |score - reference| = {err:.3e} > {tol}"
    clear = np.abs(want - 0.1) > tol
    assert np.array_equal((got >= 0.1)[clear], (want >= 0.1)[clear]), "GO calls differ outside the guard band"


def oracle_scores(path, wl, idx):
    oracle = go.Predictor(path)
    out = []
    for i in idx:
        cm = co.build_align_contact_map(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], wl.threshold, wl.generated_contacts)
        out.append(oracle.forward_pass(wl.query_seqs[i], cm))
    return np.stack(out)


def take(wl, idx):
    return ([wl.query_seqs[i] for i in idx], [wl.gapped_query[i] for i in idx], [wl.gapped_target[i] for i in idx],
            [wl.coords[i] for i in idx])


@pytest.mark.parametrize("head", list(HEADS))
def test_config2_four_heads_against_oracle(head, tmp_path):
    """configs[2]: MF / BP / CC / EC heads on a LogNormal(250, 0.6) length sample."""
    path = str(tmp_path / f"{head}.onnx")
    synth.write_gcn_model(path, synth.GCNConfig(n_terms=HEADS[head]), seed=1234 + len(head))
    pred = predict.Predictor(path)
    pred.set_engine("tc")
    wl = synth.config_workload(2, 0.0012)          # 12 proteins
    idx = list(range(len(wl)))
    got = pred.forward_structures(*take(wl, idx), threshold=wl.threshold, generated_contacts=wl.generated_contacts)
    assert got.shape == (len(wl), HEADS[head])
    assert_scores(got, oracle_scores(path, wl, idx))
    pred.close()


def test_config3_long_proteins_against_oracle(tmp_path):
    """configs[3]: L 1000-2500 (dense A tiles, LSTM-LM at full length).  The score error of fp16 activations grows
    with L (the logits scale with the pooled sum); the 8-phase dithered LSTM weights keep it inside the tolerance."""
    path = str(tmp_path / "mf.onnx")
    synth.write_gcn_model(path, synth.GCNConfig())
    pred = predict.Predictor(path)
    wl = synth.config_workload(3, 0.003)           # 6 proteins, L 1000-2500
    idx = list(range(len(wl)))
    want = oracle_scores(path, wl, idx)
    for engine, tol in (("simt", 5e-5), ("tc", TOL)):
        pred.set_engine(engine)
        got = pred.forward_structures(*take(wl, idx), threshold=wl.threshold, generated_contacts=wl.generated_contacts)
        assert_scores(got, want, tol)
    pred.close()


def test_config4_batch_scale_properties(tmp_path):
    """configs[4] at the bench's per-GPU batch size: every protein of a 16,384-protein batch gets the score it
    gets alone or in a small batch (no cross-protein leakage through the packed LSTM sub-batches, grouped GEMM
    tiles or pooling), and a sample agrees with the exact-fp32 engine inside the tolerance."""
    path = str(tmp_path / "mf.onnx")
    synth.write_gcn_model(path, synth.GCNConfig())
    pred = predict.Predictor(path)
    pred.set_engine("tc")
    wl = synth.make_workload(16384, 50, 1000, seed=5, dist="lognormal", threshold=10.0)
    batch = pred.upload(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)
    pred.run(batch, wl.threshold, wl.generated_contacts)
    full = pred.fetch_scores(batch)
    batch.close()
    assert np.isfinite(full).all() and full.min() >= 0.0 and full.max() <= 1.0
    idx = np.linspace(0, len(wl) - 1, 192).astype(int)
    sub = pred.upload(*take(wl, idx))
    pred.run(sub, wl.threshold, wl.generated_contacts)
    small = pred.fetch_scores(sub)
    assert np.abs(small - full[idx]).max() < 2e-5          # fp32 atomics in the sum-pool are the only order dependence
    pred.set_engine("simt")
    pred.run(sub, wl.threshold, wl.generated_contacts)
    assert_scores(small, pred.fetch_scores(sub))
    sub.close()
    pred.close()


def test_tc_engine_batch_invariance_and_order(tmp_path):
    path = str(tmp_path / "mf.onnx")
    synth.write_gcn_model(path, synth.GCNConfig())
    pred = predict.Predictor(path)
    pred.set_engine("tc")
    wl = synth.make_workload(300, 1, 400, seed=78, threshold=10.0)
    full = pred.forward_structures(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords, 10.0, 2)
    perm = np.random.default_rng(1).permutation(len(wl))
    shuffled = pred.forward_structures(*take(wl, perm), threshold=10.0, generated_contacts=2)
    assert np.abs(shuffled - full[perm]).max() < 2e-5
    one = pred.forward_structures(*take(wl, [7]), threshold=10.0, generated_contacts=2)
    assert np.abs(one[0] - full[7]).max() < 2e-5
    pred.close()


def test_pipeline_adapter_shares_maps_and_lm_across_heads(tmp_path):
    """`pipeline.predict_structures`: one upload, one contact-map build and one LSTM-LM run serve all four heads; the
    scores equal what each head computes on its own, rows come out like pipeline.py:318-319 writes them."""
    import csv
    import io
    from conftest import Aln
    from metagenomic_deepfri_b200 import pipeline
    preds, paths = {}, {}
    for head, C in HEADS.items():
        paths[head] = str(tmp_path / f"{head}.onnx")
        synth.write_gcn_model(paths[head], synth.GCNConfig(n_terms=C), seed=77 + C)      # same LM (lm_seed), own heads
        preds[head] = predict.Predictor(paths[head])
        preds[head].set_engine("tc")
    wl = synth.config_workload(2, 0.004)           # 40 proteins, LogNormal lengths
    alns = [Aln(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], i) for i in range(len(wl))]
    alns[5].coords = None                          # no structure: reported back, not predicted
    sinks = {h: io.StringIO() for h in HEADS}
    writers = {h: csv.writer(sinks[h], delimiter="\t") for h in HEADS}
    scores, kept, skipped = pipeline.predict_structures(preds, alns, wl.threshold, wl.generated_contacts, max_residues=4000,
                                                        writers=writers)
    assert skipped == [5] and kept == [i for i in range(len(wl)) if i != 5]
    for head, pred in preds.items():
        alone = pred.forward_structures(*take(wl, kept), threshold=wl.threshold, generated_contacts=wl.generated_contacts)
        assert scores[head].shape == (len(kept), HEADS[head])
        assert np.abs(scores[head] - alone).max() < 2e-5
        rows = list(csv.reader(io.StringIO(sinks[head].getvalue()), delimiter="\t"))
        assert [r[0] for r in rows] == [alns[i].query_name for i in kept] and all(r[1] == "gcn" for r in rows)
        assert np.allclose(np.array(rows[3][2:], np.float64), scores[head][3], atol=1e-7)
    assert_scores(scores["mf"][:4], oracle_scores(paths["mf"], wl, kept[:4]))
    # the reference helper's signature on precomputed maps
    pairs = [(alns[i], co.build_align_contact_map(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], wl.threshold,
                                                  wl.generated_contacts)) for i in kept[:6]]
    sink = io.StringIO()
    pipeline.run_prediction_loop(preds["cc"], pairs, len(pairs), "gcn", csv.writer(sink, delimiter="\t"), "cc")
    rows = list(csv.reader(io.StringIO(sink.getvalue()), delimiter="\t"))
    assert len(rows) == 6 and np.abs(np.array([r[2:] for r in rows], np.float64) - scores["cc"][:6]).max() < 2e-5
    for p in preds.values():
        p.close()


def test_block_sparse_adjacency_equals_dense_walk(tmp_path, monkeypatch):
    """The adjacency GEMM skips all-zero 128 x 64 tiles of A_hat (tc_engine.cu adj_tile_scan_kernel).  Skipped tiles
    contribute exact zeros, so the per-residue GraphConv output must equal the dense walk's bit for bit; the pooled sums
    (fp32 atomics, order not fixed) and the scores agree to rounding.  Long proteins: most of their off-diagonal tiles are
    empty, and lengths around the 64 / 128 / 256 boundaries exercise the list edges."""
    path = str(tmp_path / "mf.onnx")
    synth.write_gcn_model(path, synth.GCNConfig())
    wl_a = synth.make_workload(6, 900, 1400, seed=31, threshold=10.0)
    wl_b = synth.make_workload(10, 60, 260, seed=32, threshold=6.0)
    seqs = wl_a.query_seqs + wl_b.query_seqs
    gq, gt, co_ = wl_a.gapped_query + wl_b.gapped_query, wl_a.gapped_target + wl_b.gapped_target, wl_a.coords + wl_b.coords
    out = {}
    # two weight terms in every layer: with the mean-corrected single term the input of layers >= 2 depends on the previous
    # layer's pooled sums (fp32 atomics, order not fixed), so only the two-term path is reproducible bit for bit
    monkeypatch.setenv("MDF_XW_MEAN", "0")
    for mode in ("1", "0"):
        monkeypatch.setenv("MDF_ADJ_SPARSE", mode)         # read when the model is created
        pred = predict.Predictor(path)
        pred.set_engine("tc")
        pred._ctx.set_debug_taps(True)
        try:
            batch = pred.upload(seqs, gq, gt, co_)
            pred.run(batch, 10.0, 2)
            out[mode] = (pred.fetch(batch, "gc_last"), pred.fetch(batch, "pooled"), pred.fetch_scores(batch))
            batch.close()
        finally:
            pred._ctx.set_debug_taps(False)
            pred.close()
    assert np.array_equal(out["1"][0], out["0"][0])
    assert np.allclose(out["1"][1], out["0"][1], rtol=1e-5, atol=1e-4)
    assert np.abs(out["1"][2] - out["0"][2]).max() < 1e-5


def test_compact_axis_many_tiny_proteins(tmp_path):
    """On the compact residue axis one 64-residue k-block of Y^T can hold dozens of proteins: a batch of 300 proteins of 1-25
    residues (plus a few long ones in between) through the tensor-core engine must agree with the exact-fp32 SIMT engine, and
    with itself when the padded axis is selected."""
    import os
    path = str(tmp_path / "mf.onnx")
    synth.write_gcn_model(path, synth.GCNConfig())
    tiny = synth.make_workload(300, 1, 25, seed=41, threshold=10.0)
    big = synth.make_workload(5, 200, 700, seed=42, threshold=10.0)
    seqs, gq, gt, co_ = list(tiny.query_seqs), list(tiny.gapped_query), list(tiny.gapped_target), list(tiny.coords)
    for k in range(len(big)):                                   # long proteins interleaved at positions 0, 61, 122, ...
        at = 61 * k
        seqs.insert(at, big.query_seqs[k]); gq.insert(at, big.gapped_query[k]); gt.insert(at, big.gapped_target[k]); co_.insert(at, big.coords[k])
    pred = predict.Predictor(path)
    pred.set_engine("simt")
    want = pred.forward_structures(seqs, gq, gt, co_, threshold=10.0, generated_contacts=2)
    pred.set_engine("tc")
    got = pred.forward_structures(seqs, gq, gt, co_, threshold=10.0, generated_contacts=2)
    pred.close()
    assert_scores(got, want)
    # the two axis layouts with the same arithmetic (two weight terms in every layer: the mean-corrected single term of the
    # default path needs the compact axis) agree to the order of the pool's fp32 atomics
    res = {}
    for compact_axis in ("1", "0"):
        os.environ["MDF_COMPACT"] = compact_axis
        os.environ["MDF_XW_MEAN"] = "0"
        try:
            p2 = predict.Predictor(path)
            p2.set_engine("tc")
            res[compact_axis] = p2.forward_structures(seqs, gq, gt, co_, threshold=10.0, generated_contacts=2)
            p2.close()
        finally:
            del os.environ["MDF_COMPACT"], os.environ["MDF_XW_MEAN"]
    assert np.abs(res["1"] - res["0"]).max() < 2e-5
    assert_scores(res["1"], want)
    assert np.abs(got - res["1"]).max() < 5e-4               # single term + mean correction against two terms
