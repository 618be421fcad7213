"""GPU parity of the sequence-only DeepCNN branch (predict.pyx:91-95, csrc/cnn_tc.cu) through the C ABI against the
oracle's golden vectors.  Tolerance: GO-term scores <= 1e-3 absolute (north-star), identical calls at 0.1 outside a
+-1e-3 guard band; the max-pooled features are compared at 2e-3 relative to their scale (fp16 conv weights)."""
import csv

import numpy as np
import pytest

import gcn_oracle as go
import spec
from metagenomic_deepfri_b200 import onnx_lite as ox
from metagenomic_deepfri_b200 import predict, synth

pytestmark = pytest.mark.gpu
TOL = 1e-3


@pytest.fixture(scope="module")
def cnn(cnn_model_dir):
    return {tag: predict.Predictor(path) for tag, path in cnn_model_dir.items()}


@pytest.mark.parametrize("tag", list(spec.CNN_CASES))
def test_scores_match_golden(cnn, cnn_golden, tag):
    pred = cnn[tag]
    seqs = spec.cnn_sequences(tag)
    want = cnn_golden[f"{tag}_scores"]
    got = pred.forward_sequences(seqs)
    assert got.dtype == np.float32 and got.shape == want.shape
    err = np.abs(got - want)
    assert err.max() <= TOL, f"{tag}: max |score - oracle| = {err.max():.3e}"
    clear = np.abs(want - 0.1) > TOL
    assert np.array_equal((got >= 0.1)[clear], (want >= 0.1)[clear])
    # forward_pass(seqres) = batch of one through the same kernels (predict.pyx:91-95)
    for i in (0, len(seqs) - 1):
        one = pred.forward_pass(seqs[i])
        assert one.shape == (pred.n_terms,) and np.abs(one - want[i]).max() <= TOL


def test_pooled_features_and_batch_invariance(cnn, cnn_model_dir):
    pred = cnn["cnn_small"]
    rng = np.random.default_rng(9)
    letters = np.array(list(spec.CNN_ALPHABET))
    lens = [1, 2, 127, 128, 129, 255, 256, 257, 511, 513, 700, 64, 40, 300, 17]
    seqs = ["".join(rng.choice(letters, L)) for L in lens]
    pred.upload_sequences(seqs)
    pred.run_sequences()
    scores, pooled = pred.fetch_sequences(pooled=True)
    # the oracle's max-pooled features: run the graph up to the ReduceMax
    orc = go.OnnxOracle(cnn_model_dir["cnn_small"])
    for i in (0, 2, 3, 4, 7, 9, 10):
        S = go.seq2onehot(seqs[i])[None].astype(np.float32)
        want_pool, want = orc.run(["global_max_pooling1d/Max", "labels"], {"seq": S})
        assert np.abs(pooled[i] - want_pool[0]).max() <= 2e-3 * max(1.0, float(np.abs(want_pool).max()))
        assert np.abs(scores[i] - want[0, :, 0]).max() <= TOL
    # scores do not depend on the rest of the batch or on the position in it
    perm = rng.permutation(len(seqs))
    shuffled = pred.forward_sequences([seqs[i] for i in perm])
    assert np.array_equal(shuffled, scores[perm])
    assert np.array_equal(pred.forward_sequences(seqs[4:5])[0], scores[4])
    assert pred.forward_sequences([]).shape == (0, pred.n_terms)


def test_contract_and_errors(cnn, model_dir):
    pred = cnn["cnn_small"]
    assert pred.is_cnn and pred.input_names == ["seq"] and [a.name for a in pred.session.get_inputs()] == ["seq"]
    with pytest.raises(ValueError, match="Invalid character in sequence: J"):
        pred.forward_pass("ACJE")
    with pytest.raises(ValueError):
        pred.forward_pass("")                                 # ReduceMax over no residues
    with pytest.raises(ValueError):
        pred.forward_sequences(["ACD", ""])
    with pytest.raises(ValueError, match="one input"):
        pred.forward_pass("ACDE", np.eye(4, dtype=np.int32))  # a CNN model has no contact-map input
    with pytest.raises(ValueError):
        predict.Predictor(model_dir["small"]).forward_sequences(["ACDE"])
    # unsupported shapes fail loudly at load time (no fallback executor)
    import tempfile, os
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "odd.onnx")
        synth.write_cnn_model(p, synth.CNNConfig(filter_lens=(8,), num_filters=(100,), n_terms=5))
        with pytest.raises(NotImplementedError):
            predict.Predictor(p)


def test_pipeline_loop_drop_in_cnn(cnn, tmp_path):
    """The reference's prediction loop for unaligned queries (pipeline.py:313-319, :606-613) over the drop-in Predictor,
    and the batched adapter: same rows."""
    from metagenomic_deepfri_b200 import pipeline
    pred = cnn["cnn_small"]
    unaligned = {f"q{i}": s for i, s in enumerate(spec.cnn_sequences("cnn_small"))}
    ref_rows = []
    for query_id, sequence in unaligned.items():               # pipeline.py:313-319
        ref_rows.append([query_id, "cnn"] + pred.forward_pass(seqres=sequence).tolist())
    out = tmp_path / "prediction_matrix_mf_cnn.tsv"
    with open(out, "w", newline="") as fh:
        pipeline.run_prediction_loop(pred, unaligned.items(), len(unaligned), "cnn", csv.writer(fh, delimiter="\t"))
    rows = list(csv.reader(open(out), delimiter="\t"))
    assert [r[:2] for r in rows] == [r[:2] for r in ref_rows]
    assert np.array_equal(np.array([r[2:] for r in rows], np.float32), np.array([r[2:] for r in ref_rows], np.float32))


def test_metagenomic_batch_properties(cnn, cnn_model_dir):
    """A 2,048-sequence batch of the configs[4] length distribution through the trained models' shape (16 x 512 filters,
    widths 8..128): a bounded sample against the oracle, and size-independent properties of the whole batch - every score in
    (0, 1), identical sequences give identical scores wherever they sit in the batch, and a sequence scored alone equals its
    score inside the batch (the max-pool is order-free, so this is exact)."""
    pred = cnn["cnn_mf"]
    rng = np.random.default_rng(77)
    lengths = np.clip(np.exp(rng.normal(np.log(250.0), 0.6, size=2048)), 50, 1000).astype(np.int64)
    seqs = synth.random_sequences(rng, lengths)
    seqs[5] = seqs[1999]                                   # the same protein twice, far apart in the batch
    got = pred.forward_sequences(seqs)
    assert got.shape == (2048, 489) and np.isfinite(got).all() and (got > 0).all() and (got < 1).all()
    assert np.array_equal(got[5], got[1999])
    assert np.array_equal(pred.forward_sequences([seqs[1234]])[0], got[1234])
    orc = go.Predictor(cnn_model_dir["cnn_mf"])
    for i in (0, 777, 2047, int(np.argmax(lengths)), int(np.argmin(lengths))):
        assert np.abs(orc.forward_pass(seqs[i]) - got[i]).max() <= TOL
