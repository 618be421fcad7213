"""CPU tests that pin the GCN / CNN oracle: the NumPy ONNX interpreter (`oracle/gcn_oracle.py`)
against a second executor of the same `.onnx` files on PyTorch's CPU kernels (`oracle/torch_ref.py`:
`torch.nn.LSTM`, `F.conv2d`, `F.elu`, `F.batch_norm`, `torch.softmax`), both reading the file through Google's
protobuf runtime (`oracle/onnx_pb.py`) — and that decoder against the product's hand-written one.

Tolerance: 1e-5 absolute on scores (two fp32 implementations of the same graph; the reference's own TF-vs-ORT
check uses atol = 10e-5, weight_convert/random_100_protein_prediction.ipynb cell 1)."""
import numpy as np
import pytest
import torch

import cmap_oracle as co
import gcn_oracle as go
import onnx_pb
import spec
import torch_ref
from conftest import golden_workload
from metagenomic_deepfri_b200 import onnx_lite, synth

TOL = 1e-5


def test_oracle_does_not_import_the_product():
    import os
    here = os.path.dirname(os.path.abspath(go.__file__))
    for f in os.listdir(here):
        if f.endswith(".py"):
            text = open(os.path.join(here, f)).read()
            assert "mdf_pkg" not in text and "import metagenomic" not in text and "from metagenomic" not in text, f


@pytest.mark.parametrize("which", ["gcn", "gcn_tf2onnx", "cnn"])
def test_protobuf_runtime_decoder_agrees_with_product_decoder(which, tmp_path):
    """Same bytes through google.protobuf (oracle) and through the product's wire decoder: identical nodes, attributes
    and initialiser bytes.  Also a float_data / int64_data (non-raw) tensor, which tf2onnx emits for small constants."""
    p = str(tmp_path / "m.onnx")
    if which == "cnn":
        synth.write_cnn_model(p, synth.CNNConfig(filter_lens=(5, 8), num_filters=(128, 128), n_terms=12))
    else:
        cfg = synth.GCNConfig(**spec.SMALL)
        onnx_lite.save(synth.build_gcn_model(cfg, seed=3, style="tf2onnx" if which == "gcn_tf2onnx" else "compact"), p)
    a, b = onnx_pb.load(p), onnx_lite.load(p)
    assert a.opset == b.opset == 15
    assert [v.name for v in a.graph.inputs] == [v.name for v in b.graph.inputs]
    assert [tuple(v.shape) for v in a.graph.inputs] == [tuple(v.shape) for v in b.graph.inputs]
    assert len(a.graph.nodes) == len(b.graph.nodes)
    for x, y in zip(a.graph.nodes, b.graph.nodes):
        assert (x.op_type, x.inputs, x.outputs, x.name) == (y.op_type, y.inputs, y.outputs, y.name)
        assert x.attrs.keys() == y.attrs.keys()
        for k in x.attrs:
            if isinstance(x.attrs[k], np.ndarray):
                assert np.array_equal(x.attrs[k], y.attrs[k]) and x.attrs[k].dtype == y.attrs[k].dtype
            else:
                assert x.attrs[k] == y.attrs[k], (x.op_type, k)
    assert a.graph.initializers.keys() == b.graph.initializers.keys()
    for k, v in a.graph.initializers.items():
        w = b.graph.initializers[k]
        assert v.dtype == w.dtype and v.shape == w.shape and v.tobytes() == w.tobytes(), k


def test_decoders_agree_on_typed_data_fields():
    """Tensors stored in float_data / int64_data / int32_data instead of raw_data (written with the protobuf runtime)."""
    m = onnx_pb.ModelProto()
    m.ir_version = 8
    o = m.opset_import.add()
    o.version = 15
    t = m.graph.initializer.add()
    t.name, t.data_type = "f", 1
    t.dims.extend([2, 3])
    t.float_data.extend([0.5, -1.25, 3.0, 1e-6, 7.0, -0.0])
    t = m.graph.initializer.add()
    t.name, t.data_type = "i", 7
    t.dims.extend([3])
    t.int64_data.extend([-1, 2, 1 << 40])
    t = m.graph.initializer.add()
    t.name, t.data_type = "j", 6
    t.dims.extend([2])
    t.int32_data.extend([-7, 9])
    n = m.graph.node.add()
    n.op_type, n.name = "Identity", "id"
    n.input.append("f")
    n.output.append("y")
    b = onnx_lite.loads(m.SerializeToString())
    assert np.array_equal(b.graph.initializers["f"], np.array([[0.5, -1.25, 3.0], [1e-6, 7.0, -0.0]], np.float32))
    assert np.array_equal(b.graph.initializers["i"], np.array([-1, 2, 1 << 40], np.int64))
    assert np.array_equal(b.graph.initializers["j"], np.array([-7, 9], np.int32))


def test_torch_lstm_gate_permutation_is_not_vacuous(tmp_path):
    """The cross-check must be sensitive to the gate order: feeding torch.nn.LSTM the ONNX rows unpermuted (i,o,f,c read as
    i,f,g,o) changes the output by far more than the tolerance."""
    H, T = 16, 12
    rng = np.random.default_rng(0)
    W = rng.uniform(-1, 1, (1, 4 * H, 26)).astype(np.float32)
    R = rng.uniform(-0.5, 0.5, (1, 4 * H, H)).astype(np.float32)
    B = rng.uniform(-0.5, 0.5, (1, 8 * H)).astype(np.float32)
    X = rng.uniform(-1, 1, (T, 1, 26)).astype(np.float32)
    want = go._lstm(X, W, R, B, H)[0]
    got = torch_ref._onnx_lstm(torch.from_numpy(X), torch.from_numpy(W), torch.from_numpy(R), torch.from_numpy(B), H, None, None,
                               torch.float32)[0].numpy()
    assert np.abs(got - want).max() < 1e-6
    lstm = torch.nn.LSTM(26, H)
    with torch.no_grad():
        lstm.weight_ih_l0.copy_(torch.from_numpy(W[0]))
        lstm.weight_hh_l0.copy_(torch.from_numpy(R[0]))
        lstm.bias_ih_l0.copy_(torch.from_numpy(B[0, :4 * H]))
        lstm.bias_hh_l0.copy_(torch.from_numpy(B[0, 4 * H:]))
        wrong = lstm(torch.from_numpy(X))[0].numpy()
    assert np.abs(wrong - want[:, 0]).max() > 1e-2


@pytest.mark.parametrize("tag", list(spec.GCN_CASES))
def test_gcn_oracle_matches_torch_on_every_golden_case(tag, model_dir, gcn_golden):
    kw, seed, n, lo, hi = spec.GCN_CASES[tag]
    p_np, p_t = go.Predictor(model_dir[tag]), torch_ref.Predictor(model_dir[tag])
    wl = golden_workload(tag)
    worst = 0.0
    for i in range(n):
        cm = co.build_align_contact_map(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], spec.THRESHOLD, spec.GEN)
        y_t = p_t.forward_pass(wl.query_seqs[i], cm)
        want = gcn_golden[tag + "_scores"][i]
        assert y_t.shape == want.shape and y_t.dtype == np.float32
        worst = max(worst, float(np.abs(y_t - want).max()))                  # torch vs the committed golden vector
        if tag != "mf" or i == 0:                                             # NumPy oracle re-run (seconds per protein at full size)
            assert np.abs(p_np.forward_pass(wl.query_seqs[i], cm) - y_t).max() <= TOL
    assert worst <= TOL, f"{tag}: torch-CPU vs golden {worst:.2e}"


def test_intermediate_tensors_agree(model_dir):
    """Not only the scores: LSTM outputs, embedding, normalised degrees, last GraphConv output and the pooled vector."""
    path = model_dir["small"]
    a, b = go.OnnxOracle(path), torch_ref.TorchOnnx(path)
    wl = golden_workload("small")
    cm = co.build_align_contact_map(wl.gapped_query[1], wl.gapped_target[1], wl.coords[1], spec.THRESHOLD, spec.GEN)
    feeds = {"cmap": cm[None].astype(np.float32), "seq": co.seq2onehot(wl.query_seqs[1])[None]}
    names = ["lm/LSTM1_out", "lm/LSTM2_bm", "activation/Relu", "norm/d", "GraphConv_2/Elu", "SumPooling/Sum", "labels"]
    for name, x, y in zip(names, a.run(names, feeds), b.run(names, feeds)):
        assert x.shape == y.shape, name
        scale = max(1.0, float(np.abs(x).max()))
        assert np.abs(x - y).max() <= 2e-6 * scale * (50 if name == "SumPooling/Sum" else 1), name


@pytest.mark.parametrize("tag", list(spec.CNN_CASES))
def test_cnn_oracle_matches_torch_on_every_golden_case(tag, cnn_model_dir, cnn_golden):
    p_np, p_t = go.Predictor(cnn_model_dir[tag]), torch_ref.Predictor(cnn_model_dir[tag])
    seqs = spec.cnn_sequences(tag)
    for i, s in enumerate(seqs):
        y_t = p_t.forward_pass(s)
        assert np.abs(y_t - cnn_golden[tag + "_scores"][i]).max() <= TOL
        if i in spec.CNN_GOLDEN_CHECK[tag]:
            assert np.abs(p_np.forward_pass(s) - y_t).max() <= TOL


def notebook_case(rng, lo=60, hi=1000):
    """weight_convert/random_100_protein_prediction.ipynb cell 1: random length in [60, 1000), random sequence, random
    NON-symmetric 0/1 contact map (np.random.randint(0, 2, (L, L)))."""
    L = int(rng.integers(lo, hi))
    seq = "".join(rng.choice(list(synth.AA20), L))
    return seq, rng.integers(0, 2, (L, L)).astype(np.int32)


def test_notebook_harness_random_nonsymmetric_maps(model_dir):
    """The reference's only numerical check of the GCN (TF vs ORT, atol 10e-5) restated with the two CPU executors:
    dense random maps are the worst case for the degree normalisation and the adjacency product."""
    rng = np.random.default_rng(2024)
    p_np, p_t = go.Predictor(model_dir["small"]), torch_ref.Predictor(model_dir["small"])
    for _ in range(6):
        seq, cm = notebook_case(rng, 60, 400)
        y, z = p_np.forward_pass(seq, cm), p_t.forward_pass(seq, cm)
        assert np.isfinite(y).all() and np.abs(y - z).max() <= TOL
    p_np, p_t = go.Predictor(model_dir["mf"]), torch_ref.Predictor(model_dir["mf"])
    seq, cm = notebook_case(rng, 60, 200)
    assert np.abs(p_np.forward_pass(seq, cm) - p_t.forward_pass(seq, cm)).max() <= TOL


def test_tf2onnx_style_graph_same_scores(tmp_path):
    """The tf2onnx-faithful lowering (per-layer matmul(matmul(D, A_hat), D) with D = diag(d), LSTM nodes carrying zero
    initial states and a full-length sequence_lens) computes the same function as the compact lowering."""
    cfg = synth.GCNConfig(**spec.SMALL)
    w = synth.make_weights(cfg, seed=9)
    pa, pb = str(tmp_path / "a.onnx"), str(tmp_path / "b.onnx")
    onnx_lite.save(synth.build_gcn_model(cfg, w, style="compact"), pa)
    onnx_lite.save(synth.build_gcn_model(cfg, w, style="tf2onnx"), pb)
    wl = synth.make_workload(3, 20, 120, seed=5, threshold=10.0)
    for i in range(len(wl)):
        cm = co.build_align_contact_map(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], 10.0, 2)
        ya = go.Predictor(pa).forward_pass(wl.query_seqs[i], cm)
        yb = go.Predictor(pb).forward_pass(wl.query_seqs[i], cm)
        yt = torch_ref.Predictor(pb).forward_pass(wl.query_seqs[i], cm)
        assert np.abs(ya - yb).max() <= TOL and np.abs(yb - yt).max() <= TOL
