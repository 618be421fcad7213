"""CPU tests: ONNX wire format, graph recognition, and the GCN oracle against an independent
NumPy restatement of upstream DeepFRI's layer equations."""
import numpy as np
import pytest

import cmap_oracle as co
import gcn_oracle as go
from conftest import golden_workload
from metagenomic_deepfri_b200 import onnx_lite as ox
from metagenomic_deepfri_b200 import onnx_plan, synth
import spec


def test_wire_roundtrip():
    m = synth.build_gcn_model(synth.GCNConfig(**spec.SMALL), seed=3)
    m2 = ox.loads(ox.dumps(m))
    assert [(n.op_type, n.inputs, n.outputs, n.attrs) for n in m2.graph.nodes] == \
           [(n.op_type, n.inputs, n.outputs, n.attrs) for n in m.graph.nodes]
    for k, v in m.graph.initializers.items():
        assert v.dtype == m2.graph.initializers[k].dtype and np.array_equal(v, m2.graph.initializers[k])
    assert [v.name for v in m2.graph.inputs] == ["cmap", "seq"] and m2.opset == 15
    assert m2.graph.inputs[1].shape[-1] == 26


def test_load_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        ox.load(str(tmp_path / "missing.onnx"))
    bad = tmp_path / "bad.onnx"
    bad.write_bytes(b"\xff\xff\xff\xffnot a protobuf")
    with pytest.raises(RuntimeError):
        ox.load(str(bad))


@pytest.mark.parametrize("kw", [dict(), dict(gc_activation="Relu", gc_bias=True, n_terms=320),
                                dict(lstm_hidden=256, gc_dims=(256, 128), fc_dim=512, n_terms=40)])
def test_plan_reads_hyperparameters_from_graph(kw):
    cfg = synth.GCNConfig(**kw)
    w = synth.make_weights(cfg, seed=11)
    plan = onnx_plan.plan_from_model(ox.loads(ox.dumps(synth.build_gcn_model(cfg, w))))
    assert plan.input_names == ["cmap", "seq"]
    assert plan.lstm_hidden == cfg.lstm_hidden and plan.lm_dim == cfg.lm_dim and plan.n_terms == cfg.n_terms
    assert [x.shape[1] for x in plan.gc_W] == list(cfg.gc_dims)
    assert plan.gc_activation == {"Relu": 1, "Elu": 2}[cfg.gc_activation]
    assert all((b is not None) == cfg.gc_bias for b in plan.gc_b)
    assert np.isclose(plan.eps, cfg.eps)
    assert np.array_equal(plan.lm_W, w["LM_embedding_W"]) and np.array_equal(plan.aa_W, w["AA_embedding_W"])
    assert np.array_equal(plan.fc_W, w["dense_W"]) and np.array_equal(plan.out_W, w["labels_W"])
    assert np.array_equal(plan.lstm_R[1], w["lstm2_R"])


def test_plan_rejects_foreign_graphs():
    m = synth.build_gcn_model(synth.GCNConfig(**spec.SMALL))
    g = m.graph
    cnn = ox.Model(ox.Graph(nodes=g.nodes, initializers=g.initializers, inputs=[g.inputs[1]], outputs=g.outputs))
    with pytest.raises(onnx_plan.UnsupportedModelError):
        onnx_plan.plan_from_model(cnn)            # single-input = CNN branch
    swapped = ox.Model(ox.Graph(nodes=g.nodes, initializers=g.initializers, inputs=g.inputs[::-1], outputs=g.outputs))
    with pytest.raises(onnx_plan.UnsupportedModelError):
        onnx_plan.plan_from_model(swapped)
    no_sm = ox.Model(ox.Graph(nodes=g.nodes[:-1], initializers=g.initializers, inputs=g.inputs, outputs=g.outputs))
    with pytest.raises(onnx_plan.UnsupportedModelError):
        onnx_plan.plan_from_model(no_sm)


def heads_share_lm():
    a = synth.make_weights(synth.GCNConfig(n_terms=489), seed=1)
    b = synth.make_weights(synth.GCNConfig(n_terms=320), seed=2)
    return a, b


def test_heads_share_language_model():
    a, b = heads_share_lm()
    pa = onnx_plan.plan_from_model(synth.build_gcn_model(synth.GCNConfig(n_terms=489), a))
    pb = onnx_plan.plan_from_model(synth.build_gcn_model(synth.GCNConfig(n_terms=320), b))
    assert pa.lm_fingerprint == pb.lm_fingerprint and not np.array_equal(pa.fc_W, pb.fc_W)


def deepfri_numpy(w, cfg, seq, cmap):
    """Independent float64 restatement of upstream DeepFRI's equations (not the ONNX interpreter)."""
    S = co.seq2onehot(seq).astype(np.float64)
    H = cfg.lstm_hidden
    sig = lambda x: 1 / (1 + np.exp(-x))

    def lstm(X, W, R, B):
        W, R = W[0].astype(np.float64), R[0].astype(np.float64)
        b = (B[0, :4 * H] + B[0, 4 * H:]).astype(np.float64)
        h = np.zeros(H); c = np.zeros(H); out = []
        for x in X:
            z = W @ x + R @ h + b
            i, o, f, g = sig(z[:H]), sig(z[H:2 * H]), sig(z[2 * H:3 * H]), np.tanh(z[3 * H:])
            c = f * c + i * g
            h = o * np.tanh(c)
            out.append(h)
        return np.array(out)
    h = lstm(lstm(S, w["lstm1_W"], w["lstm1_R"], w["lstm1_B"]), w["lstm2_W"], w["lstm2_R"], w["lstm2_B"])
    x = np.maximum(h @ w["LM_embedding_W"] + w["LM_embedding_b"] + S @ w["AA_embedding_W"], 0)
    A = cmap.astype(np.float64)
    A = A - np.diag(np.diag(A)) + np.eye(len(A))
    d = 1.0 / (cfg.eps + np.sqrt(A.sum(1)))
    An = d[:, None] * A * d[None, :]
    outs = []
    for l in range(1, len(cfg.gc_dims) + 1):
        z = (An @ x) @ w[f"GraphConv_{l}_W"]
        if cfg.gc_bias:
            z = z + w[f"GraphConv_{l}_b"]
        x = np.where(z > 0, z, np.exp(np.minimum(z, 0)) - 1) if cfg.gc_activation == "Elu" else np.maximum(z, 0)
        outs.append(x)
    p = np.concatenate(outs, 1).sum(0)
    f = np.maximum(p @ w["dense_W"] + w["dense_b"], 0)
    o = (f @ w["labels_W"] + w["labels_b"]).reshape(-1, 2)
    e = np.exp(o - o.max(1, keepdims=True))
    return (e / e.sum(1, keepdims=True))[:, 0]


@pytest.mark.parametrize("tag", ["small", "small_relu_bias"])
def test_oracle_matches_layer_equations_and_golden(tag, model_dir, gcn_golden):
    kw, seed, n, lo, hi = spec.GCN_CASES[tag]
    cfg = synth.GCNConfig(**kw)
    w = synth.make_weights(cfg, seed)
    p = go.Predictor(model_dir[tag])
    wl = golden_workload(tag)
    for i in range(n):
        cm = co.build_align_contact_map(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], spec.THRESHOLD, spec.GEN)
        y = p.forward_pass(wl.query_seqs[i], cm)
        assert y.dtype == np.float32 and y.shape == (cfg.n_terms,)
        assert np.abs(y - deepfri_numpy(w, cfg, wl.query_seqs[i], cm)).max() < 2e-5
        assert np.abs(y - gcn_golden[tag + "_scores"][i]).max() < 1e-6


def test_oracle_full_size_golden(model_dir, gcn_golden):
    p = go.Predictor(model_dir["mf"])
    wl = golden_workload("mf")
    cm = co.build_align_contact_map(wl.gapped_query[0], wl.gapped_target[0], wl.coords[0], spec.THRESHOLD, spec.GEN)
    assert np.abs(p.forward_pass(wl.query_seqs[0], cm) - gcn_golden["mf_scores"][0]).max() < 1e-6
