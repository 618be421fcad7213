"""Pattern-matches a DeepFRI GCN (or sequence-only DeepCNN) `.onnx` graph into the fused pipeline's weight set.

The reference never looks inside the model file: it hands it to onnxruntime
(`mDeepFRI/predict.pyx:62-73`).  The B200 path runs a fixed fused pipeline (LSTM-LM ->
embedding -> GraphConv stack -> sum-pool -> dense -> FuncPredictor), so the
graph is *recognised* here rather than interpreted: every weight is located by its role in
the dataflow (what it is connected to and its shape), never by its name, and every
hyper-parameter (H, E, layer widths, activation, bias presence, epsilon, C) is read from the
initialisers / op types.  Anything that does not fit raises `UnsupportedModelError` with
the reason - there is no fallback executor.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Set

import numpy as np

from . import onnx_lite as ox

_SHAPE_OPS = {"Transpose", "Squeeze", "Unsqueeze", "Reshape", "Identity", "Dropout", "Cast", "Flatten"}
_ACTS = {"Relu": 1, "Elu": 2}


class UnsupportedModelError(RuntimeError):
    pass


@dataclass
class GCNPlan:
    input_names: List[str]
    n_channels: int
    lstm_hidden: int
    lstm_W: List[np.ndarray]
    lstm_R: List[np.ndarray]
    lstm_B: List[Optional[np.ndarray]]
    lm_dim: int
    aa_W: np.ndarray
    lm_W: np.ndarray
    lm_b: Optional[np.ndarray]
    gc_W: List[np.ndarray]
    gc_b: List[Optional[np.ndarray]]
    gc_activation: int
    gc_alpha: float
    eps: float
    fc_W: np.ndarray
    fc_b: Optional[np.ndarray]
    out_W: np.ndarray
    out_b: Optional[np.ndarray]
    n_terms: int
    lm_fingerprint: str = ""
    notes: List[str] = field(default_factory=list)


def _fail(msg: str):
    raise UnsupportedModelError(f"ONNX graph is not a supported DeepFRI GCN head: {msg}")


def plan_from_model(model: ox.Model) -> GCNPlan:
    g = model.graph
    init = g.initializers
    if len(g.inputs) == 1:
        _fail("single-input model: this is a sequence-only DeepCNN head, see cnn_plan_from_model")
    if len(g.inputs) != 2:
        _fail(f"expected 2 inputs (cmap, seq), found {len(g.inputs)}")
    in_names = [vi.name for vi in g.inputs]
    seq_in = [vi for vi in g.inputs if len(vi.shape) == 3 and vi.shape[-1] == 26]
    if len(seq_in) != 1:
        _fail("cannot identify the [batch, L, 26] one-hot sequence input")
    seq_name = seq_in[0].name
    cmap_name = [n for n in in_names if n != seq_name][0]
    if in_names[0] != cmap_name:
        _fail("inputs must be ordered (cmap, seq) as the reference feeds them (predict.pyx:87-90)")

    producer: Dict[str, ox.Node] = {}
    consumers: Dict[str, List[ox.Node]] = {}
    for n in g.nodes:
        for o in n.outputs:
            producer[o] = n
        for i in n.inputs:
            consumers.setdefault(i, []).append(n)
    # Constant nodes behave like initialisers
    const: Dict[str, np.ndarray] = dict(init)
    for n in g.nodes:
        if n.op_type == "Constant" and "value" in n.attrs:
            const[n.outputs[0]] = np.asarray(n.attrs["value"])

    memo: Dict[str, Set[str]] = {}

    def deps(t: str) -> Set[str]:
        """Roles a tensor depends on: 'seq', 'cmap', 'lstm', 'dyn' (dynamic x dynamic MatMul)."""
        if t in memo:
            return memo[t]
        memo[t] = set()
        out: Set[str] = set()
        if t == seq_name:
            out.add("seq")
        elif t == cmap_name:
            out.add("cmap")
        elif t in producer:
            n = producer[t]
            if n.op_type == "LSTM":
                out.add("lstm")
            dyn_in = [i for i in n.inputs if i and i not in const]
            if n.op_type == "MatMul" and len(dyn_in) == 2:
                out.add("dyn")
            for i in dyn_in:
                out |= deps(i)
        memo[t] = out
        return out

    def skip_shape_ops_back(t: str) -> str:
        while t in producer and producer[t].op_type in _SHAPE_OPS:
            t = producer[t].inputs[0]
        return t

    def next_compute(t: str) -> List[ox.Node]:
        """Consumers of t, looking through pure shape ops."""
        out = []
        for n in consumers.get(t, []):
            if n.op_type in _SHAPE_OPS:
                out += next_compute(n.outputs[0])
            else:
                out.append(n)
        return out

    def bias_after(n: ox.Node):
        """(bias array or None, tensor name after the optional bias Add)."""
        t = n.outputs[0]
        for c in next_compute(t):
            if c.op_type == "Add":
                other = [i for i in c.inputs if i in const]
                if len(other) == 1:
                    arr = np.asarray(const[other[0]], np.float32)
                    if arr.ndim >= 1 and arr.size == arr.shape[-1]:      # [N] or [1,..,N]
                        return arr.reshape(-1), c.outputs[0]
        return None, t

    # ---- LSTM stack
    lstms = [n for n in g.nodes if n.op_type == "LSTM"]
    if not lstms:
        _fail("no LSTM language-model layers found")
    H = int(lstms[0].attrs.get("hidden_size", 0))
    Ws, Rs, Bs = [], [], []
    prev_out = None
    for k, n in enumerate(lstms):
        if n.attrs.get("direction", "forward") != "forward" or n.attrs.get("layout", 0) != 0:
            _fail("only forward, layout-0 LSTM layers are supported")
        if "activations" in n.attrs and [a.lower() for a in n.attrs["activations"]] != ["sigmoid", "tanh", "tanh"]:
            _fail("LSTM uses non-default activations")
        if int(n.attrs.get("hidden_size", 0)) != H:
            _fail("stacked LSTM layers have different hidden sizes")
        if len(n.inputs) > 4 and any(n.inputs[4:]):
            _fail("LSTM with sequence_lens / initial state / peepholes is not supported")
        src = skip_shape_ops_back(n.inputs[0])
        if k == 0 and src != seq_name:
            _fail("first LSTM layer is not fed by the sequence input")
        if k > 0 and (src not in producer or producer[src] is not prev_out):
            _fail("LSTM layers are not stacked")
        W, R = const.get(n.inputs[1]), const.get(n.inputs[2])
        B = const.get(n.inputs[3]) if len(n.inputs) > 3 and n.inputs[3] else None
        if W is None or R is None:
            _fail("LSTM weights are not constant initialisers")
        exp_in = 26 if k == 0 else H
        if W.shape != (1, 4 * H, exp_in) or R.shape != (1, 4 * H, H) or (B is not None and B.shape != (1, 8 * H)):
            _fail(f"LSTM layer {k + 1}: unexpected weight shapes W{W.shape} R{R.shape}")
        Ws.append(np.ascontiguousarray(W, np.float32))
        Rs.append(np.ascontiguousarray(R, np.float32))
        Bs.append(None if B is None else np.ascontiguousarray(B, np.float32))
        prev_out = n

    # ---- constant-weight MatMuls, classified by what they depend on
    aa = lm = None
    gcs, heads = [], []
    n_dyn = 0
    for n in g.nodes:
        if n.op_type != "MatMul":
            continue
        cw = [i for i in n.inputs if i in const]
        if len(cw) == 0:
            n_dyn += 1
            continue
        if len(cw) != 1 or n.inputs[1] != cw[0]:
            _fail(f"MatMul {n.name!r}: weight must be the right operand")
        w = np.asarray(const[cw[0]], np.float32)
        if w.ndim != 2:
            _fail(f"MatMul {n.name!r}: weight rank {w.ndim}")
        d = deps(n.inputs[0])
        if "cmap" not in d and "lstm" not in d and "seq" in d:
            if aa is not None:
                _fail("more than one sequence embedding MatMul")
            aa = (n, w)
        elif "cmap" not in d and "lstm" in d:
            if lm is not None:
                _fail("more than one language-model embedding MatMul")
            lm = (n, w)
        elif "cmap" in d:
            pooled = any(p.op_type == "ReduceSum" for p in _ancestors_until_matmul(n.inputs[0], producer, const))
            (heads if pooled or heads else gcs).append((n, w))
        else:
            _fail(f"MatMul {n.name!r} has no recognised role")
    if aa is None or lm is None:
        _fail("embedding layers (AA_embedding / LM_embedding) not found")
    if aa[1].shape[0] != 26 or lm[1].shape[0] != H or aa[1].shape[1] != lm[1].shape[1]:
        _fail(f"embedding shapes {aa[1].shape} / {lm[1].shape} are inconsistent")
    E = int(aa[1].shape[1])
    lm_b, lm_t = bias_after(lm[0])
    aa_b, _ = bias_after(aa[0])
    if aa_b is not None:
        _fail("AA_embedding with a bias is not supported")
    if not gcs or n_dyn != len(gcs):
        _fail(f"found {len(gcs)} GraphConv weight MatMuls but {n_dyn} adjacency products")
    if len(heads) != 2:
        _fail(f"expected dense + output layers after pooling, found {len(heads)} MatMuls")

    gc_W, gc_b, acts, alphas = [], [], [], []
    prev = E
    for k, (n, w) in enumerate(gcs):
        if w.shape[0] != prev:
            _fail(f"GraphConv layer {k + 1}: weight {w.shape} does not follow width {prev}")
        b, t = bias_after(n)
        # activation: either directly after this MatMul (A.X).W order or after the adjacency product
        act, alpha = 0, 1.0
        cur = t
        for _ in range(3):
            nxt = next_compute(cur)
            hit = [c for c in nxt if c.op_type in _ACTS]
            if hit:
                act, alpha = _ACTS[hit[0].op_type], float(hit[0].attrs.get("alpha", 1.0))
                break
            mm = [c for c in nxt if c.op_type == "MatMul" and not any(i in const for i in c.inputs)]
            if not mm:
                break
            cur = mm[0].outputs[0]
        gc_W.append(w); gc_b.append(b); acts.append(act); alphas.append(alpha)
        prev = int(w.shape[1])
    if len(set(acts)) != 1 or len(set(alphas)) != 1:
        _fail("GraphConv layers use different activations")
    G = int(sum(w.shape[1] for w in gc_W))
    (fc_n, fc_W), (out_n, out_W) = heads
    if fc_W.shape[0] != G:
        _fail(f"dense layer expects {fc_W.shape[0]} pooled features, GraphConv stack gives {G} "
              "(per-layer outputs must be concatenated)")
    fc_b, fc_t = bias_after(fc_n)
    if not any(c.op_type == "Relu" for c in next_compute(fc_t)):
        _fail("dense layer after pooling is not followed by ReLU")
    if out_W.shape[0] != fc_W.shape[1] or out_W.shape[1] % 2:
        _fail(f"output layer weight {out_W.shape} inconsistent")
    out_b, out_t = bias_after(out_n)
    C = out_W.shape[1] // 2
    sm = [n for n in g.nodes if n.op_type == "Softmax"]
    if len(sm) != 1 or sm[0].outputs[0] != g.outputs[0].name:
        _fail("graph does not end in a single Softmax")
    if sm[0].attrs.get("axis", -1) not in (-1, 2):
        _fail("Softmax is not over the last axis")
    if not any(c.op_type == "Relu" for c in next_compute(_after_add(lm_t, aa[0].outputs[0], consumers))):
        _fail("embedding sum is not followed by ReLU")

    # ---- degree-normalisation epsilon: the scalar added to sqrt(rowsum)
    eps = None
    for n in g.nodes:
        if n.op_type == "Sqrt" and "cmap" in deps(n.inputs[0]):
            for c in next_compute(n.outputs[0]):
                if c.op_type == "Add":
                    k = [i for i in c.inputs if i in const and const[i].size == 1]
                    if k:
                        eps = float(np.asarray(const[k[0]]).reshape(()))
    if eps is None:
        _fail("degree normalisation (1 / (eps + sqrt(rowsum))) not found")

    import hashlib
    hsh = hashlib.sha256()
    for a in Ws + Rs + [b for b in Bs if b is not None]:
        hsh.update(a.tobytes())
    return GCNPlan(input_names=in_names, n_channels=26, lstm_hidden=H, lstm_W=Ws, lstm_R=Rs, lstm_B=Bs,
                   lm_dim=E, aa_W=np.ascontiguousarray(aa[1]), lm_W=np.ascontiguousarray(lm[1]), lm_b=lm_b,
                   gc_W=[np.ascontiguousarray(w) for w in gc_W], gc_b=gc_b, gc_activation=acts[0],
                   gc_alpha=alphas[0], eps=eps, fc_W=np.ascontiguousarray(fc_W), fc_b=fc_b,
                   out_W=np.ascontiguousarray(out_W), out_b=out_b, n_terms=int(C), lm_fingerprint=hsh.hexdigest())


def _ancestors_until_matmul(t: str, producer, const):
    """Nodes between tensor t and the closest upstream MatMuls (exclusive)."""
    out, stack, seen = [], [t], set()
    while stack:
        x = stack.pop()
        if x in seen or x not in producer:
            continue
        seen.add(x)
        n = producer[x]
        if n.op_type == "MatMul":
            continue
        out.append(n)
        stack += [i for i in n.inputs if i and i not in const]
    return out


def _after_add(a: str, b: str, consumers) -> str:
    """Output of the Add that sums tensors a and b (the 'Embedding' layer upstream)."""
    for n in consumers.get(a, []):
        if n.op_type == "Add" and b in n.inputs:
            return n.outputs[0]
    _fail("LM_embedding and AA_embedding are not summed")


@dataclass
class CNNPlan:
    """Sequence-only DeepCNN head (`predict.pyx:91-95`): parallel Conv1D -> concat -> BatchNormalization -> ReLU ->
    max over residues -> dense -> softmax.  Conv bias and BatchNormalization are folded into `scale` / `shift`."""
    input_names: List[str]
    n_channels: int
    conv_W: List[np.ndarray]          # [F, 26, width] each, concat order
    conv_pad_left: List[int]
    scale: np.ndarray                 # [sum F]
    shift: np.ndarray                 # [sum F]
    out_W: np.ndarray                 # [sum F, 2C]
    out_b: Optional[np.ndarray]
    n_terms: int
    notes: List[str] = field(default_factory=list)


def _cnn_fail(msg: str):
    raise UnsupportedModelError(f"ONNX graph is not a supported DeepCNN head: {msg}")


def cnn_plan_from_model(model: ox.Model) -> CNNPlan:
    g = model.graph
    if len(g.inputs) != 1:
        _cnn_fail(f"expected the one-hot sequence as the single input, found {len(g.inputs)} inputs")
    vi = g.inputs[0]
    if len(vi.shape) != 3 or vi.shape[-1] != 26:
        _cnn_fail(f"input {vi.name!r} is not [batch, L, 26]")
    seq_name = vi.name
    const: Dict[str, np.ndarray] = dict(g.initializers)
    producer: Dict[str, ox.Node] = {}
    consumers: Dict[str, List[ox.Node]] = {}
    for n in g.nodes:
        if n.op_type == "Constant" and "value" in n.attrs:
            const[n.outputs[0]] = np.asarray(n.attrs["value"])
        for o in n.outputs:
            producer[o] = n
        for i in n.inputs:
            consumers.setdefault(i, []).append(n)

    def back(t: str) -> str:
        while t in producer and producer[t].op_type in _SHAPE_OPS:
            t = producer[t].inputs[0]
        return t

    def fwd(t: str) -> List[ox.Node]:
        out = []
        for n in consumers.get(t, []):
            out += fwd(n.outputs[0]) if n.op_type in _SHAPE_OPS else [n]
        return out

    def only(nodes: List[ox.Node], what: str) -> ox.Node:
        if len(nodes) != 1:
            _cnn_fail(f"expected exactly one consumer {what}, found {[n.op_type for n in nodes]}")
        return nodes[0]

    allowed = {"Conv", "Concat", "BatchNormalization", "Relu", "ReduceMax", "GlobalMaxPool", "MatMul", "Gemm", "Add", "Softmax",
               "Constant"} | _SHAPE_OPS
    for n in g.nodes:
        if n.op_type not in allowed:
            _cnn_fail(f"unexpected operator {n.op_type} ({n.name!r}); a language-model DeepCNN variant is not supported")

    convs = [n for n in g.nodes if n.op_type == "Conv"]
    if not convs:
        _cnn_fail("no Conv layers found")
    info = {}
    for n in convs:
        if back(n.inputs[0]) != seq_name:
            _cnn_fail(f"Conv {n.name!r} is not fed by the sequence input")
        W = const.get(n.inputs[1])
        if W is None:
            _cnn_fail(f"Conv {n.name!r}: weights are not constant")
        W = np.asarray(W, np.float32)
        if W.ndim == 4 and W.shape[2] == 1:
            W3, two_d = W[:, :, 0, :], True
        elif W.ndim == 3:
            W3, two_d = W, False
        else:
            _cnn_fail(f"Conv {n.name!r}: weight shape {W.shape} is not [F, 26, 1, w] / [F, 26, w]")
        F, Cin, k = W3.shape
        if Cin != 26:
            _cnn_fail(f"Conv {n.name!r}: {Cin} input channels")
        a = n.attrs
        if a.get("group", 1) != 1 or any(int(v) != 1 for v in a.get("strides", [1])) or any(int(v) != 1 for v in a.get("dilations", [1])):
            _cnn_fail(f"Conv {n.name!r}: only group 1 / stride 1 / dilation 1")
        auto = a.get("auto_pad", "NOTSET")
        if auto == "SAME_UPPER":
            pl, pr = (k - 1) // 2, k - 1 - (k - 1) // 2
        elif auto == "SAME_LOWER":
            pl, pr = k - 1 - (k - 1) // 2, (k - 1) // 2
        elif auto == "NOTSET":
            pads = [int(v) for v in a.get("pads", [0, 0, 0, 0] if two_d else [0, 0])]
            if two_d:
                if len(pads) != 4 or pads[0] or pads[2]:
                    _cnn_fail(f"Conv {n.name!r}: pads {pads}")
                pl, pr = pads[1], pads[3]
            else:
                if len(pads) != 2:
                    _cnn_fail(f"Conv {n.name!r}: pads {pads}")
                pl, pr = pads
        else:
            _cnn_fail(f"Conv {n.name!r}: auto_pad {auto}")
        if pl + pr != k - 1:
            _cnn_fail(f"Conv {n.name!r}: padding ({pl}, {pr}) does not keep the sequence length for width {k} ('same' expected)")
        b = np.zeros(F, np.float32)
        if len(n.inputs) > 2 and n.inputs[2]:
            if n.inputs[2] not in const:
                _cnn_fail(f"Conv {n.name!r}: bias is not constant")
            b = np.asarray(const[n.inputs[2]], np.float32).reshape(-1)
        info[n.outputs[0]] = (np.ascontiguousarray(W3), int(pl), b)

    # concat order = channel order
    nxt = {id(c): c for n in convs for c in fwd(n.outputs[0])}
    if len(convs) > 1:
        cat = only(list(nxt.values()), "of the Conv outputs (Concat)")
        if cat.op_type != "Concat" or cat.attrs.get("axis") not in (1, -2):
            _cnn_fail("Conv outputs are not concatenated over the channel axis (axis 1 of [b, F, L])")
        order = [back(i) for i in cat.inputs]
        if sorted(order) != sorted(info):
            _cnn_fail("Concat inputs are not exactly the Conv outputs")
        cur = cat.outputs[0]
    else:
        order = [convs[0].outputs[0]]
        cur = convs[0].outputs[0]
    conv_W = [info[o][0] for o in order]
    pad_left = [info[o][1] for o in order]
    bias = np.concatenate([info[o][2] for o in order])
    tot = int(bias.size)
    scale, shift = np.ones(tot, np.float32), bias.copy()
    n = only(fwd(cur), "after the concatenation")
    if n.op_type == "BatchNormalization":
        try:
            gam, bet, mean, var = (np.asarray(const[i], np.float64).reshape(-1) for i in n.inputs[1:5])
        except KeyError:
            _cnn_fail("BatchNormalization parameters are not constant")
        if any(v.size != tot for v in (gam, bet, mean, var)):
            _cnn_fail("BatchNormalization width does not match the concatenated filters")
        s64 = gam / np.sqrt(var + float(n.attrs.get("epsilon", 1e-5)))
        scale = s64.astype(np.float32)
        shift = ((bias.astype(np.float64) - mean) * s64 + bet).astype(np.float32)
        n = only(fwd(n.outputs[0]), "after BatchNormalization")
    if n.op_type != "Relu":
        _cnn_fail(f"expected ReLU before the pooling, found {n.op_type}")
    n = only(fwd(n.outputs[0]), "after ReLU")
    if n.op_type == "ReduceMax":
        if [int(v) for v in n.attrs.get("axes", [])] not in ([2], [-1]):
            _cnn_fail("ReduceMax is not over the residue axis of [b, F, L]")
    elif n.op_type != "GlobalMaxPool":
        _cnn_fail(f"expected a global max-pool over residues, found {n.op_type}")
    n = only(fwd(n.outputs[0]), "after the max-pool")
    if n.op_type not in ("MatMul", "Gemm") or n.inputs[1] not in const:
        _cnn_fail(f"expected the FuncPredictor dense layer after the max-pool, found {n.op_type}")
    out_W = np.asarray(const[n.inputs[1]], np.float32)
    if n.op_type == "Gemm":
        if n.attrs.get("transA", 0) or float(n.attrs.get("alpha", 1.0)) != 1.0 or float(n.attrs.get("beta", 1.0)) != 1.0:
            _cnn_fail("Gemm with transA / alpha / beta is not supported")
        if n.attrs.get("transB", 0):
            out_W = out_W.T
    if out_W.ndim != 2 or out_W.shape[0] != tot or out_W.shape[1] % 2:
        _cnn_fail(f"output layer weight {out_W.shape} does not follow {tot} pooled channels")
    out_b = None
    if n.op_type == "Gemm" and len(n.inputs) > 2 and n.inputs[2]:
        out_b = np.asarray(const[n.inputs[2]], np.float32).reshape(-1)
        t = n.outputs[0]
    else:
        t = n.outputs[0]
        for c in fwd(t):
            if c.op_type == "Add":
                other = [i for i in c.inputs if i in const]
                if len(other) == 1:
                    out_b, t = np.asarray(const[other[0]], np.float32).reshape(-1), c.outputs[0]
    if out_b is not None and out_b.size != out_W.shape[1]:
        _cnn_fail("output bias width mismatch")
    sm = only(fwd(t), "after the output layer")
    if sm.op_type != "Softmax" or sm.outputs[0] != g.outputs[0].name or sm.attrs.get("axis", -1) not in (-1, 2):
        _cnn_fail("graph does not end in a Softmax over the last axis")
    return CNNPlan(input_names=[seq_name], n_channels=26, conv_W=conv_W, conv_pad_left=pad_left, scale=scale, shift=shift,
                   out_W=np.ascontiguousarray(out_W), out_b=out_b, n_terms=int(out_W.shape[1] // 2))


def load_plan(path: str):
    """`GCNPlan` for a two-input (cmap, seq) model, `CNNPlan` for a single-input sequence model."""
    model = ox.load(path)
    if len(model.graph.inputs) == 1:
        return cnn_plan_from_model(model)
    return plan_from_model(model)
