"""Pattern-matches a DeepFRI GCN `.onnx` graph into the fused pipeline's weight set.

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
        _fail("single-input model (the sequence-only CNN branch, predict.pyx:91-95) is out of scope")
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


def load_plan(path: str) -> GCNPlan:
    return plan_from_model(ox.load(path))
