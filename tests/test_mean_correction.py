"""The arithmetic behind the single-term X.W of GraphConv layers 2 and 3 (tc_engine.cu, DESIGN.md section 4 "Precision"), emulated in
NumPy: the rounding error of an fp16 weight matrix is coherent across residues, its coherent part is mean_p . (W - fp16(W)) per
protein, and adding that rank-one term to a product with fp16(W) alone recovers the two-term (hi + lo) result after the sum-pool.
CPU only; the GPU tests pin the kernels themselves against the exact-fp32 engine."""
import numpy as np
import pytest


def f16(x):
    return x.astype(np.float16).astype(np.float64)


def elu(x):
    return np.where(x > 0, x, np.exp(np.minimum(x, 0)) - 1)


def gcn_pooled(X0, A, Ws, mode):
    """Three GraphConv layers on one protein, Y = X W with the weights handled as `mode` says; returns the concatenated pools."""
    A = A.astype(np.float64).copy()
    np.fill_diagonal(A, 1.0)
    d = 1.0 / (1e-6 + np.sqrt(A.sum(1)))
    X, pools = X0, []
    for l, W in enumerate(Ws):
        Xin = X if mode == "exact" else f16(X)
        if mode == "exact":
            Y = X @ W
        elif mode == "hilo" or l == 0:                       # the first layer keeps its two terms in every non-exact mode
            Y = Xin @ (f16(W) + f16(W - f16(W)))
        elif mode == "hi":
            Y = Xin @ f16(W)
        else:                                                # "hi_mean": one term + per-protein mean correction
            Y = Xin @ f16(W) + (pools[-1] / len(X)) @ (W - f16(W))
        X = elu(d[:, None] * (A @ (d[:, None] * Y)))
        pools.append(X.sum(0))
    return np.concatenate(pools)


@pytest.mark.parametrize("L", [150, 900])
def test_mean_correction_recovers_the_second_weight_term(L):
    rng = np.random.default_rng(L)
    K = 192
    X0 = np.maximum(rng.normal(0.3, 0.6, size=(L, 2 * K)), 0)                 # post-ReLU embedding: non-zero mean, like the real X0
    Ws = [rng.uniform(-1, 1, size=(2 * K, K)) * np.sqrt(6.0 / (2 * K)) * 1.4] + \
         [rng.uniform(-1, 1, size=(K, K)) * np.sqrt(6.0 / K) * 1.4 for _ in range(2)]
    # banded contact map with a few long-range patches
    i = np.arange(L)
    A = (np.abs(i[:, None] - i[None, :]) <= 3).astype(np.int32)
    for _ in range(L // 20):
        a, b = rng.integers(0, L - 8, 2)
        A[a:a + 6, b:b + 6] = 1
        A[b:b + 6, a:a + 6] = 1
    ref = gcn_pooled(X0, A, Ws, "exact")
    scale = np.abs(ref).mean()
    err = {m: np.abs(gcn_pooled(X0, A, Ws, m) - ref).mean() / scale for m in ("hilo", "hi", "hi_mean")}
    # the hi term alone is several times worse than two terms; the mean-corrected single term is back at the two-term level
    assert err["hi"] > 3 * err["hilo"], err
    assert err["hi_mean"] < 1.5 * err["hilo"], err
