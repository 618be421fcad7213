"""Verbose on-GPU diagnostic: runs every entry point against the oracle and prints diffs.
(Development aid; the asserting versions live in tests/.)"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mdf_pkg  # noqa: E402

mdf_pkg.load()
from metagenomic_deepfri_b200 import synth, predict, bio_utils, contact_map_utils, batching, _lib  # noqa: E402
import cmap_oracle as co  # noqa: E402
import gcn_oracle as go  # noqa: E402


def section(name):
    print(f"\n=== {name}", flush=True)


def run(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
        sys.stdout.flush()


def check_cmap():
    section("cmap low-level")
    rng = np.random.default_rng(0)
    for n in (1, 2, 31, 32, 33, 127, 128, 129, 300, 1000):
        X = synth.random_walk_coords(rng, [n])[0]
        D = contact_map_utils.pairwise_sqeuclidean(X)
        Dr = co.pairwise_sqeuclidean(X)
        for thr in (6.0, 10.0):
            cm = bio_utils.calculate_contact_map(X, thr)
            cr = co.calculate_contact_map(X, thr)
            sp = bio_utils.calculate_contact_map(X, thr, mode="sparse")
            sr = co.calculate_contact_map(X, thr, mode="sparse")
            print(f"n={n} thr={thr} D eq {np.array_equal(D, Dr)} dense eq {np.array_equal(cm, cr)} "
                  f"sparse eq {sp.shape == sr.shape and np.array_equal(sp, sr)} nnz {sr.shape[0]}")
    X5 = rng.random((17, 5)).astype(np.float32)
    print("m=5 D eq", np.array_equal(contact_map_utils.pairwise_sqeuclidean(X5), co.pairwise_sqeuclidean(X5)))
    section("align_contact_map scatter")
    wl = synth.make_workload(40, 20, 400, seed=3, threshold=6.0)
    ok = 0
    for i in range(len(wl)):
        sp = co.calculate_contact_map(wl.coords[i], 6.0, "sparse")
        for gen in (0, 1, 2, 5):
            a = contact_map_utils.align_contact_map(wl.gapped_query[i], wl.gapped_target[i], sp, gen)
            b = co.align_contact_map(wl.gapped_query[i], wl.gapped_target[i], sp, gen)
            ok += int(np.array_equal(a, b))
    print(f"scatter equal {ok}/{len(wl) * 4}")
    section("fused build_align_contact_maps")

    class Aln:
        pass
    alns = []
    for i in range(len(wl)):
        a = Aln()
        a.target_name = f"t{i}"
        a.coords, a.gapped_sequence, a.gapped_target = wl.coords[i], wl.gapped_query[i], wl.gapped_target[i]
        alns.append(a)
    alns[3].coords = None
    # structure shorter / longer than the aligned target
    alns[5].coords = alns[5].coords[: len(alns[5].coords) // 2]
    alns[6].coords = np.concatenate([alns[6].coords, alns[6].coords[:7]])
    for thr, gen in ((6.0, 2), (10.0, 2), (10.0, 0), (6, 3)):
        t0 = time.time()
        dense = bio_utils.build_align_contact_maps(alns, thr, gen)
        packed = bio_utils.build_align_contact_maps(alns, thr, gen, packed=True)
        dt = time.time() - t0
        okd = okp = 0
        for i, a in enumerate(alns):
            want = co.build_align_contact_map(a.gapped_sequence, a.gapped_target, a.coords, thr, gen)
            if want is None:
                okd += dense[i] is None
                okp += packed[i] is None
                continue
            okd += int(np.array_equal(dense[i], want))
            okp += int(np.array_equal(batching.unpack_bits(packed[i], want.shape[0]), want))
            if not np.array_equal(dense[i], want):
                bad = np.argwhere(dense[i] != want)
                print("  mismatch protein", i, "L", want.shape[0], "n_bad", len(bad), bad[:5].tolist())
        print(f"thr={thr} gen={gen}: dense {okd}/{len(alns)} packed {okp}/{len(alns)} ({dt * 1e3:.1f} ms)")


def check_gcn():
    section("GCN forward vs oracle")
    import tempfile
    d = tempfile.mkdtemp()
    path = os.path.join(d, "mf.onnx")
    synth.write_gcn_model(path, synth.GCNConfig())
    t0 = time.time()
    pred = predict.Predictor(path)
    print(f"Predictor load {time.time() - t0:.2f}s")
    oracle = go.Predictor(path)
    wl = synth.make_workload(12, 30, 260, seed=11, threshold=10.0)
    cms = [co.build_align_contact_map(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], 10.0, 2) for i in range(len(wl))]
    want = np.stack([oracle.forward_pass(wl.query_seqs[i], cms[i]) for i in range(len(wl))])
    # intermediates from the oracle
    for i in range(2):
        t0 = time.time()
        y = pred.forward_pass(wl.query_seqs[i], cms[i])
        print(f"forward_pass L={len(wl.query_seqs[i])}: max err {np.abs(y - want[i]).max():.3e} ({(time.time() - t0) * 1e3:.1f} ms)")
    t0 = time.time()
    got = pred.forward_structures(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords, 10.0, 2)
    print(f"forward_structures n={len(wl)}: max err {np.abs(got - want).max():.3e} per-protein {np.abs(got - want).max(1)} ({(time.time() - t0) * 1e3:.1f} ms)")
    packed = [batching.pack_bits(c) for c in cms]
    got2 = pred.forward_batch(wl.query_seqs, packed)
    print(f"forward_batch: max err {np.abs(got2 - want).max():.3e}")
    # stage taps
    b = pred.upload(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)
    pred.run(b, 10.0, 2)
    sc = pred.fetch_scores(b)
    print(f"resident run: max err {np.abs(sc - want).max():.3e}")
    sess = oracle.session
    off = b.seq_off
    taps = {"lstm1": "lm/LSTM1_out", "lstm2": "lm/LSTM2_bm", "x0": "activation/Relu", "gc_last": "GraphConv_3/Elu", "pooled": "SumPooling/Sum"}
    got_t = {k: pred.fetch(b, k) for k in taps}
    deg = pred.fetch(b, "deg")
    for i in range(3):
        S = co.seq2onehot(wl.query_seqs[i])[None]
        outs = sess.run(list(taps.values()) + ["norm/d"], {"cmap": cms[i][None].astype(np.float32), "seq": S})
        for (k, _), o in zip(taps.items(), outs):
            o = np.asarray(o)
            g = got_t[k][i] if k == "pooled" else got_t[k][off[i]:off[i + 1]]
            o = o.reshape(g.shape)
            print(f"  protein {i} tap {k}: max abs err {np.abs(g - o).max():.3e} (ref absmax {np.abs(o).max():.3f})")
        print(f"  protein {i} deg err {np.abs(deg[off[i]:off[i + 1]] - outs[-1].reshape(-1)).max():.3e}")
    # error paths
    for bad in ("ACDJ",):
        try:
            pred.forward_pass(bad, np.eye(4, dtype=np.int32))
            print("invalid residue NOT rejected")
        except ValueError as e:
            print("invalid residue ->", e)
    try:
        pred.forward_pass("ACD", 2 * np.eye(3, dtype=np.int32))
        print("non-0/1 cmap NOT rejected")
    except ValueError as e:
        print("bad cmap ->", e)
    # timing of a bigger batch
    section("timing (simt engine)")
    wl = synth.config_workload(0, 0.25)
    b = pred.upload(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)
    ctx = _lib.default_context()
    for it in range(3):
        ctx.synchronize(); t0 = time.time()
        pred.run(b, 10.0, 2)
        ctx.synchronize(); dt = time.time() - t0
        print(f"  run {it}: n={len(wl)} {dt * 1e3:.1f} ms -> {len(wl) / dt:.0f} proteins/s")
    ctx.profile(True)
    pred.run(b, 10.0, 2)
    for name, ms, units in ctx.profile_report():
        print(f"    {name:24s} {ms:9.3f} ms  {units / ms / 1e9 if ms > 0 else 0:9.2f} G(units)/s")
    ctx.profile(False)
    t0 = time.time()
    sc = pred.forward_structures(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords, 10.0, 2)
    print(f"  e2e forward_structures: {(time.time() - t0) * 1e3:.1f} ms")
    print("  launches:", ctx.launch_count)


def check_tc():
    section("tensor-core engine vs oracle / simt")
    import tempfile
    d = tempfile.mkdtemp()
    path = os.path.join(d, "mf.onnx")
    synth.write_gcn_model(path, synth.GCNConfig())
    pred = predict.Predictor(path)
    _lib.default_context().set_debug_taps(True)
    oracle = go.Predictor(path)
    wl = synth.make_workload(int(os.environ.get("TC_N", "10")), 30, 330, seed=11, threshold=10.0)
    cms = [co.build_align_contact_map(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], 10.0, 2) for i in range(len(wl))]
    want = np.stack([oracle.forward_pass(wl.query_seqs[i], cms[i]) for i in range(len(wl))])
    b = pred.upload(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)
    res = {}
    for eng in ("simt", "tc"):
        pred.set_engine(eng)
        pred.run(b, 10.0, 2)
        res[eng] = (pred.fetch_scores(b), pred.fetch(b, "pooled"), pred.fetch(b, "gc_last"), pred.fetch(b, "lstm1"), pred.fetch(b, "lstm2"))
        print(f"engine {eng}: scores max err vs oracle {np.abs(res[eng][0] - want).max():.3e}  per-protein {np.abs(res[eng][0] - want).max(1)}")
    for k, name in ((3, "lstm1"), (4, "lstm2")):
        a, t = res["simt"][k], res["tc"][k]
        e = np.abs(a - t)
        print(f"  {name}: tc vs simt max abs {e.max():.3e} nan {np.isnan(t).sum()} rows>1e-2 {(e.max(1) > 1e-2).sum()}/{len(e)} cols>1e-2 {(e.max(0) > 1e-2).sum()}")
        if e.max() > 1e-2:
            off = b.seq_off
            for i in range(len(wl)):
                ei = e[off[i]:off[i + 1]]
                bad_t = np.flatnonzero(ei.max(1) > 1e-2)
                print(f"    protein {i} L={off[i+1]-off[i]} first bad step {bad_t[:3]} bad units at that step {np.flatnonzero(ei[bad_t[0]] > 1e-2)[:8] if len(bad_t) else ''}")
    for k, name in ((1, "pooled"), (2, "gc_last")):
        a, t = res["simt"][k], res["tc"][k]
        print(f"  {name}: tc vs simt max abs {np.abs(a - t).max():.3e} (absmax {np.abs(a).max():.3f}) nan {np.isnan(t).sum()}")
    off = b.seq_off
    a, t = res["simt"][2], res["tc"][2]
    for i in range(len(wl)):
        e = np.abs(a[off[i]:off[i + 1]] - t[off[i]:off[i + 1]])
        print(f"    protein {i} L={off[i + 1] - off[i]} gc_last err max {e.max():.3e} rows>1e-2: {(e.max(1) > 1e-2).sum()} cols>1e-2: {(e.max(0) > 1e-2).sum()}")
    pred.set_engine("tc")
    _lib.default_context().set_debug_taps(False)
    wl = synth.make_workload(int(os.environ.get("TC_TIMING_N", "1000")), 100, 500, seed=1, threshold=10.0)
    b2 = pred.upload(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)
    ctx = _lib.default_context()
    for it in range(3):
        ctx.synchronize(); t0 = time.time()
        pred.run(b2, 10.0, 2)
        ctx.synchronize(); dt = time.time() - t0
        print(f"  tc run {it}: n={len(wl)} {dt * 1e3:.1f} ms -> {len(wl) / dt:.0f} proteins/s")
    ctx.profile(True)
    pred.run(b2, 10.0, 2)
    for name, ms, units in ctx.profile_report():
        print(f"    {name:24s} {ms:9.3f} ms  {units / (ms * 1e-3) / 1e12 if ms > 0 else 0:9.2f} T(units)/s")
    ctx.profile(False)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cmap", "gcn"]
    if "tc" in which:
        run(check_tc)
    if "cmap" in which:
        run(check_cmap)
    if "gcn" in which:
        run(check_gcn)
