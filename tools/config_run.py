"""Run one BASELINE.json config (or a slice of it) through the tensor-core engine on the GPU: throughput,
per-stage profile, and the score difference against the exact-fp32 SIMT engine on the same batch.
(Development aid; asserting versions live in tests/.)

  python tools/config_run.py --config 4 --n 16384 [--check 256] [--head mf]
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdf_pkg  # noqa: E402

mdf_pkg.load()
from metagenomic_deepfri_b200 import synth, predict, _lib  # noqa: E402

HEADS = {"mf": 489, "bp": 1943, "cc": 320, "ec": 538}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--check", type=int, default=128, help="proteins also run through the fp32 SIMT engine")
    ap.add_argument("--head", default="mf")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--all-heads", action="store_true", help="run MF, BP, CC and EC on one uploaded batch (maps + LM shared)")
    args = ap.parse_args()
    full = {0: 1000, 1: 100_000, 2: 10_000, 3: 2000, 4: 1_000_000}[args.config]
    wl = synth.config_workload(args.config, args.n / full)
    lens = np.array([len(s) for s in wl.query_seqs])
    print(f"config {args.config}: n={len(wl)} L min/mean/max {lens.min()}/{lens.mean():.0f}/{lens.max()} residues {lens.sum()} "
          f"thr {wl.threshold}", flush=True)
    tmp = tempfile.mkdtemp()
    if args.all_heads:
        preds = {}
        for h, C in HEADS.items():
            pth = os.path.join(tmp, f"{h}.onnx")
            synth.write_gcn_model(pth, synth.GCNConfig(n_terms=C), seed=77 + C)
            preds[h] = predict.Predictor(pth)
        ctx = _lib.default_context()
        first = preds["mf"]
        for it in range(args.reps):
            ctx.synchronize()
            t0 = time.perf_counter()
            batch = first.upload(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)
            t1 = time.perf_counter()
            per = {}
            for h, p in preds.items():
                ta = time.perf_counter()
                p.run(batch, wl.threshold, wl.generated_contacts, share=True)
                ctx.synchronize()
                per[h] = (time.perf_counter() - ta) * 1e3
            dt = time.perf_counter() - t0
            batch.close()
            print(f"  4 heads run {it}: upload {1e3 * (t1 - t0):.0f} ms, heads " + ", ".join(f"{h} {v:.1f} ms" for h, v in per.items()) +
                  f" -> {len(wl) / dt:.0f} proteins/s for all four heads ({4 * len(wl) / dt:.0f} head-evaluations/s)", flush=True)
        return
    path = os.path.join(tmp, "m.onnx")
    synth.write_gcn_model(path, synth.GCNConfig(n_terms=HEADS[args.head]))
    pred = predict.Predictor(path)
    ctx = _lib.default_context()
    t0 = time.perf_counter()
    batch = pred.upload(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)
    print(f"upload (host packing + H2D) {time.perf_counter() - t0:.2f} s", flush=True)
    pred.set_engine("tc")
    for it in range(args.reps):
        ctx.synchronize()
        t0 = time.perf_counter()
        pred.run(batch, wl.threshold, wl.generated_contacts)
        ctx.synchronize()
        dt = time.perf_counter() - t0
        print(f"  tc run {it}: {dt * 1e3:.1f} ms -> {len(wl) / dt:.0f} proteins/s", flush=True)
    ctx.profile(True)
    pred.run(batch, wl.threshold, wl.generated_contacts)
    for name, ms, units in ctx.profile_report():
        print(f"    {name:24s} {ms:9.3f} ms  {units / (ms * 1e-3) / 1e12 if ms > 0 else 0:8.2f} T(units)/s")
    ctx.profile(False)
    tc = pred.fetch_scores(batch)
    print("tc scores finite:", bool(np.isfinite(tc).all()), "range", float(tc.min()), float(tc.max()))
    if args.check > 0:
        idx = np.linspace(0, len(wl) - 1, min(args.check, len(wl))).astype(int)
        sub = pred.upload([wl.query_seqs[i] for i in idx], [wl.gapped_query[i] for i in idx],
                          [wl.gapped_target[i] for i in idx], [wl.coords[i] for i in idx])
        pred.set_engine("simt")
        pred.run(sub, wl.threshold, wl.generated_contacts)
        ref = pred.fetch_scores(sub)
        pred.set_engine("tc")
        pred.run(sub, wl.threshold, wl.generated_contacts)
        got = pred.fetch_scores(sub)
        err = np.abs(got - ref).max(1)
        same = np.abs(tc[idx] - got).max()
        calls = ((got >= 0.1) != (ref >= 0.1)) & (np.abs(ref - 0.1) > 1e-3)
        print(f"tc vs simt on {len(idx)} proteins: max |d score| {err.max():.3e} (mean of per-protein max {err.mean():.3e}), "
              f"worst L={lens[idx][err.argmax()]}, differing GO calls outside the guard band: {int(calls.sum())}; "
              f"batch-composition invariance {same:.2e}")


if __name__ == "__main__":
    main()
