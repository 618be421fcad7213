"""Where does the tensor-core engine's score error come from, and how does it scale with protein length?
Compares every tap of the tc engine with the exact-fp32 SIMT engine for a few proteins per length class.
(Development aid.)   python tools/precision_probe.py [Lmin Lmax n] ..."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdf_pkg  # noqa: E402

mdf_pkg.load()
from metagenomic_deepfri_b200 import synth, predict, _lib  # noqa: E402


def main():
    classes = [(100, 150, 6), (400, 500, 6), (950, 1000, 6), (2300, 2500, 6)]
    d = tempfile.mkdtemp()
    path = os.path.join(d, "mf.onnx")
    synth.write_gcn_model(path, synth.GCNConfig())
    pred = predict.Predictor(path)
    _lib.default_context().set_debug_taps(True)
    for lo, hi, n in classes:
        wl = synth.make_workload(n, lo, hi, seed=lo, threshold=10.0)
        b = pred.upload(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)
        res = {}
        for eng in ("simt", "tc"):
            pred.set_engine(eng)
            pred.run(b, 10.0, 2)
            res[eng] = {k: pred.fetch(b, k) for k in ("lstm1", "lstm2", "x0", "gc_last", "pooled")}
            res[eng]["scores"] = pred.fetch_scores(b)
        line = [f"L {lo}-{hi}:"]
        for k in ("lstm1", "lstm2", "x0", "gc_last", "pooled", "scores"):
            a, t = res["simt"][k].astype(np.float64), res["tc"][k].astype(np.float64)
            e = a - t
            rms = np.sqrt((e ** 2).mean()) / max(np.sqrt((a ** 2).mean()), 1e-30)
            line.append(f"{k} max {np.abs(e).max():.2e} rel-rms {rms:.1e}")
        # coherent part of the per-residue error: |mean over residues of (tc - simt)| against the mean |value|
        off = b.seq_off
        for k in ("lstm2", "x0", "gc_last"):
            a, t = res["simt"][k].astype(np.float64), res["tc"][k].astype(np.float64)
            coh = []
            for i in range(len(wl)):
                e = (t - a)[off[i]:off[i + 1]]
                coh.append(np.abs(e.mean(0)).mean() / np.abs(a[off[i]:off[i + 1]]).mean())
            line.append(f"{k} coherent/|x| {np.mean(coh):.1e}")
        pe = np.abs(res["tc"]["pooled"] - res["simt"]["pooled"]).max(1) / np.abs(res["simt"]["pooled"]).max(1)
        line.append(f"pooled rel(max) per protein {pe.mean():.1e}")
        print("  ".join(line), flush=True)
        b.close()


if __name__ == "__main__":
    main()
