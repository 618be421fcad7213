"""Scratch: time the contact-map kernel variants (MDF_CMAP_VAR) on configs[1] and compare their packed maps bit for bit."""
import os, sys, json, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdf_pkg
mdf_pkg.load()
from metagenomic_deepfri_b200 import synth, predict, _lib

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 44, 42, 41, 22, 21, 14, 12, 11]
wl = synth.config_workload(1, pairs / 100_000)
path = os.path.join(tempfile.mkdtemp(), "m.onnx")
synth.write_gcn_model(path, synth.GCNConfig())
pred = predict.Predictor(path)
ctx = _lib.default_context()
batch = pred.upload(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)
ref = None
res = {}
for rnd in range(2):
    for v in variants:
        os.environ["MDF_CMAP_VAR"] = str(v)
        ts = []
        for _ in range(6):
            ctx.synchronize(); t0 = time.perf_counter()
            pred.run(batch, wl.threshold, wl.generated_contacts, upto=1)
            ctx.synchronize(); ts.append(time.perf_counter() - t0)
        dt = float(np.median(ts[2:]))
        if rnd == 0:
            packed = pred.fetch(batch, "packed")
            if ref is None:
                ref = packed.copy()
            same = bool(np.array_equal(packed, ref))
            res[v] = {"same": same}
        res[v][f"ms{rnd}"] = round(dt * 1e3, 3)
        res[v][f"Mpairs{rnd}"] = round(len(wl) / dt / 1e6, 3)
for v in variants:
    print(v, json.dumps(res[v]), flush=True)
