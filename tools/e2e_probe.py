"""Where does the end-to-end step go?  Streams configs[4] chunks from Python lists through submit_structures / wait with the
per-stage CUDA-event profile on, and prints the stage sums per chunk beside the wall clock per chunk.  (Development aid.)"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdf_pkg  # noqa: E402

mdf_pkg.load()
from metagenomic_deepfri_b200 import synth, predict, distributed, _lib  # noqa: E402
import torch  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
prof = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
wl = synth.config_workload(4, 2 * n / 1_000_000)
chunks = [synth.Workload(wl.query_seqs[lo:lo + n], wl.gapped_query[lo:lo + n], wl.gapped_target[lo:lo + n], wl.coords[lo:lo + n],
                         wl.threshold, wl.generated_contacts, wl.name) for lo in range(0, len(wl), n)][:2]
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "m.onnx")
synth.write_gcn_model(path, synth.GCNConfig(n_terms=489))
pred = predict.Predictor(path)
ctx = _lib.default_context()
C = 489
out = torch.empty((steps * n, C), dtype=torch.float32, pin_memory=True).numpy()


def submit(c, rows):
    return pred.submit_structures(c.query_seqs, c.gapped_query, c.gapped_target, c.coords, wl.threshold, 2, out=rows)


distributed.stream_chunks(submit, [(chunks[0], n)], out[:n])
distributed.stream_chunks(submit, [(chunks[1], len(chunks[1]))], out[:len(chunks[1])])
# resident leg
b = pred.upload(chunks[0].query_seqs, chunks[0].gapped_query, chunks[0].gapped_target, chunks[0].coords)
for _ in range(3):
    pred.run(b, wl.threshold, 2)
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    pred.run(b, wl.threshold, 2)
ctx.synchronize()
print(f"resident: {(time.perf_counter() - t0) / steps * 1e3:.2f} ms per chunk (back to back, no flush)")
b.close()
if prof:
    ctx.profile(True)
t0 = time.perf_counter()
distributed.stream_chunks(submit, [(chunks[s % 2], len(chunks[s % 2])) for s in range(steps)], out)
dt = time.perf_counter() - t0
print(f"e2e: {dt / steps * 1e3:.2f} ms per chunk over {steps} chunks")
if prof:
    rep = ctx.profile_report()
    ctx.profile(False)
    agg = {}
    for name, ms, units in rep:
        agg[name] = agg.get(name, 0.0) + ms
    print("stage sums per chunk:", {k: round(v / steps, 3) for k, v in agg.items()}, "total", round(sum(agg.values()) / steps, 2))
