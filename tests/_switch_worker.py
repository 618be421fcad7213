"""Worker of tests/test_gpu_switches.py: runs the tensor-core engine under the environment switches it was started with and
compares with the exact-fp32 SIMT engine (and the contact maps with the oracle).  Prints SWITCH_OK on success."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import mdf_pkg  # noqa: E402

mdf_pkg.load()
from metagenomic_deepfri_b200 import batching, predict, synth  # noqa: E402
import cmap_oracle as co  # noqa: E402
import spec  # noqa: E402

# MDF_TEST_FULL_MODEL=1: the full-size head (H = 512: the fused LSTM kernel's straight-line CTA-pair issuer), else the smallest tc shape
kw = {} if os.environ.get("MDF_TEST_FULL_MODEL") else spec.GCN_CASES["tc_small"][0]
path = os.path.join(tempfile.mkdtemp(), "m.onnx")
synth.write_gcn_model(path, synth.GCNConfig(**kw), seed=8)
pred = predict.Predictor(path)
wl = synth.make_workload(300, 1, 330, seed=21, threshold=10.0)
if os.environ.get("MDF_TEST_FULL_MODEL"):
    wl = synth.make_workload(300, 20, 260, seed=21, threshold=10.0)
pred.set_engine("simt")
want = pred.forward_structures(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords, 10.0, 2)
pred.set_engine("tc")
got = pred.forward_structures(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords, 10.0, 2)
err = float(np.abs(got - want).max())
assert err <= 1e-3, f"tc vs simt {err:.2e} under {[k for k in os.environ if k.startswith('MDF_')]}"
b = pred.upload(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)
pred.run(b, 10.0, 2, upto=1)
packed = pred.fetch(b, "packed")
for i in range(0, len(wl), 7):
    m = co.build_align_contact_map(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], 10.0, 2)
    L = m.shape[0]
    assert np.array_equal(batching.unpack_bits(packed[b.packed_off[i]:b.packed_off[i + 1]].reshape(L, -1), L), m), i
print(f"SWITCH_OK max |tc - simt| = {err:.2e}")
