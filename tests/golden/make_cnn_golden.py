"""Generates tests/golden/cnn_golden.npz: outputs of oracle/gcn_oracle.py (fp32 NumPy ONNX interpreter) on the seeded
random-init DeepCNN heads of spec.CNN_CASES.  PARITY UNPINNED like gcn_golden.npz: the reference holds no golden vectors
for Predictor and onnxruntime is absent; these pin the CUDA path and the oracle against regressions only.

    python tests/golden/make_cnn_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mdf_pkg  # noqa: E402

mdf_pkg.load()
from metagenomic_deepfri_b200 import synth  # noqa: E402
import gcn_oracle as go  # noqa: E402

sys.path.insert(0, HERE)
import spec  # noqa: E402


def main():
    g = {}
    with tempfile.TemporaryDirectory() as d:
        for tag, (kw, seed, n, lo, hi) in spec.CNN_CASES.items():
            path = os.path.join(d, tag + ".onnx")
            synth.write_cnn_model(path, synth.CNNConfig(**kw), seed=seed)
            p = go.Predictor(path)
            g[tag + "_scores"] = np.stack([p.forward_pass(s) for s in spec.cnn_sequences(tag)])
    np.savez_compressed(os.path.join(HERE, "cnn_golden.npz"), **g)
    print("cnn_golden.npz:", {k: (v.shape, float(v.min()), float(v.max())) for k, v in g.items()})


if __name__ == "__main__":
    main()
