"""Fit the sharding cost model c(L) = ALPHA L^2 + BETA L (metagenomic-deepfri_b200/sharding.py) from the per-stage CUDA-event
profile in a bench.py JSON line of the configs[4] workload: the stages whose work grows with L^2 per protein (contact maps, tile
scan, adjacency product) against sum L^2 of the chunk, the rest (LSTM-LM, embedding, X.W, head) against its residue count.

  python tools/fit_cost.py profiles/r02b_bench_config4.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdf_pkg  # noqa: E402

mdf_pkg.load()
from metagenomic_deepfri_b200 import synth  # noqa: E402

QUADRATIC = ("cmap_build_transfer", "adj_tile_scan", "graphconv_adj")


def main():
    d = json.load(open(sys.argv[1]))
    st = d["roofline"]["stages_ms"]
    lens = synth.keyed_lengths(np.arange(200_000), 5)       # the job's length law (protein i = a function of (seed 5, i))
    T = d["config"]["residues_per_step_per_gpu"]
    sumsq = float((lens.astype(np.float64) ** 2).sum()) * T / float(lens.sum())
    quad = sum(v for k, v in st.items() if k in QUADRATIC)
    lin = sum(v for k, v in st.items() if k not in QUADRATIC)
    alpha, beta = quad * 1e-3 / sumsq, lin * 1e-3 / T
    print(f"T = {T:,} residues, sum L^2 = {sumsq:.4g}; quadratic stages {quad:.2f} ms, linear stages {lin:.2f} ms")
    print(f"ALPHA, BETA = {alpha:.3g}, {beta:.4g}")


if __name__ == "__main__":
    main()
