"""Batched GPU global alignment: cell updates per second on the 16,384 query / target pairs of one
configs[4] chunk (full alignments with traceback, and score-only), against the CPU restatement (one core) and against what the
GCN path consumes per chunk (~68 ms).

  python tools/nw_bench.py [--pairs 16384] [--out gpurun_out/nw_bench.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mdf_pkg  # noqa: E402

mdf_pkg.load()
from metagenomic_deepfri_b200 import alignment, synth, _lib  # noqa: E402
import nw_oracle as nw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=16384)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "nw_bench.json"))
    args = ap.parse_args()
    wl = synth.keyed_workload_parallel(np.arange(args.pairs), 5, min(16, os.cpu_count() or 1))
    targets = [t.replace("-", "") for t in wl.gapped_target]
    cells = float(sum(len(q) * len(t) for q, t in zip(wl.query_seqs, targets)))
    B62 = (alignment.BLOSUM62_ALPHABET, alignment.BLOSUM62)
    res = {"pairs": args.pairs, "cells": cells}
    for full in (True, False):
        times = []
        for _ in range(4):
            t0 = time.perf_counter()
            scores, ops = alignment.nw_align(wl.query_seqs, targets, scoring_matrix=B62, full=full)
            times.append(time.perf_counter() - t0)
        dt = min(times[1:])
        ctx = _lib.default_context()
        ctx.profile(True)
        alignment.nw_align(wl.query_seqs, targets, scoring_matrix=B62, full=full)
        kern_ms = sum(ms for name, ms, _ in ctx.profile_report() if name == "nw_align")
        ctx.profile(False)
        res["full" if full else "score_only"] = {"call_s": dt, "pairs_per_s": args.pairs / dt, "kernels_ms": kern_ms,
                                                 "GCUPS_kernels": cells / (kern_ms * 1e-3) / 1e9, "GCUPS_call": cells / dt / 1e9}
        print(res["full" if full else "score_only"], flush=True)
    idx = np.random.default_rng(0).choice(args.pairs, min(300, args.pairs), replace=False)
    t0 = time.perf_counter()
    for i in idx:
        ws, wo = nw.align(wl.query_seqs[i], targets[i])
        assert ws == int(scores[i])
    dt = time.perf_counter() - t0
    sub = float(sum(len(wl.query_seqs[i]) * len(targets[i]) for i in idx))
    res["cpu_oracle_one_core"] = {"pairs_per_s": len(idx) / dt, "GCUPS": sub / dt / 1e9}
    print(json.dumps(res))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
