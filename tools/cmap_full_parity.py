"""BASELINE configs[1] in full: contact-map build + alignment transfer on ALL 100,000 synthetic query/target pairs,
every pair compared bit for bit with the reference's own code (`oracle/_ref` = mDeepFRI/contact_map_utils.pyx compiled
unchanged, + the NumPy glue of bio_utils.py:214-223; the C port when `_ref` is not built).

Runs 6 A and 10 A (generated contacts 2) and mixes in the edge-case families of SURVEY.md §8d: leading / trailing gaps
(5 % each from the Markov generator), '-/-' columns, target structures shorter and longer than the aligned target sequence.

The GPU side goes through the drop-in `bio_utils.build_align_contact_maps(..., packed=True)` (C ABI
`mdf_cmap_build_transfer`) in chunks; the CPU side runs in a process pool forked before CUDA is initialised.  Each
pair's bit-packed map is reduced to a BLAKE2b digest on both sides; a digest mismatch is re-examined bit by bit.

  python tools/cmap_full_parity.py [--pairs 100000] [--out gpurun_out/cmap_full_parity.json]
"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mdf_pkg  # noqa: E402

mdf_pkg.load()
from metagenomic_deepfri_b200 import synth  # noqa: E402

GEN = 2
_WL = None


def mutate_edge_cases(wl, seed=2):
    """Deterministic edge-case families on top of the Markov alignments (which already carry leading / trailing gaps)."""
    rng = np.random.default_rng(seed)
    kinds = {"structure_shorter": 0, "structure_longer": 0, "double_gap_columns": 0}
    for i in range(len(wl)):
        k = i % 50
        if k == 7 and len(wl.coords[i]) > 12:                   # structure shorter than the aligned target sequence
            wl.coords[i] = np.ascontiguousarray(wl.coords[i][:len(wl.coords[i]) - int(rng.integers(1, 10))])
            kinds["structure_shorter"] += 1
        elif k == 13:                                            # structure longer
            extra = synth.random_walk_coords(rng, [int(rng.integers(1, 10))])[0] + wl.coords[i][-1]
            wl.coords[i] = np.ascontiguousarray(np.concatenate([wl.coords[i], extra]).astype(np.float32))
            kinds["structure_longer"] += 1
        elif k == 21:                                            # '-/-' columns (count as query gaps, contact_map_utils.pyx:64-80)
            q, t = wl.gapped_query[i], wl.gapped_target[i]
            for pos in sorted(rng.integers(0, len(q) + 1, size=3), reverse=True):
                q, t = q[:pos] + "-" + q[pos:], t[:pos] + "-" + t[pos:]
            wl.gapped_query[i], wl.gapped_target[i] = q, t
            kinds["double_gap_columns"] += 1
    return kinds


class Aln:
    """Minimal stand-in for mDeepFRI.alignment.AlignmentResult (alignment.py:65-150)."""

    def __init__(self, gq, gt, coords):
        self.target_name, self.gapped_sequence, self.gapped_target, self.coords = "t", gq, gt, coords


def pack_rows(dense):
    L = dense.shape[0]
    rw = ((L + 127) // 128) * 16                                 # bytes per row (rows padded to 128 bits)
    padded = np.zeros((L, rw * 8), np.uint8)
    padded[:, :L] = dense != 0
    return np.packbits(padded, axis=1, bitorder="little")


def _ref_map(i, thr):
    import cmap_oracle as co
    ref = co.ref_module()
    if ref is not None:
        D = ref.pairwise_sqeuclidean(_WL.coords[i])
        sp = np.argwhere((D < thr ** 2).astype(np.int32) == 1).astype(np.int32)           # bio_utils.py:220-223
        return ref.align_contact_map(_WL.gapped_query[i], _WL.gapped_target[i], sp, GEN)
    return co.build_align_contact_map(_WL.gapped_query[i], _WL.gapped_target[i], _WL.coords[i], thr, GEN)


def _worker(args):
    lo, hi, thr = args
    out = []
    for i in range(lo, hi):
        out.append(hashlib.blake2b(pack_rows(_ref_map(i, thr)).tobytes(), digest_size=16).digest())
    return lo, out


def main():
    global _WL
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=100_000)
    ap.add_argument("--chunk", type=int, default=4000)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cmap_full_parity.json"))
    args = ap.parse_args()
    t0 = time.perf_counter()
    wl = synth.config_workload(1, args.pairs / 100_000)
    kinds = mutate_edge_cases(wl)
    lead = sum(q.startswith("-") or t.startswith("-") for q, t in zip(wl.gapped_query, wl.gapped_target))
    trail = sum(q.endswith("-") or t.endswith("-") for q, t in zip(wl.gapped_query, wl.gapped_target))
    kinds.update(leading_gap=lead, trailing_gap=trail)
    _WL = wl
    n = len(wl)
    print(f"{n} pairs generated in {time.perf_counter() - t0:.1f} s; edge cases {kinds}", flush=True)
    import cmap_oracle as co
    cpu_impl = "reference contact_map_utils.pyx (oracle/_ref)" if co.ref_module() is not None else "C port (oracle/cmap_oracle.c)"
    pool = mp.get_context("fork").Pool(args.procs)              # forked BEFORE the CUDA context exists
    results = {}
    try:
        for thr in (6.0, 10.0):
            t1 = time.perf_counter()
            jobs = [(lo, min(n, lo + 250), thr) for lo in range(0, n, 250)]
            pending = pool.imap_unordered(_worker, jobs)
            # GPU side while the pool works
            from metagenomic_deepfri_b200 import bio_utils
            gpu_digest = [None] * n
            tg = 0.0
            for lo in range(0, n, args.chunk):
                hi = min(n, lo + args.chunk)
                alns = [Aln(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i]) for i in range(lo, hi)]
                t2 = time.perf_counter()
                maps = bio_utils.build_align_contact_maps(alns, thr, GEN, packed=True)
                tg += time.perf_counter() - t2
                for k, m in enumerate(maps):
                    gpu_digest[lo + k] = hashlib.blake2b(np.ascontiguousarray(m).view(np.uint8).tobytes(), digest_size=16).digest()
            cpu_digest = [None] * n
            for lo, out in pending:
                cpu_digest[lo:lo + len(out)] = out
            bad = [i for i in range(n) if gpu_digest[i] != cpu_digest[i]]
            bits = 0
            for i in bad[:50]:
                want = _ref_map(i, thr)
                a = Aln(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i])
                got = bio_utils.build_align_contact_map(a, thr, GEN)[1]
                bits += int((got != want).sum())
            cells = int(sum(len(s) ** 2 for s in wl.query_seqs))
            results[f"{thr:g}A"] = {"pairs_compared": n, "mismatching_pairs": len(bad), "differing_cells_in_first_50": bits,
                                    "cells_compared": cells, "gpu_seconds_incl_copies": round(tg, 2),
                                    "wall_seconds": round(time.perf_counter() - t1, 1)}
            print(thr, results[f"{thr:g}A"], flush=True)
    finally:
        pool.close()
    line = {"what": "BASELINE configs[1], every pair, bit-packed maps, GPU (mdf_cmap_build_transfer) vs " + cpu_impl,
            "generated_contacts": GEN, "edge_cases": kinds, "results": results, "cpu_procs": args.procs,
            "mismatches_total": sum(r["mismatching_pairs"] for r in results.values())}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(line, fh, indent=1)
    print(json.dumps(line))
    if line["mismatches_total"]:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
