"""Coordinate ingest throughput (host only): PDB text -> C-alpha coordinates with the library's parser on
1..N threads, the C-alpha cache write, and cache lookups (pointer views) - against what the GPU path consumes
(~235 k structures/s per B200).  The reference's biotite path is absent from this image; its published order of magnitude is
milliseconds per structure.

  python tools/ingest_bench.py [--n 20000] [--out gpurun_out/ingest_bench.json]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import mdf_pkg  # noqa: E402

mdf_pkg.load()
from metagenomic_deepfri_b200 import ingest, synth  # noqa: E402
import pdb_oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ingest_bench.json"))
    args = ap.parse_args()
    wl = synth.keyed_workload_parallel(np.arange(args.n), 5, min(16, os.cpu_count() or 1))
    rng = np.random.default_rng(0)
    targets = [t.replace("-", "") for t in wl.gapped_target]
    t0 = time.perf_counter()
    texts = [synth.pdb_text(t, c, rng).encode() for t, c in zip(targets, wl.coords)]
    mb = sum(len(t) for t in texts) / 1e6
    print(f"{args.n} synthetic PDB texts (full backbone), {mb:.0f} MB, written in {time.perf_counter() - t0:.1f} s", flush=True)
    res = {"structures": args.n, "pdb_text_MB": mb, "host_cpus": os.cpu_count(), "parse": {}}
    for th in (1, 4, 16):
        if th > (os.cpu_count() or 1):
            continue
        t0 = time.perf_counter()
        coords = ingest.calpha_from_pdb_texts(texts, "A", threads=th)
        dt = time.perf_counter() - t0
        res["parse"][str(th)] = {"seconds": dt, "structures_per_s": args.n / dt, "MB_per_s": mb / dt}
        print(th, "threads:", res["parse"][str(th)], flush=True)
    k = min(200, args.n)
    t0 = time.perf_counter()
    for t in texts[:k]:
        pdb_oracle.extract_residues_coordinates(t.decode(), "A")
    res["python_restatement_structures_per_s"] = k / (time.perf_counter() - t0)
    path = os.path.join(tempfile.mkdtemp(), "db.mdfca")
    ids = [f"AF-{i:08d}-F1" for i in range(args.n)]
    t0 = time.perf_counter()
    ingest.write_cache(path, ids, coords)
    res["cache_write_s"] = time.perf_counter() - t0
    res["cache_MB"] = os.path.getsize(path) / 1e6
    cache = ingest.CoordsCache(path)
    perm = [ids[i] for i in np.random.default_rng(1).permutation(args.n)]
    t0 = time.perf_counter()
    views = cache.get(perm)
    dt = time.perf_counter() - t0
    res["cache_lookup_structures_per_s"] = args.n / dt
    assert all(v is not None for v in views) and np.array_equal(views[0], coords[ids.index(perm[0])])
    print(json.dumps(res))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
