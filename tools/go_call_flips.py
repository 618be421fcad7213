"""GO-call identity per BASELINE config (north-star: "identical above-threshold GO calls"; the reference keeps a term iff
score >= 0.1, pipeline.py:698-715).

For a seeded sample of every GCN config the CUDA path (tensor-core engine, whole path from coordinates) is compared with the
CPU restatement of the reference (compiled contact_map_utils.pyx / C port for the maps, the torch-CPU executor of the same
.onnx file for the network).  Reported per config: scores compared, max |difference|, how many oracle scores lie within 1e-3
of the 0.1 threshold (where no non-bit-identical arithmetic can be pinned), how many of THOSE are called differently, and how
many calls differ outside that band (must be 0).

  python tools/go_call_flips.py [--out gpurun_out/go_call_flips.json]
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
from metagenomic_deepfri_b200 import synth, predict  # noqa: E402
import cmap_oracle as co  # noqa: E402
import torch_ref  # noqa: E402

HEADS = {"mf": 489, "bp": 1943, "cc": 320, "ec": 538}
TOL = 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "go_call_flips.json"))
    ap.add_argument("--scale", type=float, default=1.0, help="scales the sample sizes")
    args = ap.parse_args()
    tmp = tempfile.mkdtemp()
    paths = {}
    for h, C in HEADS.items():
        paths[h] = os.path.join(tmp, f"{h}.onnx")
        synth.write_gcn_model(paths[h], synth.GCNConfig(n_terms=C), seed=1234 if h == "mf" else 77 + C)
    cases = [
        ("configs[0] 1k proteins L100-500 MF", synth.config_workload(0, 0.2 * args.scale), ["mf"]),
        ("configs[2] L50-1000 lognormal, MF+BP+CC+EC", synth.config_workload(2, 0.008 * args.scale), list(HEADS)),
        ("configs[3] L1000-2500 MF", synth.config_workload(3, 0.016 * args.scale), ["mf"]),
        ("configs[4] keyed metagenomic MF", synth.keyed_workload(np.arange(0, 16384, max(1, int(64 / args.scale))), 5), ["mf"]),
    ]
    report = []
    for name, wl, heads in cases:
        t0 = time.perf_counter()
        maps = [co.build_align_contact_map(wl.gapped_query[i], wl.gapped_target[i], wl.coords[i], wl.threshold, wl.generated_contacts)
                for i in range(len(wl))]
        row = {"config": name, "proteins": len(wl), "heads": heads, "scores": 0, "max_abs_diff": 0.0, "in_band": 0, "in_band_flips": 0,
               "out_of_band_flips": 0, "calls_oracle": 0}
        for h in heads:
            gpu = predict.Predictor(paths[h])
            cpu = torch_ref.Predictor(paths[h])
            got = gpu.forward_structures(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords, wl.threshold, wl.generated_contacts)
            want = np.stack([cpu.forward_pass(s, m) for s, m in zip(wl.query_seqs, maps)])
            band = np.abs(want - 0.1) <= TOL
            diff = (got >= 0.1) != (want >= 0.1)
            row["scores"] += int(want.size)
            row["max_abs_diff"] = max(row["max_abs_diff"], float(np.abs(got - want).max()))
            row["in_band"] += int(band.sum())
            row["in_band_flips"] += int(diff[band].sum())
            row["out_of_band_flips"] += int(diff[~band].sum())
            row["calls_oracle"] += int((want >= 0.1).sum())
            gpu.close()
        row["seconds"] = round(time.perf_counter() - t0, 1)
        print(json.dumps(row), flush=True)
        report.append(row)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"tolerance": TOL, "threshold": 0.1, "oracle": "reference .pyx / C port maps + torch-CPU executor of the .onnx file",
               "rows": report}, open(args.out, "w"), indent=1)
    if any(r["out_of_band_flips"] or r["max_abs_diff"] > TOL for r in report):
        raise SystemExit(1)


if __name__ == "__main__":
    main()
