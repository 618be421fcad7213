#!/usr/bin/env python
"""bench.py — proteins/s of the structure-branch hot path (contact-map build + alignment transfer +
DeepFRI GCN MF forward) on N B200s of one node, one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference ...                     # the reference's CPU path (oracle) on host cores

One "step" = one pass of the whole path over one batch of synthetic proteins.  The default workload is
the configuration BASELINE.json's metric is quoted on, configs[4] (the 1M-protein metagenomic MF job,
L ~ LogNormal(median 250, sigma 0.6) clipped to [50, 1000]), processed the way the sharded job runs it:
in batches of 16,384 proteins per GPU (Markov-gapped alignments, random-walk C-alpha structures, 10 A
contact maps, random-init MF head C=489 in the reference's ONNX layout).  `--workload config0` selects
configs[0] (1,000 proteins, L~U{100..500}), the reference's own CPU-runnable case.  Weak scaling: every
rank processes its own copy of the batch (same seed, so per-GPU work is exactly fixed); no collective on the
compute path, one final gather of the score matrices.

Prints ONE JSON line on rank 0 (contract in the task statement).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import mdf_pkg  # noqa: E402

mdf_pkg.load()
from metagenomic_deepfri_b200 import synth  # noqa: E402

THRESHOLD, GEN = 10.0, 2
WORKLOADS = {
    # name: (description, default proteins per step per GPU, generator)
    "config4": ("BASELINE configs[4]: 1M-protein metagenomic MF job, L~LogNormal(250, 0.6) clipped [50,1000], run in "
                "batches of {n} proteins per GPU per step; gapped alignments, 10A maps, MF head C=489", 16384,
                lambda n, seed: synth.make_workload(n, 50, 1000, seed=seed, dist="lognormal", threshold=THRESHOLD)),
    "config0": ("BASELINE configs[0]: {n} synthetic proteins L~U{{100..500}}, gapped alignments, 10A maps, MF head C=489", 1000,
                lambda n, seed: synth.make_workload(n, 100, 500, seed=seed, threshold=THRESHOLD)),
}


def make_batch(args, rank):
    desc, default_n, gen = WORKLOADS[args.workload]
    n = args.proteins or default_n
    base_seed = 5 if args.workload == "config4" else 1
    # every rank draws the same batch: per-GPU work is exactly fixed as N grows (weak scaling); ranks differ in nothing but
    # the device they run on
    return gen(n, base_seed), desc.format(n=n)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="config4", choices=sorted(WORKLOADS))
    ap.add_argument("--proteins", type=int, default=0, help="proteins per step per GPU (0 = the workload's default)")
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--cpu-sample", type=int, default=400, help="proteins in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-lanes", type=int, default=1, help="concurrent contexts per GPU in the end-to-end leg (1 or 2)")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "tflops_burst": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic(stage):
    """dram bytes (read + write) per launch of the stage's kernel from the committed `ncu --set full` capture of this
    same command (profiles/ncu_traffic.json, written by tools/ncu_summary.py), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(stage)
    return None


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_path_proteins_per_s(wl, model_path, idx):
    """The reference's CPU path on `idx`: compiled contact_map_utils.pyx (oracle/_ref) when present,
    else the C port, + NumPy glue of bio_utils.py:214-223, then the fp32 ONNX interpreter standing in
    for onnxruntime (absent from this image), one protein per call like pipeline.py:301-319."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cmap_oracle as co
    import gcn_oracle as go
    ref = co.ref_module()
    pred = go.Predictor(model_path)
    t0 = time.perf_counter()
    for i in idx:
        c = wl.coords[i]
        if ref is not None:
            D = ref.pairwise_sqeuclidean(c)
            sp = np.argwhere((D < THRESHOLD ** 2).astype(np.int32) == 1).astype(np.int32)
            cm = ref.align_contact_map(wl.gapped_query[i], wl.gapped_target[i], sp, GEN)
        else:
            cm = co.build_align_contact_map(wl.gapped_query[i], wl.gapped_target[i], c, THRESHOLD, GEN)
        pred.forward_pass(wl.query_seqs[i], cm)
    dt = time.perf_counter() - t0
    kind = "port"   # the GCN half is a port (onnxruntime unavailable); the cmap half runs the reference itself when built
    return len(idx) / dt, dt, kind, ("reference .pyx" if ref is not None else "C port")


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [d.get("num_threads", 1) for d in threadpool_info() if d.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl, workload = make_batch(args, 0)
    per_step = 6
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "mf.onnx")
        synth.write_gcn_model(path, synth.GCNConfig())
        rng = np.random.default_rng(0)
        times = []
        for s in range(args.warmup + args.steps):
            idx = rng.choice(len(wl), per_step, replace=False)
            pps, dt, kind, cm_impl = cpu_path_proteins_per_s(wl, path, idx)
            if s >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = per_step * args.steps / total
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": "proteins/sec (GCN MF fwd incl. cmap)", "value": value, "unit": "proteins/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "sample": f"{per_step} proteins per step drawn from one {len(wl)}-protein batch",
                   "cmap": cm_impl, "gcn": "fp32 NumPy ONNX interpreter (onnxruntime absent)"},
        "cpu_baseline": {"value": value, "unit": "proteins/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step * args.steps} proteins, batch=1 per call, host has {os.cpu_count()} cpus"},
        "e2e": {"value": value, "unit": "proteins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from metagenomic_deepfri_b200 import _lib, predict
    torch.cuda.set_device(local_rank)
    stream = torch.cuda.Stream()             # a real (non-NULL) stream shared by torch and the library
    torch.cuda.set_stream(stream)
    ctx = _lib.Context(local_rank, stream=stream.cuda_stream)
    wl, workload = make_batch(args, rank)
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, f"mf_{rank}.onnx")
    synth.write_gcn_model(path, synth.GCNConfig())
    pred = predict.Predictor(path, context=ctx)
    if args.engine != "auto":
        pred.set_engine(args.engine)
    n = len(wl)
    C = pred.n_terms
    T = sum(len(s) for s in wl.query_seqs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    # ---- resident leg: inputs already in HBM when the timed region starts
    batch = pred.upload(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords)
    for _ in range(args.warmup):
        pred.run(batch, THRESHOLD, GEN)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ctx.profile(True)                           # per-stage CUDA events on the launch stream, inside the timed steps (roofline object)
    for s in range(args.steps):
        flush.zero_()                           # L2 flush between timed iterations (outside the events)
        ev[s][0].record(stream)
        pred.run(batch, THRESHOLD, GEN)
        ev[s][1].record(stream)
    barrier()
    timed_stages = ctx.profile_report()
    ctx.profile(False)
    launches = ctx.launch_count - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    scores_resident = pred.fetch_scores(batch)

    # ---- end-to-end leg: pinned host buffers -> H2D -> path -> D2H, through the public call
    inputs = predict.PathInputs(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords, pin=True)
    out = inputs.output_buffer(C)
    for _ in range(args.warmup):
        pred.forward_inputs(inputs, THRESHOLD, GEN, out)
    barrier()
    # the job's one collective: the final result gather.  Source = the pinned score buffer of the last step, destination = a
    # pinned [world, n, C] matrix on rank 0 allocated before the timed region (pageable staging of 8 x 32 MB cost more than a step)
    gathered = torch.empty((world, n, C), dtype=torch.float32, pin_memory=True) if (world > 1 and rank == 0) else None

    def gather_scores(host_scores):
        sd = torch.from_numpy(host_scores).cuda(non_blocking=True)
        gl = [torch.empty_like(sd) for _ in range(world)] if rank == 0 else None
        dist.gather(sd, gl, dst=0)
        if rank != 0:
            return None
        for r in range(world):
            gathered[r].copy_(gl[r], non_blocking=True)
        torch.cuda.synchronize()
        return gathered.numpy()
    if world > 1:
        gather_scores(out)                                      # warm-up: NCCL builds its gather channels lazily
    barrier()
    # Two lanes per GPU: a second context (own stream + workspace) and Predictor in a second host thread, so that one
    # batch's host-side packing, H2D and D2H overlap the other batch's kernels.  Every step still does the full
    # H2D -> path -> D2H of its own batch through the same public call (ctypes releases the GIL).
    lanes = [(pred, inputs, out)]
    if args.e2e_lanes > 1:
        stream2 = torch.cuda.Stream()
        ctx2 = _lib.Context(local_rank, stream=stream2.cuda_stream)
        pred2 = predict.Predictor(path, context=ctx2)
        if args.engine != "auto":
            pred2.set_engine(args.engine)
        inputs2 = predict.PathInputs(wl.query_seqs, wl.gapped_query, wl.gapped_target, wl.coords, pin=True)
        out2 = inputs2.output_buffer(C)
        for _ in range(max(1, args.warmup)):
            pred2.forward_inputs(inputs2, THRESHOLD, GEN, out2)
        lanes.append((pred2, inputs2, out2))
    barrier()

    def lane_worker(lane, nsteps):
        p, i, o = lane
        for _ in range(nsteps):
            p.forward_inputs(i, THRESHOLD, GEN, o)              # synchronous: returns with scores on the host
    split = [args.steps // len(lanes) + (1 if k < args.steps % len(lanes) else 0) for k in range(len(lanes))]
    threads = [threading.Thread(target=lane_worker, args=(lanes[k], split[k])) for k in range(1, len(lanes))]
    t0 = time.perf_counter()
    for th in threads:
        th.start()
    lane_worker(lanes[0], split[0])
    for th in threads:
        th.join()
    local_scores = out                                          # nothing writes the pinned buffer after the last step
    for _, _, o in lanes[1:]:
        assert np.abs(o - local_scores).max() < 1e-5, "the two end-to-end lanes disagree"
    if world > 1:
        all_scores = gather_scores(local_scores)
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    assert np.abs(local_scores - scores_resident).max() < 1e-5, "resident and end-to-end legs disagree"

    # max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    # ---- per-stage profile for the roofline object: the stage events recorded live inside the timed steps, averaged per step
    PROF_PASSES = max(1, args.steps)
    agg = {}
    for name, ms, units in timed_stages:
        a = agg.setdefault(name, [0.0, 0.0, 0])
        a[0] += ms / PROF_PASSES; a[1] += units / PROF_PASSES; a[2] += 1
    for a in agg.values():
        a[2] //= PROF_PASSES
    peaks = measured_peaks()
    stage_total = sum(a[0] for a in agg.values())
    dom = max(agg.items(), key=lambda kv: kv[1][0])
    dname, (dms, dunits, dcount) = dom
    if dname == "cmap_build_transfer":
        roof = {"bound": "hbm", "achieved": dunits / (dms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"}
    else:
        roof = {"bound": "tensor", "achieved": dunits / (dms * 1e-3) / 1e12, "peak": peaks["tflops"], "unit": "TFLOP/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["traffic"] = ncu_traffic(dname)
    roof["kernel"] = dname
    roof["launches_per_step"] = dcount
    roof["share_of_step"] = dms / stage_total if stage_total else None
    roof["peak_source"] = peaks["source"] + (" sustained bf16" if roof["bound"] == "tensor" else "")
    roof["stages_ms"] = {k: round(v[0], 4) for k, v in agg.items()}

    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        idx = np.random.default_rng(0).choice(n, min(args.cpu_sample, n), replace=False)
        pps, dt, kind, cm_impl = cpu_path_proteins_per_s(wl, path, idx)
        cpu = {"value": pps, "unit": "proteins/s", "cores": blas_threads(), "kind": kind,
               "sample": f"{len(idx)} of the {n} proteins of one step ({dt:.1f} s), batch=1 per call; cmap: {cm_impl}; "
                         f"GCN: fp32 NumPy ONNX interpreter standing in for onnxruntime; host has {os.cpu_count()} cpus"}
    total_proteins = n * world * args.steps
    line = {
        "metric": "proteins/sec (GCN MF fwd incl. cmap)", "value": total_proteins / (dev_ms * 1e-3), "unit": "proteins/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if pred.engine == "simt" else "f16 (hi/lo split weights) with f32 accumulate",
        "data": "synthetic",
        "config": {"workload": workload, "proteins_per_step_per_gpu": n, "residues_per_step_per_gpu": T,
                   "threshold_A": THRESHOLD, "generated_contacts": GEN, "engine": pred.engine,
                   "l2": "flushed between timed iterations (256 MiB memset outside the event pairs)",
                   "timing": "per-step CUDA events on the launch stream, summed over steps, max over ranks"},
        "e2e": {"value": total_proteins / (e2e_ms * 1e-3), "unit": "proteins/s",
                "h2d_bytes_per_step": inputs.h2d_bytes, "d2h_bytes_per_step": int(out.nbytes),
                "ms_per_step": e2e_ms / args.steps,
                "lanes_per_gpu": len(lanes),
                "note": "Predictor.forward_inputs: pinned host buffers -> H2D -> all kernels -> D2H scores, wall clock; "
                        "lanes_per_gpu contexts run such calls concurrently so copies overlap kernels"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


def main():
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner) are sent to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        import torch
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
