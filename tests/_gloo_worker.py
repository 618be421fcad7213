"""Worker of tests/test_host.py::test_two_rank_gloo_gather (launched by torch.distributed.run)."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mdf_pkg  # noqa: E402

mdf_pkg.load()
from metagenomic_deepfri_b200 import distributed  # noqa: E402

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(0)
lengths = rng.integers(50, 1000, size=301)
C = 7


def fake_forward(idx):      # a deterministic stand-in for Predictor.forward_structures on one shard
    return (np.asarray(idx, np.float32)[:, None] * 10 + np.arange(C, dtype=np.float32)[None, :])


out = distributed.predict_sharded(fake_forward, lengths, C, rank, world, max_residues=20000)
if rank == 0:
    want = fake_forward(np.arange(len(lengths)))
    assert out is not None and np.array_equal(out, want)
    print("GLOO_GATHER_OK")
else:
    assert out is None
dist.destroy_process_group()
