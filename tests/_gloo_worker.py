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

# the streaming form bench.py uses: chunks of this rank's bin through submit / wait jobs (two in flight), rows in chunk order
class FakeJob:
    def __init__(self, idx, rows):
        self.idx, self.rows, self.done = idx, rows, False

    def wait(self):
        assert not self.done
        self.rows[...] = fake_forward(self.idx)
        self.done = True


mine = distributed.shard_job(lengths, world, max_proteins=40, max_residues=20000)[rank]
assert all(len(c) <= 40 and lengths[c].sum() <= 20000 for c in mine)
local = np.zeros((sum(len(c) for c in mine), C), np.float32)
jobs = []
distributed.stream_chunks(lambda ch, rows: jobs.append(FakeJob(ch, rows)) or jobs[-1], ((c, len(c)) for c in mine), local)
assert all(j.done for j in jobs)
out2 = distributed.gather_scores(np.concatenate(mine), local, len(lengths), C)
if rank == 0:
    assert np.array_equal(out2, fake_forward(np.arange(len(lengths))))
# the node-local shared result matrix: scatter per chunk while "computing", the final gather is a barrier
board = distributed.ScoreBoard(len(lengths), C)
local3 = np.zeros_like(local)
distributed.stream_chunks(lambda ch, rows: FakeJob(ch, rows), ((c, len(c)) for c in mine), local3,
                          on_done=lambda k, rows: board.put(mine[k], rows))
final = board.finish()
if rank == 0:
    assert final is not None and np.array_equal(final, fake_forward(np.arange(len(lengths))))
else:
    assert final is None
board.close()
# contiguous row ranges written in place (the step-contract end-to-end leg of bench.py)
counts = np.cumsum([0] + [37 + r for r in range(world)])
board2 = distributed.ScoreBoard(int(counts[-1]), C)
board2.rows(counts[rank], counts[rank + 1])[...] = fake_forward(np.arange(counts[rank], counts[rank + 1]))
final2 = board2.finish()
if rank == 0:
    assert np.array_equal(final2, fake_forward(np.arange(counts[-1])))
board2.close()
if rank == 0:
    want = fake_forward(np.arange(len(lengths)))
    assert out is not None and np.array_equal(out, want)
    print("GLOO_GATHER_OK")
else:
    assert out is None
dist.destroy_process_group()
