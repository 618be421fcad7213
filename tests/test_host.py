"""CPU tests of the host-side logic: ragged batching, bit packing, LPT sharding, the world_size-2
gloo gather, and argument validation that happens before any GPU call."""
import os
import subprocess
import sys

import numpy as np
import pytest

from metagenomic_deepfri_b200 import batching, sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pack_unpack_bits_roundtrip():
    rng = np.random.default_rng(0)
    for L in (1, 31, 32, 33, 127, 128, 129, 300):
        d = (rng.random((L, L)) < 0.1).astype(np.int32)
        p = batching.pack_bits(d)
        assert p.shape == (L, batching.packed_row_words(L)) and p.dtype == np.uint32
        assert np.array_equal(batching.unpack_bits(p, L), d)
    assert batching.unpack_bits(np.zeros((0, 0), np.uint32), 0).shape == (0, 0)


def test_pack_structures_offsets():
    wl = synth.make_workload(7, 3, 50, seed=1)
    ps = batching.pack_structures(wl.gapped_query, wl.gapped_target, wl.coords)
    assert ps.aln_off[-1] == len(ps.q_aln) == len(ps.t_aln)
    assert list(np.diff(ps.seq_off)) == [len(s) for s in wl.query_seqs]
    assert list(np.diff(ps.coord_off)) == [len(c) for c in wl.coords]
    assert ps.coords.dtype == np.float32 and ps.coords.shape == (ps.coord_off[-1], 3)
    assert list(np.diff(ps.packed_off)) == [len(s) * batching.packed_row_words(len(s)) for s in wl.query_seqs]
    with pytest.raises(ValueError):
        batching.pack_structures(["AC-"], ["AC"], [np.zeros((2, 3), np.float32)])
    with pytest.raises(ValueError):
        batching.pack_structures(["AC"], ["AC"], [np.zeros((2, 2), np.float32)])
    empty = batching.pack_structures([], [], [])
    assert empty.n == 0 and empty.coords.shape == (0, 3)


def test_lpt_bins_are_balanced_and_complete():
    rng = np.random.default_rng(5)
    lengths = np.clip(np.exp(rng.normal(np.log(250), 0.6, 5000)), 50, 1000).astype(int)
    for world in (1, 2, 4, 8):
        bins = sharding.lpt_bins(lengths, world)
        allidx = np.sort(np.concatenate(bins))
        assert np.array_equal(allidx, np.arange(len(lengths)))
        assert sharding.imbalance(bins, lengths) < 1.01
        for b in bins:
            assert np.all(np.diff(b) > 0)                  # ascending protein index inside a bin: a random mix of lengths
    chunks = sharding.chunks_by_residues(bins[0], lengths, 20000)
    assert np.array_equal(np.concatenate(chunks), bins[0])
    assert all(lengths[c].sum() <= 20000 or len(c) == 1 for c in chunks)


def test_two_rank_gloo_gather():
    """N>1 path on CPU: two processes, gloo backend, LPT shards, final gather on rank 0."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", script],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_GATHER_OK" in r.stdout


def test_config_workloads_shapes():
    for idx, scale in ((0, 0.01), (1, 0.0002), (2, 0.002), (3, 0.002), (4, 0.00002)):
        wl = synth.config_workload(idx, scale)
        assert len(wl) >= 1 and all(len(q) == len(t) for q, t in zip(wl.gapped_query, wl.gapped_target))
    lens = [len(s) for s in synth.config_workload(3, 0.005).query_seqs]
    assert min(lens) >= 1000 and max(lens) <= 2500


def test_pipeline_chunks_by_residue_budget():
    from metagenomic_deepfri_b200 import pipeline
    assert pipeline._chunks([], 100) == []
    assert pipeline._chunks([50, 60, 10, 200, 5], 100) == [(0, 1), (1, 3), (3, 4), (4, 5)]
    assert pipeline._chunks([10, 10, 10], 1000) == [(0, 3)]


def test_pack_checked_matches_per_sequence_encoding():
    """predict._pack_checked validates a whole batch in one pass with the error behaviour of predict.pyx:17-48."""
    from metagenomic_deepfri_b200 import predict
    seqs = ["ACDE", "", "MKV-X", "W"]
    b, off = predict._pack_checked(seqs)
    assert b == b"ACDEMKV-XW" and off.tolist() == [0, 4, 4, 9, 10] and off.dtype == np.int64
    b, off = predict._pack_checked([])
    assert b == b"" and off.tolist() == [0]
    with pytest.raises(ValueError, match="Invalid character in sequence: J"):
        predict._pack_checked(["ACD", "AJC"])
    with pytest.raises(ValueError, match="Invalid character in sequence: a"):
        predict._pack_checked(["aCD"])
    with pytest.raises(UnicodeEncodeError):
        predict._pack_checked(["ACé"])


def test_cmap_ragged_sizes_only_needs_no_gpu():
    """`mdf_cmap_build_transfer_ragged(packed_out=NULL)`: query lengths and canonical offsets from the gapped query strings, on
    the host (this is what sizes the output arena of `bio_utils.build_align_contact_maps(out=...)`)."""
    import numpy as np
    from metagenomic_deepfri_b200 import _lib
    host = _lib.pyhost()
    fn = _lib.fn_addr("mdf_cmap_build_transfer_ragged")
    gq = ["AC-DE", "----", "", "M" * 129 + "-" * 3 + "K", "A-" * 300]
    rc, poff, soff = host.cmap_ragged(fn, 0, gq, None, None, 36.0, 2, 0, 0)
    assert rc == 0
    lens = [len(q) - q.count("-") for q in gq]
    assert np.frombuffer(soff, np.int64).tolist() == np.concatenate([[0], np.cumsum(lens)]).tolist()
    words = [L * _lib.packed_row_words(L) for L in lens]
    assert np.frombuffer(poff, np.int64).tolist() == np.concatenate([[0], np.cumsum(words)]).tolist()
    with pytest.raises(UnicodeEncodeError):
        host.cmap_ragged(fn, 0, ["AC\u00e9"], None, None, 36.0, 2, 0, 0)
