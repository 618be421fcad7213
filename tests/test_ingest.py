"""CPU tests of the coordinate ingest: the library's PDB parser (`csrc/ingest.cu`, host code behind the C
ABI) against the column-slice restatement of biotite's rules (`oracle/pdb_oracle.py`), and the C-alpha cache.  No GPU needed."""
import os

import numpy as np
import pytest

import pdb_oracle
from metagenomic_deepfri_b200 import ingest, synth


def case(rng, L, **kw):
    seq = "".join(rng.choice(list(synth.AA20), L))
    ca = synth.random_walk_coords(rng, [L])[0] + rng.uniform(-400, 400, 3).astype(np.float32)
    return seq, ca, synth.pdb_text(seq, ca, rng, **kw)


VARIANTS = [dict(), dict(backbone=False), dict(extra_chain=True), dict(hetatm=True), dict(altloc_every=3), dict(models=3),
            dict(crlf=True), dict(first_res=-5), dict(chain="C", extra_chain=True), dict(hetatm=True, altloc_every=2, models=2, extra_chain=True)]


@pytest.mark.parametrize("kw", VARIANTS)
def test_parser_matches_biotite_rules(kw):
    rng = np.random.default_rng(11)
    for L in (1, 7, 120):
        seq, ca, text = case(rng, L, **kw)
        chain = kw.get("chain", "A")
        want_seq, want_xyz = pdb_oracle.extract_residues_coordinates(text, chain)
        got_seq, got_xyz = ingest.extract_residues_coordinates(text, chain, "pdb")
        assert want_seq == seq and got_seq == seq
        assert got_xyz.dtype == np.float32 and got_xyz.shape == (L, 3)
        assert np.array_equal(got_xyz, want_xyz)                                  # bit for bit: float(line[30:38]) -> float32
        assert np.array_equal(got_xyz, np.round(ca.astype(np.float64), 3).astype(np.float32))


def test_parser_edge_cases():
    rng = np.random.default_rng(3)
    seq, ca, text = case(rng, 30)
    with pytest.raises(ValueError, match="Chain B not found in structure."):       # bio_utils.py:243-244
        ingest.extract_residues_coordinates(text, "B", "pdb")
    with pytest.raises(ValueError, match="Chain B not found"):
        pdb_oracle.extract_residues_coordinates(text, "B")
    with pytest.raises(NotImplementedError):
        ingest.extract_residues_coordinates(text, "A", "mmcif")
    # no trailing newline, full-width coordinate fields, ANISOU / REMARK noise
    noisy = "REMARK 1\n" + text.rstrip("\n").replace("END", "ANISOU    1  CA  ALA A   1     1000   1000   1000      0      0      0       C\nEND")
    assert np.array_equal(ingest.extract_residues_coordinates(noisy, "A", "pdb")[1], pdb_oracle.extract_residues_coordinates(noisy, "A")[1])
    wide = text.replace(text.splitlines()[3][30:54], "-123.456-999.9991234.567")
    a, b = ingest.extract_residues_coordinates(wide, "A", "pdb")[1], pdb_oracle.extract_residues_coordinates(wide, "A")[1]
    assert np.array_equal(a, b)
    # a non-standard residue on an ATOM record: rejected like ProteinSequence does, accepted through a substitution table
    odd = text.replace(" ALA A", " MSE A").replace(" GLY A", " MSE A").replace(" LEU A", " MSE A")
    if "MSE" in odd:
        with pytest.raises(ValueError, match="MSE"):
            ingest.extract_residues_coordinates(odd, "A", "pdb")
        s1, _ = ingest.extract_residues_coordinates(odd, "A", "pdb", substitutions={"MSE": "MET"})
        s2, _ = pdb_oracle.extract_residues_coordinates(odd, "A", substitutions={"MSE": "MET"})
        assert s1 == s2 and "M" in s1
    # a structure whose chain holds no C-alpha at all: zero rows, no error (the reference returns an empty array)
    only_n = "\n".join(l for l in text.splitlines() if " CA " not in l[12:17]) + "\n"
    s, xyz = ingest.extract_residues_coordinates(only_n, "A", "pdb")
    assert s == "" and xyz.shape == (0, 3)


def test_batch_parse_and_cache_roundtrip(tmp_path):
    rng = np.random.default_rng(5)
    cases = [case(rng, int(L), **VARIANTS[i % len(VARIANTS)]) for i, L in enumerate(rng.integers(1, 400, 64))]
    texts = [c[2] for c in cases]
    texts[10] = texts[10].replace(" A ", " Q ")                                    # chain A missing -> skipped
    chains = ["C" if VARIANTS[i % len(VARIANTS)].get("chain") == "C" else "A" for i in range(64)]
    got = ingest.calpha_from_pdb_texts([t for t, c in zip(texts, chains) if c == "A"], "A", threads=4)
    k = 0
    for i, (t, c) in enumerate(zip(texts, chains)):
        if c != "A":
            continue
        if i == 10:
            assert got[k] is None
        else:
            assert np.array_equal(got[k], pdb_oracle.extract_residues_coordinates(t, "A")[1])
        k += 1
    ids = [f"AF-P{i:05d}-F1-model_v4" for i in range(64)]
    keep = [i for i in range(64) if chains[i] == "A"]
    path = str(tmp_path / "db.mdfca")
    skipped = ingest.write_cache_from_pdb_texts(path, [ids[i] for i in keep], [texts[i] for i in keep], "A", threads=3)
    assert skipped == [ids[10]]
    cache = ingest.CoordsCache(path)
    assert len(cache) == len(keep) - 1 and cache.ids() == [ids[i] for i in keep if i != 10]
    order = [ids[i] for i in reversed(keep)] + ["not-in-the-database"]
    views = cache.get(order)
    assert views[-1] is None and views[[ids[i] for i in reversed(keep)].index(ids[10])] is None
    for name, v in zip(order[:-1], views[:-1]):
        i = ids.index(name)
        if i != 10:
            assert v.dtype == np.float32 and not v.flags.writeable
            assert np.array_equal(v, pdb_oracle.extract_residues_coordinates(texts[i], "A")[1])
    cache.close()
    # direct write from arrays, empty structures, duplicates, corrupt files
    ingest.write_cache(path, ["a", "b", "c"], [np.zeros((0, 3)), np.arange(6.0).reshape(2, 3), np.ones((1, 3))])
    c2 = ingest.CoordsCache(path)
    a, b, c = c2.get(["a", "b", "c"])
    assert a.shape == (0, 3) and np.array_equal(b, np.arange(6, dtype=np.float32).reshape(2, 3)) and np.array_equal(c, np.ones((1, 3), np.float32))
    c2.close()
    with pytest.raises(ValueError, match="duplicate id"):
        ingest.write_cache(path, ["x", "x"], [np.zeros((1, 3)), np.zeros((1, 3))])
    with pytest.raises(FileNotFoundError):
        ingest.CoordsCache(str(tmp_path / "missing.mdfca"))
    bad = tmp_path / "bad.mdfca"
    bad.write_bytes(b"not a cache" * 20)
    with pytest.raises(RuntimeError):
        ingest.CoordsCache(str(bad))
    empty = str(tmp_path / "empty.mdfca")
    ingest.write_cache(empty, [], [])
    assert len(ingest.CoordsCache(empty)) == 0
