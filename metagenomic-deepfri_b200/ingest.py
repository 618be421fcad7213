"""Coordinate ingest in front of the contact-map kernel.

Reference: `extract_calpha_coords` (`pdb.py:130-162`) turns every FoldComp hit into PDB text and parses it with biotite
(`bio_utils.extract_residues_coordinates`, `bio_utils.py:281-302`; selection `bio_utils.py:230-255`: chain "A", atom name "CA",
`hetero == False`).  Here the text is parsed by the library (`csrc/ingest.cu`, column slices with biotite's selection rules,
threads over structures) and - since a structure database is parsed again and again for every query set - kept as a
C-alpha cache: one mmap-able file of float32 [L, 3] blocks with an id hash table, whose lookups hand the batched path
pointers into the mapping.  FoldComp decompression and mmCIF are not handled: the cache is written from PDB text or from
coordinate arrays the caller already holds.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib


def extract_residues_coordinates(structure_string: str, chain: str = "A", filetype: str = "pdb",
                                 substitutions: Optional[Dict[str, str]] = None) -> Tuple[str, np.ndarray]:
    """`bio_utils.py:281-302` for `filetype="pdb"`: (one-letter residues, float32 [n, 3] C-alpha coordinates) of `chain`.
    `substitutions` maps non-standard three-letter names to standard ones before the conversion (the reference applies its
    `bio_utils.substitutions` table, `bio_utils.py:48-190`; pass that dict to reproduce it); a name that is still unknown raises
    ValueError like biotite's ProteinSequence.  ValueError("Chain X not found in structure.") as `bio_utils.py:243-244`."""
    if filetype != "pdb":
        raise NotImplementedError(f"Filetype {filetype} not supported.")
    if len(chain) != 1:
        raise ValueError(f"Chain {chain} not found in structure.")
    text = structure_string.encode("utf-8", "replace") if isinstance(structure_string, str) else bytes(structure_string)
    L = _lib.lib()
    n = C.c_int(0)
    _lib.check(L.mdf_pdb_calpha(text, len(text), chain.encode("ascii"), None, None, None, 0, C.byref(n)))
    coords = np.empty((n.value, 3), np.float32)
    names = np.empty((n.value, 3), np.uint8)
    _lib.check(L.mdf_pdb_calpha(text, len(text), chain.encode("ascii"), coords.ctypes.data, None, names.ctypes.data, n.value, C.byref(n)))
    from .ingest_tables import THREE_TO_ONE
    res = []
    for row in names:
        name = row.tobytes().decode("ascii", "replace").strip()
        name = (substitutions or {}).get(name, name)
        if name not in THREE_TO_ONE:
            raise ValueError(f"'{name}' is not a valid amino acid")
        res.append(THREE_TO_ONE[name])
    return "".join(res), coords


def calpha_from_pdb_texts(texts: Sequence, chain: str = "A", threads: int = 8) -> List[Optional[np.ndarray]]:
    """C-alpha coordinates of many PDB texts at once (the loop of `pdb.py:150-156`), parsed on `threads` host threads.  A text
    without the chain gives None (the reference's callers skip such hits, `pipeline.py:432-444`)."""
    n = len(texts)
    if n == 0:
        return []
    host = _lib.pyhost()
    keep = [t if isinstance(t, (bytes, str)) else bytes(t) for t in texts]
    ptrs, lens32 = host.pointers(keep)
    lens = np.frombuffer(lens32, np.int32).astype(np.int64)
    rows = np.empty(n, np.int32)
    total = C.c_int64(0)
    L = _lib.lib()
    _lib.check(L.mdf_pdb_calpha_batch(n, ptrs, _lib.lp(lens), chain.encode("ascii"), threads, rows.ctypes.data, None, 0, C.byref(total)))
    flat = np.empty((total.value, 3), np.float32)
    _lib.check(L.mdf_pdb_calpha_batch(n, ptrs, _lib.lp(lens), chain.encode("ascii"), threads, rows.ctypes.data, flat.ctypes.data,
                                      total.value, C.byref(total)))
    out, at = [], 0
    for r in rows:
        if r < 0:
            out.append(None)
        else:
            out.append(flat[at:at + r])
            at += int(r)
    return out


def write_cache(path: str, ids: Sequence[str], coords: Sequence[np.ndarray]) -> None:
    """Writes the C-alpha cache `path` (atomically: a temporary file is renamed into place)."""
    n = len(ids)
    if len(coords) != n:
        raise ValueError("ids and coords differ in length")
    arrs = [np.ascontiguousarray(c, np.float32).reshape(-1, 3) for c in coords]
    rows = np.array([a.shape[0] for a in arrs], np.int32)
    cptr = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in arrs])
    idb = [s.encode("utf-8") for s in ids]
    iptr = (C.c_char_p * max(n, 1))(*idb)
    _lib.check(_lib.lib().mdf_coords_cache_create(os.fsencode(path), n, iptr, rows.ctypes.data, cptr))


def write_cache_from_pdb_texts(path: str, ids: Sequence[str], texts: Sequence, chain: str = "A", threads: int = 8) -> List[str]:
    """Parses `texts` and writes the cache; returns the ids that were skipped (chain not found)."""
    coords = calpha_from_pdb_texts(texts, chain, threads)
    kept = [(i, c) for i, c in zip(ids, coords) if c is not None]
    write_cache(path, [i for i, _ in kept], [c for _, c in kept])
    return [i for i, c in zip(ids, coords) if c is None]


class CoordsCache:
    """Read side of the C-alpha cache: `get(ids)` returns float32 [L, 3] views into the mapping (None for unknown ids) - the
    `coords` list `Predictor.submit_structures` / `bio_utils.build_align_contact_maps` take, without a copy."""

    def __init__(self, path: str):
        h = C.c_void_p()
        _lib.check(_lib.lib().mdf_coords_cache_open(os.fsencode(path), C.byref(h)))
        self._h = h
        self.path = path

    def __len__(self) -> int:
        return int(_lib.lib().mdf_coords_cache_size(self._h))

    def ids(self) -> List[str]:
        buf = C.create_string_buffer(4096)
        out = []
        for q in range(len(self)):
            _lib.check(_lib.lib().mdf_coords_cache_entry(self._h, q, buf, len(buf), None))
            out.append(buf.value.decode("utf-8"))
        return out

    def get(self, ids: Sequence[str]) -> List[Optional[np.ndarray]]:
        n = len(ids)
        if n == 0:
            return []
        idb = [s.encode("utf-8") for s in ids]
        iptr = (C.c_char_p * n)(*idb)
        ptrs = (C.c_void_p * n)()
        rows = np.empty(n, np.int32)
        _lib.check(_lib.lib().mdf_coords_cache_lookup(self._h, n, iptr, ptrs, rows.ctypes.data, None))
        out: List[Optional[np.ndarray]] = []
        for p, r in zip(ptrs, rows):
            if r < 0:
                out.append(None)
            elif r == 0:
                out.append(np.zeros((0, 3), np.float32))
            else:
                a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(int(r), 3))
                a.flags.writeable = False
                out.append(a)
        self._keepalive = self          # views borrow the mapping: keep the cache object alive as long as they are used
        return out

    def close(self) -> None:
        if getattr(self, "_h", None):
            _lib.lib().mdf_coords_cache_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
