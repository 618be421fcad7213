"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement of the reference's coordinate ingest for PDB text: `extract_residues_coordinates(structure_string, chain,
filetype="pdb")` (`mDeepFRI/bio_utils.py:281-302`) = `load_structure` (`:258-278`: biotite `PDBFile.read(...).get_structure()[0]`)
followed by `get_residues_coordinates` (`:230-255`: chain filter, `atom_name == "CA"`, `hetero == False`, three-letter ->
one-letter through `ProteinSequence`).

The parsing itself lives in a third-party dependency that is absent from this image and from /root/reference: **biotite**
(`pyproject.toml`: `biotite>=1.0`, un-pinned).  Its published behaviour for `PDBFile.get_structure(model=None, altloc="first")`
is restated here with plain string slices:
  * atom records = lines starting with "ATOM" or "HETATM"; `hetero` = the line starts with "HETATM";
  * models are delimited by MODEL / ENDMDL records, `[0]` keeps the first;
  * fixed columns (0-based slices): atom_name = line[12:16].strip(), altloc = line[16], res_name = line[17:20].strip(),
    chain_id = line[21].strip(), res_id = int(line[22:26]), ins_code = line[26].strip(),
    x, y, z = float(line[30:38]), float(line[38:46]), float(line[46:54]) stored as float32;
  * altloc="first": within each residue (a new residue starts when chain_id, res_id, ins_code or res_name changes) only atoms
    without an alternate location, or with the first alternate-location id that occurs in that residue, are kept.
PARITY UNPINNED against biotite itself: the reference's only test of this function (`tests/test_bio_utils.py:18-21`) downloads
a structure from the AlphaFold database at run time (no network here); it pins `sequence[:10]` and a coordinate checksum of a
file that is not in the tree.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

THREE_TO_ONE = {
    "ALA": "A", "ARG": "R", "ASN": "N", "ASP": "D", "CYS": "C", "GLN": "Q", "GLU": "E", "GLY": "G", "HIS": "H", "ILE": "I",
    "LEU": "L", "LYS": "K", "MET": "M", "PHE": "F", "PRO": "P", "SER": "S", "THR": "T", "TRP": "W", "TYR": "Y", "VAL": "V",
    "ASX": "B", "GLX": "Z", "UNK": "X", "SEC": "U", "PYL": "O", "XLE": "J",
}


def first_model_atoms(text: str):
    atoms = []
    for line in text.splitlines():
        if line.startswith("ENDMDL"):
            break
        if (line.startswith("ATOM") or line.startswith("HETATM")) and len(line) >= 54:
            atoms.append(dict(hetero=line.startswith("HETATM"), atom_name=line[12:16].strip(), altloc=line[16], res_name=line[17:20].strip(),
                              chain_id=line[21].strip(), res_id=line[22:26], ins_code=line[26],
                              coord=(float(line[30:38]), float(line[38:46]), float(line[46:54]))))
    return atoms


def filter_first_altloc(atoms):
    out, cur, first = [], None, None
    for a in atoms:
        rid = (a["chain_id"], a["res_id"], a["ins_code"], a["res_name"])
        if rid != cur:
            cur, first = rid, None
        if a["altloc"] != " ":
            if first is None:
                first = a["altloc"]
            if a["altloc"] != first:
                continue
        out.append(a)
    return out


def extract_residues_coordinates(structure_string: str, chain: str = "A", substitutions: Optional[Dict[str, str]] = None
                                 ) -> Tuple[str, np.ndarray]:
    atoms = filter_first_altloc(first_model_atoms(structure_string))
    if chain not in {a["chain_id"] for a in atoms}:
        raise ValueError(f"Chain {chain} not found in structure.")           # bio_utils.py:243-244
    ca = [a for a in atoms if a["chain_id"] == chain and a["atom_name"] == "CA" and not a["hetero"]]      # :246-249
    names = [(substitutions or {}).get(a["res_name"], a["res_name"]) for a in ca]
    for n in names:
        if n not in THREE_TO_ONE:
            raise ValueError(f"'{n}' is not a valid amino acid")
    coords = np.array([a["coord"] for a in ca], dtype=np.float64).astype(np.float32).reshape(-1, 3)
    return "".join(THREE_TO_ONE[n] for n in names), coords
