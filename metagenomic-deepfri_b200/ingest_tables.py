"""Three-letter -> one-letter residue names of biotite's ProteinSequence alphabet (20 standard residues + B, Z, X, U, O, J)."""
THREE_TO_ONE = {
    "ALA": "A", "ARG": "R", "ASN": "N", "ASP": "D", "CYS": "C", "GLN": "Q", "GLU": "E", "GLY": "G", "HIS": "H", "ILE": "I",
    "LEU": "L", "LYS": "K", "MET": "M", "PHE": "F", "PRO": "P", "SER": "S", "THR": "T", "TRP": "W", "TYR": "Y", "VAL": "V",
    "ASX": "B", "GLX": "Z", "UNK": "X", "SEC": "U", "PYL": "O", "XLE": "J",
}
