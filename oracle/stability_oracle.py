"""CPU restatement of the reference's stability metric — TEST INFRASTRUCTURE ONLY (see oracle/geoldm_oracle.py header).

Follows qm9/analyze.py:209-245 (check_stability) and qm9/bond_analyze.py:93-146 (margins, allowed valences,
get_bond_order, geom_predictor) with numpy float32 arithmetic, vectorised over pairs.  Parity PINNED: checked against the
unmodified reference on tests/golden/stability.npz (oracle/make_golden_stability.py).
"""
from __future__ import annotations

import numpy as np

# (first, second, single, double, triple) typical lengths in pm; 0 = not tabulated.  Rows hold for both orders except
# the one-way rows at the end (the reference's double-bond table lists C-S but not S-C).
_ROWS = [
    ("H", "H", 74, 0, 0), ("H", "C", 109, 0, 0), ("H", "N", 101, 0, 0), ("H", "O", 96, 0, 0), ("H", "F", 92, 0, 0),
    ("H", "B", 119, 0, 0), ("H", "Si", 148, 0, 0), ("H", "P", 144, 0, 0), ("H", "As", 152, 0, 0), ("H", "S", 134, 0, 0),
    ("H", "Cl", 127, 0, 0), ("H", "Br", 141, 0, 0), ("H", "I", 161, 0, 0),
    ("C", "C", 154, 134, 120), ("C", "N", 147, 129, 116), ("C", "O", 143, 120, 113), ("C", "F", 135, 0, 0),
    ("C", "Si", 185, 0, 0), ("C", "P", 184, 0, 0), ("C", "S", 182, 0, 0), ("C", "Cl", 177, 0, 0), ("C", "Br", 194, 0, 0),
    ("C", "I", 214, 0, 0),
    ("N", "N", 145, 125, 110), ("N", "O", 140, 121, 0), ("N", "F", 136, 0, 0), ("N", "Cl", 175, 0, 0),
    ("N", "Br", 214, 0, 0), ("N", "S", 168, 0, 0), ("N", "I", 222, 0, 0), ("N", "P", 177, 0, 0),
    ("O", "O", 148, 121, 0), ("O", "F", 142, 0, 0), ("O", "Br", 172, 0, 0), ("O", "S", 151, 0, 0), ("O", "P", 163, 150, 0),
    ("O", "Si", 163, 0, 0), ("O", "Cl", 164, 0, 0), ("O", "I", 194, 0, 0),
    ("F", "F", 142, 0, 0), ("F", "S", 158, 0, 0), ("F", "Si", 160, 0, 0), ("F", "Cl", 166, 0, 0), ("F", "Br", 178, 0, 0),
    ("F", "P", 156, 0, 0), ("F", "I", 187, 0, 0),
    ("B", "Cl", 175, 0, 0), ("Si", "Si", 233, 0, 0), ("Si", "S", 200, 0, 0), ("Si", "Cl", 202, 0, 0),
    ("Si", "Br", 215, 0, 0), ("Si", "I", 243, 0, 0),
    ("Cl", "Cl", 199, 0, 0), ("Cl", "P", 203, 0, 0), ("Cl", "S", 207, 0, 0), ("Cl", "Br", 214, 0, 0),
    ("S", "S", 204, 0, 0), ("S", "Br", 225, 0, 0), ("S", "P", 210, 186, 0), ("S", "I", 234, 0, 0),
    ("Br", "Br", 228, 0, 0), ("Br", "P", 222, 0, 0), ("P", "P", 221, 0, 0), ("I", "I", 266, 0, 0),
]
_ONE_WAY_DOUBLE = [("C", "S", 160)]
MARGINS = (10, 5, 3)
VALENCES = {"H": [1], "C": [4], "N": [3], "O": [2], "F": [1], "B": [3], "Al": [3], "Si": [4], "P": [3, 5], "S": [4],
            "Cl": [1], "As": [3], "Br": [1], "I": [1], "Hg": [1, 2], "Bi": [3, 5]}


def _lengths():
    tab = [{}, {}, {}]
    for a, b, *ls in _ROWS:
        for k, v in enumerate(ls):
            if v:
                tab[k][(a, b)] = tab[k][(b, a)] = v
    for a, b, v in _ONE_WAY_DOUBLE:
        tab[1][(a, b)] = v
    return tab


LENGTHS = _lengths()


def bond_order(first: str, second: str, distance, check_exists: bool) -> int:
    """bond_analyze.py:101-127; ``distance`` is a numpy float32 scalar in Angstrom."""
    d = np.float32(100) * np.float32(distance)
    if (first, second) not in LENGTHS[0]:
        if check_exists:
            return 0
        raise KeyError((first, second))
    if d < np.float32(LENGTHS[0][(first, second)] + MARGINS[0]):
        if (first, second) in LENGTHS[1] and d < np.float32(LENGTHS[1][(first, second)] + MARGINS[1]):
            if (first, second) in LENGTHS[2] and d < np.float32(LENGTHS[2][(first, second)] + MARGINS[2]):
                return 3
            return 2
        return 1
    return 0


def check_stability(positions, atom_type, dataset_info):
    """(molecule_stable, nr_stable_atoms, n_atoms, nr_bonds[n]) for one molecule."""
    pos = np.asarray(positions, dtype=np.float32)
    types = np.asarray(atom_type).astype(np.int64)
    dec = dataset_info["atom_decoder"]
    geom = dataset_info["name"] == "geom"
    n = len(pos)
    bonds = np.zeros(n, dtype=np.int64)
    for i in range(n):
        for j in range(i + 1, n):
            diff = pos[i] - pos[j]
            sq = diff * diff
            dist = np.sqrt(np.float32(np.float32(sq[0] + sq[1]) + sq[2]))
            if geom:
                a, b = sorted([types[i], types[j]])
                order = bond_order(dec[a], dec[b], dist, True)
            else:
                order = bond_order(dec[types[i]], dec[types[j]], dist, False)
            bonds[i] += order
            bonds[j] += order
    stable = sum(int(bonds[i] in VALENCES[dec[types[i]]]) for i in range(n))
    return stable == n, stable, n, bonds
