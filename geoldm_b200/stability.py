"""Bond-order stability metric on the GPU.

Mirrors qm9/analyze.py:check_stability (:209-245) and analyze_stability_for_molecules (:323-371) with
qm9/bond_analyze.py (get_bond_order :101-127, geom_predictor :136-146, allowed_bonds :96-99): for every atom pair the
float32 distance (in pm) is compared against the single/double/triple-bond thresholds of the two elements; an atom is
stable when its summed bond order is an allowed valence; a molecule is stable when all its atoms are.  The reference runs
a Python double loop per molecule on the host (about 0.1 s per GEOM molecule); here the whole batch is one launch of
`geoldm_stability` on ragged-packed coordinates (integer / table work, bit-exact with the reference's float32 compares).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

# Typical bond lengths in pm, "A-B-length" triples; symmetric unless listed in _ONE_WAY.  Data as tabulated by the
# reference (bond_analyze.py:1-47, from wiredchemist.com / chemistry-reference.com).
_SINGLE = """H-H-74 H-C-109 H-N-101 H-O-96 H-F-92 H-B-119 H-Si-148 H-P-144 H-As-152 H-S-134 H-Cl-127 H-Br-141 H-I-161
C-C-154 C-N-147 C-O-143 C-F-135 C-Si-185 C-P-184 C-S-182 C-Cl-177 C-Br-194 C-I-214
N-N-145 N-O-140 N-F-136 N-Cl-175 N-Br-214 N-S-168 N-I-222 N-P-177
O-O-148 O-F-142 O-Br-172 O-S-151 O-P-163 O-Si-163 O-Cl-164 O-I-194
F-F-142 F-S-158 F-Si-160 F-Cl-166 F-Br-178 F-P-156 F-I-187
B-Cl-175 Si-Si-233 Si-S-200 Si-Cl-202 Si-Br-215 Si-I-243
Cl-Cl-199 Cl-P-203 Cl-S-207 Cl-Br-214 S-S-204 S-Br-225 S-P-210 S-I-234
Br-Br-228 Br-P-222 P-P-221 I-I-266"""
_DOUBLE = "C-C-134 C-N-129 C-O-120 N-N-125 N-O-121 O-O-121 O-P-150 P-S-186"
_DOUBLE_ONE_WAY = "C-S-160"           # listed for (C, S) only: the lookup is ordered (bond_analyze.py:40-44)
_TRIPLE = "C-C-120 C-N-116 C-O-113 N-N-110"
_MARGINS = (10, 5, 3)                 # bond_analyze.py:93
_ALLOWED = {'H': (1,), 'C': (4,), 'N': (3,), 'O': (2,), 'F': (1,), 'B': (3,), 'Al': (3,), 'Si': (4,), 'P': (3, 5),
            'S': (4,), 'Cl': (1,), 'As': (3,), 'Br': (1,), 'I': (1,), 'Hg': (1, 2), 'Bi': (3, 5)}


def _table(spec, one_way=""):
    out = {}
    for tok in spec.split():
        a, b, v = tok.split("-")
        out[(a, b)] = out[(b, a)] = int(v)
    for tok in one_way.split():
        a, b, v = tok.split("-")
        out[(a, b)] = int(v)
    return out


_B1, _B2, _B3 = _table(_SINGLE), _table(_DOUBLE, _DOUBLE_ONE_WAY), _table(_TRIPLE)


def bond_tables(atom_decoder):
    """(thr [3, T, T] fp32 with -1 where the ordered pair has no entry, allowed [T] uint32 valence bitmasks)."""
    T = len(atom_decoder)
    thr = np.full((3, T, T), -1.0, dtype=np.float32)
    for k, (tab, margin) in enumerate(zip((_B1, _B2, _B3), _MARGINS)):
        for i, a in enumerate(atom_decoder):
            for j, b in enumerate(atom_decoder):
                if (a, b) in tab:
                    thr[k, i, j] = tab[(a, b)] + margin
    allowed = np.zeros(T, dtype=np.uint32)
    for i, a in enumerate(atom_decoder):
        for v in _ALLOWED[a]:
            allowed[i] |= np.uint32(1 << v)
    return thr, allowed


_CACHE = {}


def _device_tables(dataset_info, device):
    key = (tuple(dataset_info['atom_decoder']), str(device))
    if key not in _CACHE:
        thr, allowed = bond_tables(dataset_info['atom_decoder'])
        _CACHE[key] = (torch.from_numpy(thr).to(device), torch.from_numpy(allowed.astype(np.int32)).to(device))
    return _CACHE[key]


def _pair_mode(dataset_info):
    name = dataset_info['name']
    if name in ('qm9', 'qm9_second_half', 'qm9_first_half'):
        return 0          # (type_i, type_j) in index order, every pair must be tabulated
    if name == 'geom':
        return 1          # pair sorted by type index, missing entries mean "no bond"
    raise ValueError(name)


def stability_ragged(x, atom_type, mol_off, dataset_info):
    """x [N,3] fp32 CUDA, atom_type [N] int32, mol_off [B+1] int32 -> (nr_bonds [N] int32, n_stable [B] int32)."""
    if not x.is_cuda:
        raise _lib.GeoldmError("geoldm_b200 has no CPU path: stability inputs must be CUDA tensors")
    dev = x.device
    thr, allowed = _device_tables(dataset_info, dev)
    x = x.contiguous().float()
    atom_type = atom_type.to(torch.int32).contiguous()
    mol_off = mol_off.to(torch.int32).contiguous()
    B, N = mol_off.numel() - 1, x.shape[0]
    nr_bonds = torch.empty(N, dtype=torch.int32, device=dev)
    n_stable = torch.empty(B, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(_lib.lib().geoldm_stability(B, _lib.ptr(mol_off), _lib.ptr(x), _lib.ptr(atom_type),
                                           len(dataset_info['atom_decoder']), _lib.ptr(thr), _lib.ptr(allowed),
                                           _pair_mode(dataset_info), _lib.ptr(nr_bonds), _lib.ptr(n_stable), st),
               "geoldm_stability")
    return nr_bonds, n_stable


def check_stability(positions, atom_type, dataset_info, debug=False, device="cuda"):
    """Single molecule, reference signature (qm9/analyze.py:209): (molecule_stable, nr_stable_bonds, n_atoms)."""
    pos = torch.as_tensor(np.asarray(positions) if not isinstance(positions, torch.Tensor) else positions)
    assert pos.dim() == 2 and pos.shape[1] == 3
    at = torch.as_tensor(np.asarray(atom_type) if not isinstance(atom_type, torch.Tensor) else atom_type)
    n = pos.shape[0]
    off = torch.tensor([0, n], dtype=torch.int32, device=device)
    _, n_stable = stability_ragged(pos.to(device, torch.float32), at.to(device), off, dataset_info)
    k = int(n_stable[0])
    return k == n, k, n


def analyze_stability_for_molecules(molecule_list, dataset_info):
    """qm9/analyze.py:323-371 on the device: {'mol_stable', 'atm_stable'} fractions (RDKit metrics: not available ->
    None, as in the reference when rdkit is missing)."""
    one_hot, x, node_mask = molecule_list['one_hot'], molecule_list['x'], molecule_list['node_mask']
    if not isinstance(one_hot, torch.Tensor):
        raise TypeError("pass padded tensors (the output of sampling.sample)")
    dev = x.device
    bs, n_max = x.shape[0], x.shape[1]
    counts = node_mask.reshape(bs, n_max).sum(1).long()
    keep = torch.arange(n_max, device=dev).unsqueeze(0) < counts.unsqueeze(1)        # the first n atoms, as [:n]
    atom_type = one_hot.argmax(2)[keep]
    pos = x[keep]
    off = torch.zeros(bs + 1, dtype=torch.int32, device=dev)
    off[1:] = counts.cumsum(0)
    _, n_stable = stability_ragged(pos, atom_type, off, dataset_info)
    mol_stable = int((n_stable.long() == counts).sum())
    validity = {'mol_stable': mol_stable / float(bs), 'atm_stable': int(n_stable.sum()) / float(int(counts.sum()))}
    return validity, None
