"""Graph construction on the device -- the step right before the hot path.

`neighbor_list` replaces `neighbor_list_and_relative_vec` (/root/reference/hamgnn/models/base_model.py:87-178, ASE on the
CPU), `generate_graph` replaces `BaseModel.generate_graph` (:237-288) including `find_matching_columns_of_A_in_B`
(:180-233): the directed edges i -> (j, S) with 0 < d < rc_i + rc_j, rc = radius_scale x the OpenMX cutoff radius of the
element (table :25-40), sorted by (i, j, S), their Cartesian shifts, the inverse-edge index, and for every DFT edge of
`data` its position in the internal graph (`matching_edges`).  Unlike the reference, node indices of the k-th crystal of a
batch are offset by the CUMULATIVE atom count (the reference adds only the previous crystal's count, :266-267: a bug for
batches of more than two crystals).  All arithmetic runs in libhamgnn_b200.so (hgb_neighbor_list, hgb_edge_lookup); the
exclusive scan between the count and the fill pass is a torch.cumsum on the device.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import lib as L

# OpenMX PAO cutoff radii in bohr, by atomic number (base_model.py:25-40); elements not listed: 10.0 (DEFAULT_RADIUS)
_SYMBOLS = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd "
            "Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi").split()
_OPENMX = dict(H=6.0, He=8.0, Li=8.0, Be=7.0, B=7.0, C=6.0, N=6.0, O=6.0, F=6.0, Ne=9.0, Na=9.0, Mg=9.0, Al=7.0, Si=7.0, P=7.0, S=7.0,
               Cl=7.0, Ar=9.0, K=10.0, Ca=9.0, Sc=9.0, Ti=7.0, V=6.0, Cr=6.0, Mn=6.0, Fe=5.5, Co=6.0, Ni=6.0, Cu=6.0, Zn=6.0, Ga=7.0,
               Ge=7.0, As=7.0, Se=7.0, Br=7.0, Kr=10.0, Rb=11.0, Sr=10.0, Y=10.0, Zr=7.0, Nb=7.0, Mo=7.0, Tc=7.0, Ru=7.0, Rh=7.0,
               Pd=7.0, Ag=7.0, Cd=7.0, In=7.0, Sn=7.0, Sb=7.0, Te=7.0, I=7.0, Xe=11.0, Cs=12.0, Ba=10.0, La=8.0, Ce=8.0, Pr=8.0,
               Nd=8.0, Pm=8.0, Sm=8.0, Dy=8.0, Ho=8.0, Lu=8.0, Hf=9.0, Ta=7.0, W=7.0, Re=7.0, Os=7.0, Ir=7.0, Pt=7.0, Au=7.0,
               Hg=8.0, Tl=8.0, Pb=8.0, Bi=8.0)
DEFAULT_RADIUS = 10.0
RADIUS_BY_Z = np.full(128, DEFAULT_RADIUS)
for _z, _s in enumerate(_SYMBOLS, start=1):
    if _s in _OPENMX:
        RADIUS_BY_Z[_z] = _OPENMX[_s]


def neighbor_list(z: torch.Tensor, pos: torch.Tensor, cell: torch.Tensor, radius_scale: float = 1.0, pbc=(True, True, True),
                  radius_type: str = "openmx"):
    """One crystal.  -> dict(edge_index [2,E], cell_shift [E,3] int64, nbr_shift [E,3] fp32, offset [N+1]) on pos.device."""
    if radius_type != "openmx":
        raise NotImplementedError(f"radius table '{radius_type}' is not part of the hot path (OpenMX only)")
    L.require_cuda(pos, z)
    dev = pos.device
    n = int(pos.shape[0])
    cell_h = cell.detach().reshape(3, 3).double().cpu().numpy()
    rad_tab = torch.from_numpy(RADIUS_BY_Z * float(radius_scale)).to(dev)
    radius = rad_tab[z.long()].contiguous()
    rmax = 2.0 * float(radius.max()) if n else 0.0
    vol = abs(np.linalg.det(cell_h))
    reps = []
    for a in range(3):
        h = vol / np.linalg.norm(np.cross(cell_h[(a + 1) % 3], cell_h[(a + 2) % 3]))
        reps.append(int(math.ceil(rmax / h)) if pbc[a] else 0)
    pos64 = pos.detach().double().contiguous()
    cell_c = (C.c_double * 9)(*cell_h.reshape(-1).tolist())
    reps_c = (C.c_int32 * 3)(*reps)
    lib, st = L.load(), L.stream_ptr(dev)
    deg = torch.empty(n, dtype=torch.int64, device=dev)
    L.check(lib.hgb_neighbor_list(pos64.data_ptr(), radius.data_ptr(), cell_c, reps_c, n, deg.data_ptr(), None, 0, None, None, None, st),
            "hgb_neighbor_list")
    offset = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    offset[1:] = torch.cumsum(deg, 0)
    E = int(offset[-1])
    edge_index = torch.empty(2, E, dtype=torch.int64, device=dev)
    cell_shift = torch.empty(E, 3, dtype=torch.int64, device=dev)
    nbr_shift = torch.empty(E, 3, dtype=torch.float32, device=dev)
    L.check(lib.hgb_neighbor_list(pos64.data_ptr(), radius.data_ptr(), cell_c, reps_c, n, None, offset.data_ptr(), E, L.ptr(edge_index),
                                  L.ptr(cell_shift), L.ptr(nbr_shift), st), "hgb_neighbor_list")
    return dict(edge_index=edge_index, cell_shift=cell_shift, nbr_shift=nbr_shift, offset=offset)


def edge_lookup(q_index, q_shift, g_index, g_shift, g_offset, inverse: bool = False) -> torch.Tensor:
    """Position of every query edge (src, dst, shift) -- or of its inverse (dst, src, -shift) -- in a graph sorted by
    (src, dst, shift); -1 where absent."""
    L.require_cuda(q_index, g_index)
    nq, ng = int(q_index.shape[1]), int(g_index.shape[1])
    out = torch.empty(nq, dtype=torch.int64, device=q_index.device)
    rc = L.load().hgb_edge_lookup(L.i64c(q_index).data_ptr(), L.i64c(q_shift).data_ptr(), nq, int(inverse), L.i64c(g_index).data_ptr(),
                                  L.i64c(g_shift).data_ptr(), g_offset.data_ptr(), ng, out.data_ptr(), L.stream_ptr(q_index.device))
    L.check(rc, "hgb_edge_lookup")
    return out


def generate_graph(data, radius_scale: float, radius_type: str = "openmx"):
    """Internal message-passing graph of a (batched) `data` + the map of its DFT edges into it (base_model.py:237-288)."""
    z, pos, batch = data["z"], data["pos"], data["batch"]
    cells = data["cell"].reshape(-1, 3, 3)
    counts = torch.bincount(batch, minlength=cells.shape[0]).tolist()
    ei, cs, ns, offs = [], [], [], []
    n0 = e0 = 0
    for k, nk in enumerate(counts):
        g = neighbor_list(z[n0:n0 + nk], pos[n0:n0 + nk], cells[k], radius_scale, radius_type=radius_type)
        ei.append(g["edge_index"] + n0)                 # cumulative offset (the reference adds counts[k-1] only)
        cs.append(g["cell_shift"])
        ns.append(g["nbr_shift"])
        offs.append(g["offset"][:-1] + e0)
        n0 += nk
        e0 += int(g["edge_index"].shape[1])
    edge_index, cell_shift, nbr_shift = torch.cat(ei, dim=1), torch.cat(cs), torch.cat(ns)
    offset = torch.cat(offs + [torch.tensor([e0], dtype=torch.int64, device=pos.device)])
    out = dict(z=z, pos=pos, batch=batch, edge_index=edge_index, cell_shift=cell_shift, nbr_shift=nbr_shift.to(pos.dtype), offset=offset)
    out["inv_edge_idx_global"] = edge_lookup(edge_index, cell_shift, edge_index, cell_shift, offset, inverse=True)
    if "edge_index" in data and "cell_shift" in data:
        m = edge_lookup(data["edge_index"], data["cell_shift"].to(torch.int64), edge_index, cell_shift, offset)
        if bool((m < 0).any()):
            raise AssertionError("Please increase radius_scale factor!")     # base_model.py:191
        out["matching_edges"] = m
    return out
