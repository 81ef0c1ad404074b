"""Graph containers, `graph_data.npz` loading and synthetic crystal graphs.

* `Data` / `Batch`: the small subset of the `torch_geometric.data.Data` / `Batch` surface that the
  reference hot path touches (attribute + item access, `in`, `to_dict`, `.to(device)`), with PyG's
  collate rule -- keys containing "index" are offset by the running node count and concatenated on the
  last dim, everything else is concatenated on dim 0 -- which is why `inv_edge_idx` stays per-graph
  local (/root/reference/hamgnn/models/hamgnn_output.py:2985-2990).
* `load_graph_data_npz`: reads the reference's on-disk format, `np.savez(path, graph={idx: Data})`
  (/root/reference/DFT_interfaces/openmx/graph_data_gen.py:357-380; loader
  /root/reference/hamgnn/data/graph_data.py:110-159).  When torch_geometric is not importable a stub
  module is registered so the pickled `Data` objects can still be revived into our `Data`.
* synthetic crystals: the reference's neighbour rule -- directed edge i->j over all
  periodic images iff 0 < d < r_i + r_j with OpenMX cutoff radii in bohr
  (/root/reference/hamgnn/models/base_model.py:25-40, 146-154) -- with `edge_index[0]` sorted and
  `inv_edge_idx` = index of (j->i, -shift) as in graph_data_gen.py:293-295.
"""
from __future__ import annotations

import math
import pickle
import sys
import types
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

ANG2BOHR = 1.8897261246

# OpenMX PAO cutoff radii (bohr) for the elements the synthetic workloads use
# (/root/reference/hamgnn/models/base_model.py:26-40).
OPENMX_RADII = {1: 6.0, 5: 7.0, 6: 6.0, 7: 6.0, 8: 6.0, 14: 7.0, 16: 7.0, 42: 7.0, 31: 7.0, 33: 7.0, 13: 7.0,
                15: 7.0, 34: 7.0, 52: 7.0, 83: 8.0, 3: 8.0, 9: 6.0, 11: 9.0, 12: 9.0, 17: 7.0, 22: 7.0, 29: 6.0,
                30: 6.0, 32: 7.0, 50: 7.0, 51: 7.0, 53: 7.0, 74: 7.0, 79: 7.0}


class Data:
    """Attribute/item-access bag of tensors (PyG `Data` stand-in)."""

    def __init__(self, **kw):
        object.__setattr__(self, "_store", {})
        for k, v in kw.items():
            self._store[k] = v

    # attribute / item protocol -------------------------------------------------------
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        store = object.__getattribute__(self, "_store")
        if k in store:
            return store[k]
        raise AttributeError(k)

    def __setattr__(self, k, v):
        self._store[k] = v

    def __getitem__(self, k):
        return self._store[k]

    def __setitem__(self, k, v):
        self._store[k] = v

    def __contains__(self, k):
        return k in self._store

    def __delitem__(self, k):
        del self._store[k]

    def keys(self):
        return list(self._store.keys())

    def to_dict(self):
        return dict(self._store)

    def __setstate__(self, state):  # revive pickled torch_geometric Data (>=2.0: {'_store': {...}})
        object.__setattr__(self, "_store", {})
        if isinstance(state, dict):
            st = state.get("_store", state)
            if hasattr(st, "_mapping"):
                st = st._mapping
            if isinstance(st, dict):
                for k, v in st.items():
                    if not k.startswith("_"):
                        self._store[k] = v

    @property
    def num_nodes(self):
        return int(self._store["z"].shape[0]) if "z" in self._store else int(self._store["pos"].shape[0])

    def to(self, device, non_blocking=False):
        for k, v in self._store.items():
            if torch.is_tensor(v):
                self._store[k] = v.to(device, non_blocking=non_blocking)
        return self

    def pin_memory(self):
        for k, v in self._store.items():
            if torch.is_tensor(v):
                self._store[k] = v.pin_memory()
        return self

    def clone(self):
        return Data(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in self._store.items()})

    def __repr__(self):
        parts = [f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={v!r}" for k, v in self._store.items()]
        return f"{type(self).__name__}({', '.join(parts)})"


class Batch(Data):
    @staticmethod
    def from_data_list(graphs: Sequence[Data]) -> "Batch":
        keys = graphs[0].keys()
        out: Dict[str, torch.Tensor] = {}
        offs = np.cumsum([0] + [g.num_nodes for g in graphs])
        for k in keys:
            vals = [g[k] for g in graphs]
            if not torch.is_tensor(vals[0]):
                out[k] = vals
                continue
            if "index" in k:
                out[k] = torch.cat([v + int(o) for v, o in zip(vals, offs)], dim=-1)
            else:
                out[k] = torch.cat(vals, dim=0)
        out["batch"] = torch.cat([torch.full((g.num_nodes,), i, dtype=torch.long) for i, g in enumerate(graphs)])
        out["ptr"] = torch.as_tensor(offs, dtype=torch.long)
        b = Batch(**out)
        object.__setattr__(b, "num_graphs", len(graphs))
        return b


# ------------------------------------------------------------------------------------ npz
def _install_pyg_stub():
    """Register minimal `torch_geometric.data.*` modules so that pickled PyG Data objects unpickle
    into `Data` when PyG itself is absent."""
    try:
        import torch_geometric  # noqa: F401
        return
    except Exception:
        pass
    names = ["torch_geometric", "torch_geometric.data", "torch_geometric.data.data", "torch_geometric.data.storage"]
    for n in names:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)

    class _Storage(dict):
        def __setstate__(self, state):
            m = state.get("_mapping", state) if isinstance(state, dict) else {}
            self.update(m)

        @property
        def _mapping(self):
            return dict(self)

    class _PygData(Data):
        pass

    for n in names[1:3]:
        sys.modules[n].Data = _PygData
    for cls in ("GlobalStorage", "BaseStorage", "NodeStorage", "EdgeStorage"):
        setattr(sys.modules["torch_geometric.data.storage"], cls, _Storage)
    sys.modules["torch_geometric"].data = sys.modules["torch_geometric.data"]


def load_graph_data_npz(path: str) -> List[Data]:
    """Reference loader contract (hamgnn/data/graph_data.py:110-159): `graph` entry holds a dict
    {idx: Data-or-dict}; dict entries carry numpy arrays."""
    _install_pyg_stub()
    raw = np.load(path, allow_pickle=True)["graph"].item()
    out = []
    for g in raw.values():
        if isinstance(g, Data):
            d = Data(**g.to_dict())
        elif isinstance(g, dict):
            d = Data(**{k: (torch.from_numpy(np.asarray(v)) if isinstance(v, np.ndarray) else v) for k, v in g.items()})
        else:  # real PyG Data
            d = Data(**{k: g[k] for k in g.keys()}) if not callable(getattr(g, "keys", None)) else Data(**{k: g[k] for k in g.keys()})
        out.append(d)
    return out


def save_graph_data_npz(path: str, graphs: Sequence[Data]):
    """Writes the dict-of-numpy flavour of the reference format (graph_data.py:149-159 accepts it)."""
    payload = {i: {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in g.to_dict().items()} for i, g in enumerate(graphs)}
    np.savez(path, graph=payload)


# ------------------------------------------------------------------------------------ neighbour lists
def build_graph(z: np.ndarray, pos_bohr: np.ndarray, cell_bohr: np.ndarray, pbc=(True, True, True),
                radius_scale: float = 1.0, nao_max: int = 19, seed: int = 0, with_targets: bool = True,
                dtype=torch.float32, soc: bool = False) -> Data:
    """Reference neighbour rule on a periodic cell; returns a `Data` with the graph_data.npz fields."""
    from scipy.spatial import cKDTree

    z = np.asarray(z, dtype=np.int64)
    pos = np.asarray(pos_bohr, dtype=np.float64)
    cell = np.asarray(cell_bohr, dtype=np.float64)
    n = len(z)
    rad = np.array([OPENMX_RADII[int(a)] for a in z]) * radius_scale
    rmax = 2 * rad.max()
    # number of images needed along each axis
    vol = abs(np.linalg.det(cell))
    heights = [vol / np.linalg.norm(np.cross(cell[(a + 1) % 3], cell[(a + 2) % 3])) for a in range(3)]
    reps = [int(math.ceil(rmax / h)) if p else 0 for h, p in zip(heights, pbc)]
    shifts = np.array([(a, b, c) for a in range(-reps[0], reps[0] + 1) for b in range(-reps[1], reps[1] + 1)
                       for c in range(-reps[2], reps[2] + 1)], dtype=np.int64)
    img_pos = (pos[None, :, :] + (shifts @ cell)[:, None, :]).reshape(-1, 3)
    img_atom = np.tile(np.arange(n), len(shifts))
    img_shift = np.repeat(shifts, n, axis=0)
    tree_img = cKDTree(img_pos)
    tree0 = cKDTree(pos)
    pairs = tree0.query_ball_tree(tree_img, r=rmax)
    src, dst, sh = [], [], []
    for i, nb in enumerate(pairs):
        nb = np.asarray(nb, dtype=np.int64)
        if nb.size == 0:
            continue
        d = np.linalg.norm(img_pos[nb] - pos[i], axis=1)
        j = img_atom[nb]
        keep = (d > 1e-8) & (d < rad[i] + rad[j])
        nb, j = nb[keep], j[keep]
        # deterministic order: by neighbour atom then shift
        key = np.lexsort((img_shift[nb, 2], img_shift[nb, 1], img_shift[nb, 0], j))
        src.append(np.full(len(j), i))
        dst.append(j[key])
        sh.append(img_shift[nb][key])
    src = np.concatenate(src)
    dst = np.concatenate(dst)
    sh = np.concatenate(sh)
    E = len(src)
    # inverse edge: (dst -> src, -shift), found by sorting a composite integer key
    smax = int(np.abs(sh).max()) if E else 0
    base = 2 * smax + 1

    def _key(a, b, s):
        k = a.astype(np.int64) * n + b.astype(np.int64)
        for c in range(3):
            k = k * base + (s[:, c] + smax)
        return k

    key = _key(src, dst, sh)
    order = np.argsort(key, kind="stable")
    pos_in_sorted = np.searchsorted(key[order], _key(dst, src, -sh))
    inv = order[pos_in_sorted].astype(np.int64)
    assert np.array_equal(key[inv], _key(dst, src, -sh)), "graph is not closed under edge inversion"
    nbr_shift = sh.astype(np.float64) @ cell
    d = Data(
        z=torch.from_numpy(z), pos=torch.from_numpy(pos).to(dtype), cell=torch.from_numpy(cell).to(dtype)[None],
        node_counts=torch.tensor([n], dtype=torch.long), edge_index=torch.from_numpy(np.stack([src, dst])),
        inv_edge_idx=torch.from_numpy(inv), nbr_shift=torch.from_numpy(nbr_shift).to(dtype),
        cell_shift=torch.from_numpy(sh), doping_charge=torch.zeros(1, dtype=dtype), total_energy=torch.zeros(1, dtype=dtype))
    if with_targets:
        g = torch.Generator().manual_seed(seed)
        nn2 = nao_max * nao_max

        def sym_on(x):
            m = x.view(-1, nao_max, nao_max)
            return (0.5 * (m + m.transpose(1, 2))).reshape(-1, nn2)

        def sym_off(x):
            m = x.view(-1, nao_max, nao_max)
            return (0.5 * (m + m[d.inv_edge_idx].transpose(1, 2))).reshape(-1, nn2)

        d.Hon0 = sym_on(0.1 * torch.randn(n, nn2, generator=g)).to(dtype)
        d.Hoff0 = sym_off(0.1 * torch.randn(E, nn2, generator=g)).to(dtype)
        d.Hon = sym_on(d.Hon0 + 0.01 * torch.randn(n, nn2, generator=g)).to(dtype)
        d.Hoff = sym_off(d.Hoff0 + 0.01 * torch.randn(E, nn2, generator=g)).to(dtype)
        d.Son = sym_on(torch.rand(n, nn2, generator=g)).to(dtype)
        d.Soff = sym_off(0.1 * torch.rand(E, nn2, generator=g)).to(dtype)
        if soc:
            add_soc_targets(d, nao_max, seed=seed + 1, dtype=dtype)
    return d


def add_soc_targets(d: Data, nao_max: int, seed: int = 0, dtype=torch.float32) -> Data:
    """Synthetic spin-orbit fields of a graph_data.npz entry (graph_data_gen.py SOC branch): Hon/Hoff/Hon0/Hoff0 become
    the real parts of the (2 nao)^2 spin-orbital blocks, iHon/iHoff/iHon0/iHoff0 the imaginary parts (Hermitian with
    the inverse edge), Lon/Loff [*, nao^2, 3] the orbital angular-momentum matrices of the so3 basis and
    Hon_nonsoc/Hoff_nonsoc the spin-less blocks used with add_H_nonsoc."""
    g = torch.Generator().manual_seed(seed)
    n, E = d.z.shape[0], d.edge_index.shape[1]
    M = 2 * nao_max
    inv = d.inv_edge_idx

    def herm(re, im, partner):
        c = torch.complex(re, im).view(-1, M, M)
        o = c if partner is None else c[partner]
        c = 0.5 * (c + o.conj().transpose(1, 2))
        return c.real.reshape(-1, M * M).to(dtype).contiguous(), c.imag.reshape(-1, M * M).to(dtype).contiguous()

    d.Hon_nonsoc, d.Hoff_nonsoc = d.Hon.clone(), d.Hoff.clone()
    d.Hon0, d.iHon0 = herm(0.1 * torch.randn(n, M * M, generator=g), 0.1 * torch.randn(n, M * M, generator=g), None)
    d.Hoff0, d.iHoff0 = herm(0.1 * torch.randn(E, M * M, generator=g), 0.1 * torch.randn(E, M * M, generator=g), inv)
    d.Hon, d.iHon = herm(d.Hon0 + 0.01 * torch.randn(n, M * M, generator=g), d.iHon0 + 0.01 * torch.randn(n, M * M, generator=g), None)
    d.Hoff, d.iHoff = herm(d.Hoff0 + 0.01 * torch.randn(E, M * M, generator=g), d.iHoff0 + 0.01 * torch.randn(E, M * M, generator=g), inv)
    d.Lon = torch.randn(n, nao_max * nao_max, 3, generator=g).to(dtype)
    d.Loff = torch.randn(E, nao_max * nao_max, 3, generator=g).to(dtype)
    return d


def _jitter(pos, sigma_ang, rng):
    return pos + rng.normal(0.0, sigma_ang, size=pos.shape)


def bulk_silicon(rep=(1, 1, 1), seed=0, **kw) -> Data:
    """C1: diamond-structure Si, primitive 2-atom cell a=5.431 A (E=172 at rep=1)."""
    a = 5.431
    prim = 0.5 * a * np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]], dtype=float)
    basis = np.array([[0, 0, 0], [0.25, 0.25, 0.25]]) @ (a * np.eye(3))
    return _supercell([14, 14], basis, prim, rep, seed, **kw)


def diamond_carbon(rep=(2, 2, 2), seed=0, **kw) -> Data:
    a = 3.567
    prim = 0.5 * a * np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]], dtype=float)
    basis = np.array([[0, 0, 0], [0.25, 0.25, 0.25]]) @ (a * np.eye(3))
    return _supercell([6, 6], basis, prim, rep, seed, **kw)


def graphene(rep=(4, 4, 1), seed=0, **kw) -> Data:
    a = 2.46
    prim = np.array([[a, 0, 0], [a / 2, a * math.sqrt(3) / 2, 0], [0, 0, 20.0]])
    basis = np.array([[0, 0, 0], (prim[0] + prim[1]) / 3])
    return _supercell([6, 6], basis, prim, rep, seed, pbc=(True, True, False), **kw)


def mos2_monolayer(rep=(1, 1, 1), seed=0, **kw) -> Data:
    a, dz = 3.16, 1.56
    prim = np.array([[a, 0, 0], [a / 2, a * math.sqrt(3) / 2, 0], [0, 0, 25.0]])
    c = (prim[0] + prim[1]) / 3
    basis = np.array([[0, 0, 0], c + [0, 0, dz], c - [0, 0, dz]])
    return _supercell([42, 16, 16], basis, prim, rep, seed, pbc=(True, True, False), **kw)


def _supercell(zs, basis_ang, prim_ang, rep, seed, pbc=(True, True, True), jitter=0.02, **kw) -> Data:
    rng = np.random.default_rng(seed)
    pos, z = [], []
    for i in range(rep[0]):
        for j in range(rep[1]):
            for k in range(rep[2]):
                t = i * prim_ang[0] + j * prim_ang[1] + k * prim_ang[2]
                pos.append(basis_ang + t)
                z += list(zs)
    pos = _jitter(np.concatenate(pos), jitter, rng)
    cell = np.array([rep[0] * prim_ang[0], rep[1] * prim_ang[1], rep[2] * prim_ang[2]])
    return build_graph(np.array(z), pos * ANG2BOHR, cell * ANG2BOHR, pbc=pbc, seed=seed, **kw)


def twisted_bilayer_graphene(m: int = 28, seed: int = 0, jitter=0.02, interlayer=3.35, **kw) -> Data:
    """C5: commensurate twisted bilayer graphene, twist index m => N = 4(3m^2+3m+1) atoms
    (m=28: N=9748, theta~1.16 deg).  Layer 2 is layer 1 rotated about an AA site."""
    a = 2.46
    a1, a2 = np.array([a, 0.0]), np.array([a / 2, a * math.sqrt(3) / 2])
    t1 = m * a1 + (m + 1) * a2
    t2 = -(m + 1) * a1 + (2 * m + 1) * a2
    t1p = (m + 1) * a1 + m * a2
    theta = math.atan2(t1[1], t1[0]) - math.atan2(t1p[1], t1p[0])
    R = np.array([[math.cos(theta), -math.sin(theta)], [math.sin(theta), math.cos(theta)]])
    n_layer = 2 * (3 * m * m + 3 * m + 1)
    rng_i = np.arange(-3 * m - 3, 3 * m + 4)
    I, J = np.meshgrid(rng_i, rng_i, indexing="ij")
    lat = I.reshape(-1, 1) * a1 + J.reshape(-1, 1) * a2
    pts = np.concatenate([lat, lat + (a1 + a2) / 3])
    T = np.stack([t1, t2])  # rows
    layers = []
    for rot in (np.eye(2), R):
        p = pts @ rot.T
        frac = p @ np.linalg.inv(T)
        frac = np.round(frac, 9)
        keep = np.all((frac >= -1e-7) & (frac < 1 - 1e-7), axis=1)
        q = p[keep]
        assert len(q) == n_layer, (len(q), n_layer)
        layers.append(q)
    rng = np.random.default_rng(seed)
    pos = np.concatenate([np.c_[layers[0], np.full(n_layer, 10.0)], np.c_[layers[1], np.full(n_layer, 10.0 + interlayer)]])
    pos = _jitter(pos, jitter, rng)
    cell = np.array([[t1[0], t1[1], 0], [t2[0], t2[1], 0], [0, 0, 25.0]])
    z = np.full(len(pos), 6)
    return build_graph(z, pos * ANG2BOHR, cell * ANG2BOHR, pbc=(True, True, False), seed=seed, **kw)


def random_mixed_cell(n_atoms=24, species=(1, 6, 7, 8, 14, 16), seed=0, **kw) -> Data:
    """C4-style random periodic cell with mixed Z (density ~ 0.03 atoms/bohr^3 x0.5)."""
    rng = np.random.default_rng(seed)
    L = (n_atoms / 0.012) ** (1 / 3)  # bohr
    cell = np.eye(3) * L + rng.normal(0, 0.05 * L, size=(3, 3))
    # rejection-sample positions with a minimum distance
    pos = []
    while len(pos) < n_atoms:
        p = rng.random(3) @ cell
        if all(np.linalg.norm(p - q) > 2.2 for q in pos):
            pos.append(p)
    z = rng.choice(np.array(species), size=n_atoms)
    return build_graph(z, np.array(pos), cell, seed=seed, **kw)
