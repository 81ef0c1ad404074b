"""Band-energy head of HamGNN_out (hamgnn/models/hamgnn_output.py:1675-1996 `calculate_band_energies`).

For every crystal of the batch: H(k), S(k) on the crystal's defined orbitals (hgb_band_kspace, csrc/band.cu), the generalized
eigenproblem H c = e S c through the Cholesky factor of S (S = L L^H, H~ = L^-1 H L^-H, eigh -- cuSOLVER via torch.linalg, the
same library calls the reference makes), eigenvectors back-transformed with L^-H, the band gap around the half-filled band and
the optional band window.  Returns what the reference returns: (band_energies [sum of bands, num_k], wavefunctions (flat),
band_gaps [n_crystals], H_sym (flat transformed Hamiltonians))."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Sequence, Union

import torch

from . import lib as L

# valence electrons per species of the OpenMX pseudo-atomic orbital sets (hamgnn_output.py:349-364), Z -> count
OPENMX_NUM_VALENCE = {1: 1, 2: 2, 3: 3, 4: 2, 5: 3, 6: 4, 7: 5, 8: 6, 9: 7, 10: 8, 11: 9, 12: 8, 13: 3, 14: 4, 15: 5, 16: 6, 17: 7,
                      18: 8, 19: 9, 20: 10, 21: 11, 22: 12, 23: 13, 24: 14, 25: 15, 26: 16, 27: 17, 28: 18, 29: 19, 30: 20, 31: 13,
                      32: 4, 33: 15, 34: 6, 35: 7, 36: 8, 37: 9, 38: 10, 39: 11, 40: 12, 41: 13, 42: 14, 43: 15, 44: 14, 45: 15,
                      46: 16, 47: 17, 48: 12, 49: 13, 50: 14, 51: 15, 52: 16, 53: 7, 54: 8, 55: 9, 56: 10, 57: 11, 58: 12, 59: 13,
                      60: 14, 61: 15, 62: 16, 66: 20, 67: 21, 71: 11, 72: 12, 73: 13, 74: 12, 75: 15, 76: 14, 77: 15, 78: 16,
                      79: 17, 80: 18, 81: 19, 82: 14, 83: 15}


class BandEnergyHead:
    def __init__(self, nao_max: int, basis_def: Dict[int, Sequence[int]], num_valence: Dict[int, int], num_k: int,
                 band_num_control: Union[None, int, float, Dict[int, int]] = None):
        self.nao_max, self.num_k = int(nao_max), int(num_k)
        self.basis_def = {int(z): list(v) for z, v in basis_def.items()}
        self.num_valence = {int(z): int(v) for z, v in num_valence.items()}
        self.band_num_control = ({int(k): int(v) for k, v in band_num_control.items()} if isinstance(band_num_control, dict)
                                 else band_num_control)
        mask = torch.zeros(99, self.nao_max, dtype=torch.bool)
        for z, orbs in self.basis_def.items():
            mask[z, orbs] = True
        self._orb_mask = mask
        nv = torch.zeros(99, dtype=torch.long)
        for z, c in self.num_valence.items():
            nv[z] = c
        self._nv = nv

    def kspace_tables(self, src_local, dst_local, z):
        """Host side of hgb_band_kspace for one crystal: (orb_index [na * nao] int32, n_orb, seg_ptr, seg_edge, n_segs) -- the compact
        index of every defined orbital and the edges grouped by (i, j) (stable sort: the images of a pair keep their input order)."""
        dev = z.device
        na = z.shape[0]
        defined = self._orb_mask.to(dev)[z]                                        # [na, nao]
        flat = defined.reshape(-1)
        orb_index = torch.where(flat, torch.cumsum(flat.to(torch.int32), 0, dtype=torch.int32) - 1,
                                torch.full_like(flat, -1, dtype=torch.int32)).to(torch.int32).contiguous()
        n_orb = int(flat.sum())
        E = src_local.shape[0]
        key = src_local * na + dst_local
        order = torch.sort(key, stable=True).indices.contiguous()
        ks = key[order]
        starts = torch.ones(E, dtype=torch.bool, device=dev)
        if E > 1:
            starts[1:] = ks[1:] != ks[:-1]
        seg_ptr = torch.cat([torch.nonzero(starts).reshape(-1), torch.tensor([E], device=dev)]).to(torch.int64).contiguous()
        n_segs = int(seg_ptr.numel() - 1) if E > 0 else 0
        return orb_index, n_orb, seg_ptr, order, n_segs

    def kspace(self, hon, hoff, son, soff, src_local, dst_local, nbr_shift, kvec, z):
        """H(k), S(k) [num_k, n_orb, n_orb] complex64 of one crystal (device tensors; src / dst local atom indices)."""
        dev = hon.device
        na, nao = hon.shape[0], self.nao_max
        orb_index, n_orb, seg_ptr, order, n_segs = self.kspace_tables(src_local, dst_local, z)
        nk = kvec.shape[0]
        hk = torch.empty(nk, n_orb, n_orb, 2, device=dev, dtype=torch.float32)
        sk = torch.empty(nk, n_orb, n_orb, 2, device=dev, dtype=torch.float32)
        rc = L.load().hgb_band_kspace(L.f32c(hon).data_ptr(), L.f32c(hoff).data_ptr(), L.f32c(son).data_ptr(), L.f32c(soff).data_ptr(),
                                      na, nao, seg_ptr.data_ptr(), n_segs, order.data_ptr(), L.i64c(src_local).data_ptr(),
                                      L.i64c(dst_local).data_ptr(), L.f32c(nbr_shift).data_ptr(), L.f32c(kvec).data_ptr(), nk,
                                      orb_index.data_ptr(), n_orb, hk.data_ptr(), sk.data_ptr(), L.stream_ptr(dev))
        L.check(rc, "hgb_band_kspace")
        return torch.view_as_complex(hk), torch.view_as_complex(sk)

    def __call__(self, onsite_h, offsite_h, data):
        """`calculate_band_energies(onsite_hamiltonian, offsite_hamiltonian, data)` of the reference; data needs z, batch,
        node_counts, edge_index, nbr_shift, Son, Soff and k_vecs [n_crystals, num_k, 3]."""
        L.require_cuda(onsite_h, offsite_h)
        dev = onsite_h.device
        src, dst = data["edge_index"][0], data["edge_index"][1]
        z, batch = data["z"], data["batch"]
        counts = [int(c) for c in data["node_counts"].tolist()]
        nb = len(counts)
        e_counts = torch.bincount(batch[src], minlength=nb).tolist()
        nv = self._nv.to(dev)[z]
        valence = torch.zeros(nb, dtype=torch.long, device=dev).index_add_(0, batch, nv).tolist()
        bands_fixed = None
        if isinstance(self.band_num_control, dict):
            per = torch.zeros(99, dtype=torch.long)
            for zz, c in self.band_num_control.items():
                per[zz] = c
            bands_fixed = torch.zeros(nb, dtype=torch.long, device=dev).index_add_(0, batch, per.to(dev)[z]).tolist()
        son, soff, shift, kv = data["Son"], data["Soff"], data["nbr_shift"], data["k_vecs"]
        a0 = e0 = 0
        energies, waves, gaps, hsym = [], [], [], []
        for c in range(nb):
            na, ne = counts[c], int(e_counts[c])
            sl_a, sl_e = slice(a0, a0 + na), slice(e0, e0 + ne)
            hk, sk = self.kspace(onsite_h[sl_a], offsite_h[sl_e], son[sl_a], soff[sl_e], (src[sl_e] - a0).contiguous(),
                                 (dst[sl_e] - a0).contiguous(), shift[sl_e], kv[c].to(dev).float(), z[sl_a])
            chol = torch.linalg.cholesky(sk)
            chol_inv = torch.linalg.inv(chol)
            chol_h_inv = torch.linalg.inv(chol.mH)
            ht = torch.bmm(torch.bmm(chol_inv, hk), chol_h_inv)
            ev, vec = torch.linalg.eigh(ht)
            vec = torch.einsum("ijk,ika->iaj", chol_h_inv, vec)
            half = math.ceil(valence[c] / 2)
            gaps.append((ev[:, half].min() - ev[:, half - 1].max()).unsqueeze(0))
            bc = self.band_num_control
            if bc is not None:
                if bands_fixed is not None:
                    ev, vec = ev[:, :bands_fixed[c]], vec[:, :bands_fixed[c], :]
                else:
                    win = max(1, int(bc * half)) if isinstance(bc, float) else min(bc, half)
                    ev, vec = ev[:, half - win:half + win], vec[:, half - win:half + win, :]
            energies.append(ev.transpose(-1, -2))
            waves.append(vec.reshape(-1))
            hsym.append(ht.reshape(-1))
            a0 += na
            e0 += ne
        return torch.cat(energies, 0), torch.cat(waves, 0), torch.cat(gaps, 0), torch.cat(hsym, 0)
