"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the band-energy head.

PARITY UNPINNED like the rest of oracle/ (the reference cannot be imported here and ships no vectors for this path): this file
restates `HamGNNPlusPlusOut.calculate_band_energies` (hamgnn/models/hamgnn_output.py:1675-1996, export_reciprocal_values=False)
step by step in numpy fp64 with explicit loops -- dense [num_k, Na, Na, nao, nao] scatter with phase factors (:1775-1857), atom /
orbital axis swap and reshape (:1886-1891), selection of the defined orbitals (:1896-1904), Cholesky-transformed eigenproblem
(:1911-1928), band gap (:1930-1936), band window (:1938-1955), output packing (:1973-1996)."""
from __future__ import annotations

import math

import numpy as np


def band_energies(onsite_h, offsite_h, son, soff, edge_index, z, batch, node_counts, nbr_shift, k_vecs, nao_max, basis_def,
                  num_valence, band_num_control=None, return_kspace=False):
    onsite_h, offsite_h = np.asarray(onsite_h, np.float64), np.asarray(offsite_h, np.float64)
    son, soff = np.asarray(son, np.float64), np.asarray(soff, np.float64)
    src, dst = np.asarray(edge_index[0]), np.asarray(edge_index[1])
    z, batch = np.asarray(z), np.asarray(batch)
    nbr_shift, k_vecs = np.asarray(nbr_shift, np.float64), np.asarray(k_vecs, np.float64)
    counts = [int(c) for c in node_counts]
    nb, nk, nao = len(counts), k_vecs.shape[1], nao_max
    orb = np.zeros((99, nao), dtype=bool)
    for zz, idx in basis_def.items():
        orb[int(zz), list(idx)] = True
    ecount = np.bincount(batch[src], minlength=nb)
    a0 = e0 = 0
    energies, waves, gaps, hsym, kspace = [], [], [], [], []
    for c in range(nb):
        na, ne = counts[c], int(ecount[c])
        hk = np.zeros((nk, na, na, nao, nao), dtype=np.complex128)
        sk = np.zeros_like(hk)
        for a in range(na):
            hk[:, a, a] += onsite_h[a0 + a].reshape(nao, nao)
            sk[:, a, a] += son[a0 + a].reshape(nao, nao)
        for e in range(e0, e0 + ne):
            i, j = src[e] - a0, dst[e] - a0
            phase = np.exp(2j * np.pi * (k_vecs[c] @ nbr_shift[e]))          # [nk]
            hk[:, i, j] += phase[:, None, None] * offsite_h[e].reshape(nao, nao)
            sk[:, i, j] += phase[:, None, None] * soff[e].reshape(nao, nao)
        hk = np.swapaxes(hk, -2, -3).reshape(nk, na * nao, na * nao)
        sk = np.swapaxes(sk, -2, -3).reshape(nk, na * nao, na * nao)
        keep = orb[z[a0:a0 + na]].reshape(-1)
        hk, sk = hk[:, keep][:, :, keep], sk[:, keep][:, :, keep]
        kspace.append((hk, sk))
        chol = np.linalg.cholesky(sk)
        ci = np.linalg.inv(chol)
        chi = np.linalg.inv(np.conj(np.swapaxes(chol, -1, -2)))
        ht = ci @ hk @ chi
        ev, vec = np.linalg.eigh(ht)
        vec = np.einsum("ijk,ika->iaj", chi, vec)
        nval = sum(int(num_valence[int(zz)]) for zz in z[a0:a0 + na])
        half = math.ceil(nval / 2)
        gaps.append(ev[:, half].min() - ev[:, half - 1].max())
        if band_num_control is not None:
            if isinstance(band_num_control, dict):
                nbands = sum(int(band_num_control.get(int(zz), 0)) for zz in z[a0:a0 + na])
                ev, vec = ev[:, :nbands], vec[:, :nbands, :]
            else:
                win = max(1, int(band_num_control * half)) if isinstance(band_num_control, float) else min(band_num_control, half)
                ev, vec = ev[:, half - win:half + win], vec[:, half - win:half + win, :]
        energies.append(ev.T)
        waves.append(vec.reshape(-1))
        hsym.append(ht.reshape(-1))
        a0 += na
        e0 += ne
    out = (np.concatenate(energies, 0), np.concatenate(waves, 0), np.asarray(gaps), np.concatenate(hsym, 0))
    return out + (kspace,) if return_kspace else out
