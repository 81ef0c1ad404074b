"""Real-basis Wigner 3j / Clebsch-Gordan constant tables for the CUDA kernels (host side, numpy fp64).

Replaces `e3nn.o3.wigner_3j` as used by the reference at
/root/reference/hamgnn/physics/Clebsch_Gordan_coefficients.py:22-26 and, implicitly, inside every
`o3.TensorProduct` (hamgnn/nn/message_passing.py:81-96).  Convention:
    C = Re[ einsum('ij,kl,mn,ikn->jlm', Q(l1), Q(l2), conj(Q(l3))^T, CG_su2) ],  C /= ||C||_F
with Q(l) the complex<-real spherical-harmonic change of basis multiplied by (-i)^l.

The kernels never see dense [2l1+1, 2l2+1, 2l3+1] tensors: `cg_nnz` emits the coordinate list
(i, j, k, value) of the ~12.5 % non-zeros, sorted by output component k, which is what gets staged in
shared memory.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Tuple

import numpy as np


def _cg_su2(j1: int, m1: int, j2: int, m2: int, j3: int, m3: int) -> float:
    """<j1 m1; j2 m2 | j3 m3>, Racah's closed form evaluated in exact integer arithmetic."""
    if m1 + m2 != m3 or not (abs(j1 - j2) <= j3 <= j1 + j2):
        return 0.0
    f = math.factorial
    pref_num = (2 * j3 + 1) * f(j1 + j2 - j3) * f(j1 - j2 + j3) * f(-j1 + j2 + j3)
    pref_den = f(j1 + j2 + j3 + 1)
    pref2 = f(j1 + m1) * f(j1 - m1) * f(j2 + m2) * f(j2 - m2) * f(j3 + m3) * f(j3 - m3)
    s = 0.0
    for k in range(0, j1 + j2 - j3 + 1):
        d = [k, j1 + j2 - j3 - k, j1 - m1 - k, j2 + m2 - k, j3 - j2 + m1 + k, j3 - j1 - m2 + k]
        if min(d) < 0:
            continue
        den = 1
        for x in d:
            den *= f(x)
        s += (-1) ** k / den
    return math.sqrt(pref_num / pref_den) * math.sqrt(pref2) * s


def _q_real_to_complex(l: int) -> np.ndarray:
    q = np.zeros((2 * l + 1, 2 * l + 1), dtype=np.complex128)
    r = 1.0 / math.sqrt(2.0)
    for m in range(1, l + 1):
        # rows: complex index, cols: real index (m=-l..l)
        q[l - m, l + m] = r
        q[l - m, l - m] = -1j * r
        q[l + m, l + m] = (-1) ** m * r
        q[l + m, l - m] = 1j * (-1) ** m * r
    q[l, l] = 1.0
    return ((-1j) ** l) * q


@lru_cache(maxsize=None)
def _w3j_cached(l1: int, l2: int, l3: int) -> np.ndarray:
    cg = np.zeros((2 * l1 + 1, 2 * l2 + 1, 2 * l3 + 1))
    for m1 in range(-l1, l1 + 1):
        for m2 in range(-l2, l2 + 1):
            m3 = m1 + m2
            if abs(m3) <= l3:
                cg[l1 + m1, l2 + m2, l3 + m3] = _cg_su2(l1, m1, l2, m2, l3, m3)
    q1, q2, q3 = _q_real_to_complex(l1), _q_real_to_complex(l2), _q_real_to_complex(l3)
    c = np.einsum("ij,kl,mn,ikn->jlm", q1, q2, np.conj(q3.T), cg.astype(np.complex128))
    assert np.abs(c.imag).max() < 1e-9, (l1, l2, l3)
    c = c.real
    c = c / np.linalg.norm(c)
    c[np.abs(c) < 1e-14] = 0.0
    c.setflags(write=False)
    return c


def wigner_3j(l1: int, l2: int, l3: int) -> np.ndarray:
    """Dense real 3j tensor, fp64, Frobenius norm 1 (read-only view of a cached array)."""
    if not (abs(l1 - l2) <= l3 <= l1 + l2):
        raise ValueError(f"({l1},{l2},{l3}) violates the triangle rule")
    return _w3j_cached(l1, l2, l3)


def cg_nnz(l1: int, l2: int, l3: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Coordinate list of the non-zeros of wigner_3j(l1,l2,l3) sorted by (k, i, j)."""
    c = wigner_3j(l1, l2, l3)
    i, j, k = np.nonzero(c)
    order = np.lexsort((j, i, k))
    i, j, k = i[order], j[order], k[order]
    return i.astype(np.int32), j.astype(np.int32), k.astype(np.int32), c[i, j, k].astype(np.float64)


def normalize2mom_const(name: str) -> float:
    """Second-moment normalisation constant of e3nn's `normalize2mom`:
    c = E_{z~N(0,1)}[f(z)^2]^(-1/2).  e3nn estimates it from 1e6 seeded fp64 samples; we reproduce that
    estimator so the constants agree with e3nn to all printed digits (silu 1.6791767924,
    ssp 1.8782046685, tanh 1.5937334473)."""
    return _n2m(name)


@lru_cache(maxsize=None)
def _n2m(name: str) -> float:
    import torch

    gen = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
    if name == "silu":
        f = torch.nn.functional.silu(z)
    elif name == "ssp":
        f = torch.nn.functional.softplus(z) - math.log(2.0)
    elif name == "tanh":
        f = torch.tanh(z)
    else:
        raise ValueError(name)
    return float(f.pow(2).mean().pow(-0.5))


# ------------------------------------------------------------------------------------------------------------
# Wigner-D constants for the edge-aligned ("rotated frame") message kernel.
#
# Basis: the real spherical harmonics the kernels use (csrc/edge_embed.cu): polar axis z, index l+m,
# Y_{l,+m} ~ Re (x+iy)^m, Y_{l,-m} ~ Im (x+iy)^m, component normalisation.  D^l(R) is defined by
# Y_l(R v) = D^l(R) Y_l(v).  A rotation about z by psi is sparse in this basis ("Z(psi)"):
#     Y'_{+m} = cos(m psi) Y_{+m} - sin(m psi) Y_{-m},   Y'_{-m} = sin(m psi) Y_{+m} + cos(m psi) Y_{-m}
# and a rotation about y is  D(R_y(psi)) = J Z(psi) J^T  with the constant matrix J = D(S), S = R_x(-90 deg)
# (S maps z -> y).  The kernels build D(R_y(-theta) R_z(-phi)) per edge from these two pieces.
def real_sh(l: int, v: np.ndarray) -> np.ndarray:
    """Standard real spherical harmonics of unit vectors v[...,3] in fp64, component normalisation (|Y_l|^2 = 2l+1)."""
    v = np.asarray(v, dtype=np.float64)
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    out = np.zeros(v.shape[:-1] + (2 * l + 1,))
    cm, sm = [np.ones_like(x)], [np.zeros_like(x)]
    for m in range(1, l + 1):
        cm.append(cm[-1] * x - sm[-1] * y)
        sm.append(sm[-1] * x + cm[-2] * y)
    for m in range(0, l + 1):
        qmm = 1.0
        for k in range(1, m + 1):
            qmm *= 2 * k - 1
        q2, q1 = None, None
        for ll in range(m, l + 1):
            if ll == m:
                q = qmm * np.ones_like(z)
            elif ll == m + 1:
                q = (2 * m + 1) * z * q1
            else:
                q = ((2 * ll - 1) * z * q1 - (ll + m - 1) * q2) / (ll - m)
            q2, q1 = q1, q
        n = math.sqrt((2 * l + 1) * math.factorial(l - m) / math.factorial(l + m))
        if m == 0:
            out[..., l] = n * q1
        else:
            out[..., l + m] = math.sqrt(2.0) * n * q1 * cm[m]
            out[..., l - m] = math.sqrt(2.0) * n * q1 * sm[m]
    return out


def wigner_D_numeric(l: int, R: np.ndarray) -> np.ndarray:
    """D^l(R) with Y_l(R v) = D Y_l(v), by a least-squares fit on fixed sample directions (fp64, ~1e-14)."""
    rng = np.random.default_rng(12345 + l)
    v = rng.normal(size=(6 * (2 * l + 1) + 8, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    A = real_sh(l, v)                    # [n, d]
    B = real_sh(l, v @ np.asarray(R, dtype=np.float64).T)
    X, *_ = np.linalg.lstsq(A, B, rcond=None)   # A X = B  ->  D = X^T
    return X.T


@lru_cache(maxsize=None)
def wigner_J(l: int) -> np.ndarray:
    """J = D^l(S), S = rotation about x by -90 degrees (z -> y): D(R_y(psi)) = J Z(psi) J^T."""
    S = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, -1.0, 0.0]])
    J = wigner_D_numeric(l, S)
    J[np.abs(J) < 1e-13] = 0.0
    J.setflags(write=False)
    return J


def wigner_offsets(lmax: int):
    """Float offsets of the per-edge D^l blocks (row-major d x d, each block starting on a multiple of 4) and
    the padded row length."""
    offs, cur = [], 0
    for l in range(lmax + 1):
        offs.append(cur)
        cur += ((2 * l + 1) ** 2 + 3) // 4 * 4
    return offs, cur
