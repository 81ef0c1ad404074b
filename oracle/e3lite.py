"""ORACLE (test infrastructure, NOT product code) -- minimal CPU restatement of the e3nn 0.5.0
semantics that the HamGNN hot path relies on.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file.  The product path (hamgnn_b200/) never imports anything from oracle/.

PARITY UNPINNED: e3nn (pinned 0.5.0 by /root/reference/HamGNN.yaml:42) is an un-vendored dependency
of the reference and is not installable in the build container; the reference ships no tests, golden
vectors or fixtures.  This file restates e3nn's *published* algorithms:

  * Irreps / Irrep algebra, sort(), simplify()         (e3nn/o3/_irreps.py)
  * wigner_3j via su(2) Clebsch-Gordan + real basis    (e3nn/o3/_wigner.py)
  * spherical harmonics, 'component' normalisation     (e3nn/o3/_spherical_harmonics.py)
  * TensorProduct ('uvw' / 'uvu' instructions)         (e3nn/o3/_tensor_product/_tensor_product.py)
  * Linear                                             (e3nn/o3/_linear.py)
  * FullyConnectedNet, normalize2mom                   (e3nn/nn/_fc.py, e3nn/math/_normalize_activation.py)
  * Gate (sorted "_Sortcut" input layout), Activation  (e3nn/nn/_gate.py)

and is anchored on the reference's call sites (hamgnn/nn/message_passing.py:81-96,
hamgnn/nn/tensor_products.py:25-47, hamgnn/nn/interaction_blocks.py:306-317, ...).
Substitute pins (tests/test_oracle_so3.py): closed-form w3j values, ||Y_l||^2 = 2l+1,
Y_{l+1} ~ +w3j . Y_l Y_1, and rotation/inversion equivariance of every module.

Everything is written for clarity, in the dense "as executed by e3nn" formulation
(einsum 'uvw,ijk,zuvij->zwk' per instruction, materialised mid tensors).
"""
from __future__ import annotations

import math
from fractions import Fraction
from functools import lru_cache
from typing import List, Sequence, Tuple

import torch


# --------------------------------------------------------------------------------------
# Irreps
# --------------------------------------------------------------------------------------
class Irrep(tuple):
    """(l, p) with p in {+1, -1}; tuple ordering => for equal l, odd (p=-1) sorts before even."""

    def __new__(cls, l, p=None):
        if p is None:
            if isinstance(l, Irrep):
                return l
            if isinstance(l, str):
                s = l.strip()
                p = {"e": 1, "o": -1}[s[-1]]
                l = int(s[:-1])
            else:
                l, p = l
        assert l >= 0 and p in (1, -1)
        return super().__new__(cls, (int(l), int(p)))

    @property
    def l(self):
        return self[0]

    @property
    def p(self):
        return self[1]

    @property
    def dim(self):
        return 2 * self[0] + 1

    def __repr__(self):
        return f"{self.l}{'e' if self.p == 1 else 'o'}"

    def __mul__(self, other):
        other = Irrep(other)
        p = self.p * other.p
        return [Irrep(l, p) for l in range(abs(self.l - other.l), self.l + other.l + 1)]


class Irreps(tuple):
    """Tuple of (mul, Irrep)."""

    def __new__(cls, irreps=None):
        if isinstance(irreps, Irreps):
            return super().__new__(cls, irreps)
        out = []
        if irreps is None:
            irreps = []
        if isinstance(irreps, str):
            s = irreps.strip()
            if s:
                for tok in s.split("+"):
                    tok = tok.strip()
                    if "x" in tok:
                        mul, ir = tok.split("x")
                        out.append((int(mul), Irrep(ir)))
                    else:
                        out.append((1, Irrep(tok)))
        else:
            for item in irreps:
                if isinstance(item, Irrep):
                    out.append((1, item))
                elif isinstance(item, str):
                    out.extend(Irreps(item))
                else:
                    mul, ir = item
                    out.append((int(mul), Irrep(ir)))
        return super().__new__(cls, out)

    @property
    def dim(self):
        return sum(mul * ir.dim for mul, ir in self)

    @property
    def num_irreps(self):
        return sum(mul for mul, _ in self)

    @property
    def lmax(self):
        return max(ir.l for _, ir in self)

    def slices(self):
        s, i = [], 0
        for mul, ir in self:
            s.append(slice(i, i + mul * ir.dim))
            i += mul * ir.dim
        return s

    def sort(self):
        """e3nn Irreps.sort(): returns (irreps, p, inv) with p[i_old] = i_new."""
        out = sorted((ir, i, mul) for i, (mul, ir) in enumerate(self))
        inv = tuple(i for _, i, _ in out)
        p = [0] * len(inv)
        for new, old in enumerate(inv):
            p[old] = new
        return Irreps([(mul, ir) for ir, _, mul in out]), tuple(p), inv

    def simplify(self):
        out = []
        for mul, ir in self:
            if out and out[-1][1] == ir:
                out[-1] = (out[-1][0] + mul, ir)
            elif mul > 0:
                out.append((mul, ir))
        return Irreps(out)

    def __add__(self, other):
        return Irreps(tuple(self) + tuple(Irreps(other)))

    def __mul__(self, n):
        return Irreps(tuple(self) * int(n))

    __rmul__ = __mul__

    def __repr__(self):
        return "+".join(f"{mul}x{ir}" for mul, ir in self)


# --------------------------------------------------------------------------------------
# wigner_3j  (e3nn/o3/_wigner.py: _so3_clebsch_gordan, change_basis_real_to_complex)
# --------------------------------------------------------------------------------------
def _su2_cg_coeff(j1, m1, j2, m2, j3, m3) -> float:
    """<j1 m1 j2 m2 | j3 m3> by the Racah formula, exact rational arithmetic under the sqrt."""
    if m3 != m1 + m2:
        return 0.0
    vmin = int(max(-j1 + j2 + m3, -j1 + m1, 0))
    vmax = int(min(j2 + j3 + m1, j3 - j1 + j2, j3 + m3))
    f = math.factorial

    C = Fraction((2 * j3 + 1) * f(j3 + j1 - j2) * f(j3 - j1 + j2) * f(j1 + j2 - j3) * f(j3 + m3) * f(j3 - m3),
                 f(j1 + j2 + j3 + 1) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2))
    S = Fraction(0)
    for v in range(vmin, vmax + 1):
        S += Fraction((-1) ** (v + j2 + m2) * f(j2 + j3 + m1 - v) * f(j1 - m1 + v),
                      f(v) * f(j3 - j1 + j2 - v) * f(j3 + m3 - v) * f(v + j1 - j2 - m3))
    return math.sqrt(float(C)) * float(S)


def _su2_cg(j1, j2, j3) -> torch.Tensor:
    mat = torch.zeros(2 * j1 + 1, 2 * j2 + 1, 2 * j3 + 1, dtype=torch.float64)
    if abs(j1 - j2) <= j3 <= j1 + j2:
        for m1 in range(-j1, j1 + 1):
            for m2 in range(-j2, j2 + 1):
                if abs(m1 + m2) <= j3:
                    mat[j1 + m1, j2 + m2, j3 + m1 + m2] = _su2_cg_coeff(j1, m1, j2, m2, j3, m1 + m2)
    return mat


def _real_to_complex(l) -> torch.Tensor:
    q = torch.zeros(2 * l + 1, 2 * l + 1, dtype=torch.complex128)
    s = 1 / math.sqrt(2)
    for m in range(-l, 0):
        q[l + m, l + abs(m)] = s
        q[l + m, l - abs(m)] = -1j * s
    q[l, l] = 1
    for m in range(1, l + 1):
        q[l + m, l + abs(m)] = (-1) ** m * s
        q[l + m, l - abs(m)] = 1j * (-1) ** m * s
    return (-1j) ** l * q


@lru_cache(maxsize=None)
def _w3j64(l1, l2, l3) -> torch.Tensor:
    Q1, Q2, Q3 = _real_to_complex(l1), _real_to_complex(l2), _real_to_complex(l3)
    C = _su2_cg(l1, l2, l3).to(torch.complex128)
    C = torch.einsum("ij,kl,mn,ikn->jlm", Q1, Q2, torch.conj(Q3.T), C)
    assert torch.all(C.imag.abs() < 1e-9)
    C = C.real
    return C / C.norm()


def wigner_3j(l1, l2, l3, dtype=None) -> torch.Tensor:
    """Real-basis Wigner 3j tensor [2l1+1, 2l2+1, 2l3+1], Frobenius norm 1."""
    assert abs(l1 - l2) <= l3 <= l1 + l2
    return _w3j64(l1, l2, l3).to(dtype or torch.get_default_dtype()).clone()


# --------------------------------------------------------------------------------------
# spherical harmonics (e3nn convention: y is the polar axis, 'component' normalisation raw)
# --------------------------------------------------------------------------------------
def _std_real_sh(l: int, x, y, z):
    """sqrt(4 pi) * standard real spherical harmonics Y_lm (m=-l..l, polar axis z, no Condon-Shortley
    sign in the real form: Y_{1,1} ~ +x) of the UNIT vector (x,y,z).  Closed form through associated
    Legendre recursion; written independently of the product's CUDA recurrence."""
    ct = z
    st = torch.sqrt(torch.clamp(x * x + y * y, min=0))
    phi = torch.atan2(y, x)
    out = []
    # P_l^m(ct) with Condon-Shortley phase removed at the end
    P = {}
    P[(0, 0)] = torch.ones_like(ct)
    for m in range(1, l + 1):
        P[(m, m)] = -(2 * m - 1) * st * P[(m - 1, m - 1)]
    for m in range(0, l):
        P[(m + 1, m)] = (2 * m + 1) * ct * P[(m, m)]
    for m in range(0, l + 1):
        for ll in range(m + 2, l + 1):
            P[(ll, m)] = ((2 * ll - 1) * ct * P[(ll - 1, m)] - (ll + m - 1) * P[(ll - 2, m)]) / (ll - m)
    for m in range(-l, l + 1):
        am = abs(m)
        N = math.sqrt((2 * l + 1) * math.factorial(l - am) / math.factorial(l + am))
        base = N * P[(l, am)] * (-1) ** am  # remove the CS phase
        if m < 0:
            out.append(math.sqrt(2) * base * torch.sin(am * phi))
        elif m == 0:
            out.append(base)
        else:
            out.append(math.sqrt(2) * base * torch.cos(am * phi))
    return torch.stack(out, dim=-1)


def spherical_harmonics(ls: Sequence[int], vec: torch.Tensor, normalize: bool = True,
                        normalization: str = "component") -> torch.Tensor:
    """e3nn o3.spherical_harmonics(ls, vec, normalize, normalization).

    e3nn's (x, y, z) input has y as the polar axis: e3nn SH(x,y,z) == sqrt(4pi) * standard real SH
    evaluated at the standard vector (z, x, y)  (SURVEY.md Appendix A.3).
    """
    if normalize:
        vec = torch.nn.functional.normalize(vec, dim=-1)
    ex, ey, ez = vec[..., 0], vec[..., 1], vec[..., 2]
    sx, sy, sz = ez, ex, ey
    outs = []
    for l in ls:
        y = _std_real_sh(l, sx, sy, sz)
        if normalization == "integral":
            y = y / math.sqrt(4 * math.pi)
        elif normalization == "norm":
            y = y / math.sqrt(2 * l + 1)
        else:
            assert normalization == "component"
        outs.append(y)
    return torch.cat(outs, dim=-1)


# --------------------------------------------------------------------------------------
# normalize2mom / FullyConnectedNet
# --------------------------------------------------------------------------------------
def shifted_softplus(x):
    """hamgnn/toolbox/nequip/nn/nonlinearities.py:6"""
    return torch.nn.functional.softplus(x) - math.log(2.0)


@lru_cache(maxsize=None)
def _second_moment_const(name: str) -> float:
    f = {"silu": torch.nn.functional.silu, "ssp": shifted_softplus, "tanh": torch.tanh, "abs": torch.abs}[name]
    gen = torch.Generator(device="cpu").manual_seed(0)
    z = torch.randn(1_000_000, generator=gen, dtype=torch.float64)
    return float(f(z).pow(2).mean().pow(-0.5))


class Normalize2Mom:
    def __init__(self, name: str):
        self.name = name
        self.f = {"silu": torch.nn.functional.silu, "ssp": shifted_softplus, "tanh": torch.tanh, "abs": torch.abs}[name]
        self.cst = _second_moment_const(name)

    def __call__(self, x):
        return self.f(x) * self.cst


class FullyConnectedNet(torch.nn.Module):
    """e3nn.nn.FullyConnectedNet(hs, act): layer{i}.weight [h_in,h_out] ~ randn, no bias,
    hidden: act(x @ W / sqrt(h_in)), last: x @ W / sqrt(h_in)."""

    def __init__(self, hs: List[int], act: str = "silu"):
        super().__init__()
        self.hs = list(hs)
        self.act = Normalize2Mom(act)
        for i, (h1, h2) in enumerate(zip(hs, hs[1:])):
            layer = torch.nn.Module()
            layer.weight = torch.nn.Parameter(torch.randn(h1, h2))
            self.add_module(f"layer{i}", layer)

    def forward(self, x):
        n = len(self.hs) - 1
        for i in range(n):
            w = getattr(self, f"layer{i}").weight
            x = x @ (w / math.sqrt(self.hs[i]))
            if i < n - 1:
                x = self.act(x)
        return x


# --------------------------------------------------------------------------------------
# Linear
# --------------------------------------------------------------------------------------
class Linear(torch.nn.Module):
    """e3nn o3.Linear(irreps_in, irreps_out): one [mul_in, mul_out] block per (i_in, i_out) pair of equal
    irrep (loop i_in outer, i_out inner), flat `weight` in that order, path weight 1/sqrt(sum of mul_in
    over all blocks feeding i_out)."""

    def __init__(self, irreps_in, irreps_out):
        super().__init__()
        self.irreps_in = Irreps(irreps_in)
        self.irreps_out = Irreps(irreps_out)
        self.instr = [(i, o) for i, (_, ii) in enumerate(self.irreps_in)
                      for o, (_, io) in enumerate(self.irreps_out) if ii == io]
        self.weight_numel = sum(self.irreps_in[i][0] * self.irreps_out[o][0] for i, o in self.instr)
        self.weight = torch.nn.Parameter(torch.randn(self.weight_numel))
        self._fan = {}
        for i, o in self.instr:
            self._fan[o] = self._fan.get(o, 0) + self.irreps_in[i][0]

    def forward(self, x):
        sin, sout = self.irreps_in.slices(), self.irreps_out.slices()
        out = [None] * len(self.irreps_out)
        off = 0
        for i, o in self.instr:
            mi, ir = self.irreps_in[i]
            mo, _ = self.irreps_out[o]
            w = self.weight[off:off + mi * mo].view(mi, mo)
            off += mi * mo
            fan = self._fan[o]
            xi = x[:, sin[i]].reshape(-1, mi, ir.dim)
            y = torch.einsum("uw,zui->zwi", w, xi) / math.sqrt(fan)
            out[o] = y if out[o] is None else out[o] + y
        res = []
        for o, (mo, ir) in enumerate(self.irreps_out):
            if out[o] is None:
                res.append(x.new_zeros(x.shape[0], mo * ir.dim))
            else:
                res.append(out[o].reshape(x.shape[0], mo * ir.dim))
        return torch.cat(res, dim=1)


# --------------------------------------------------------------------------------------
# TensorProduct
# --------------------------------------------------------------------------------------
class TensorProduct(torch.nn.Module):
    """e3nn o3.TensorProduct with irrep_normalization='component', path_normalization='element'.
    instructions: (i1, i2, io, mode in {'uvw','uvu'}, has_weight).  shared/internal weights => flat
    parameter `weight`; otherwise weight [z, weight_numel] is passed to forward."""

    def __init__(self, irreps_in1, irreps_in2, irreps_out, instructions, internal_weights=True,
                 shared_weights=True):
        super().__init__()
        self.irreps_in1, self.irreps_in2, self.irreps_out = Irreps(irreps_in1), Irreps(irreps_in2), Irreps(irreps_out)
        self.instructions = [tuple(ins) for ins in instructions]
        self.internal = internal_weights
        self.shapes = []
        for i1, i2, io, mode, hw in self.instructions:
            m1, m2, mo = self.irreps_in1[i1][0], self.irreps_in2[i2][0], self.irreps_out[io][0]
            if not hw:
                self.shapes.append(None)
            elif mode == "uvw":
                self.shapes.append((m1, m2, mo))
            elif mode == "uvu":
                assert mo == m1
                self.shapes.append((m1, m2))
            else:
                raise NotImplementedError(mode)
        self.weight_numel = sum(math.prod(s) for s in self.shapes if s is not None)
        if internal_weights:
            assert shared_weights
            self.weight = torch.nn.Parameter(torch.randn(self.weight_numel))

    def _coef(self, k):
        i1, i2, io, mode, hw = self.instructions[k]
        ne = lambda ins: (self.irreps_in1[ins[0]][0] * self.irreps_in2[ins[1]][0]) if ins[3] == "uvw" \
            else self.irreps_in2[ins[1]][0]
        x = sum(ne(ins) for ins in self.instructions if ins[2] == io)
        alpha = self.irreps_out[io][1].dim / x
        return math.sqrt(alpha)

    def forward(self, x1, x2, weight=None):
        if self.internal:
            weight = self.weight
        s1, s2 = self.irreps_in1.slices(), self.irreps_in2.slices()
        Z = x1.shape[0]
        outs = [None] * len(self.irreps_out)
        off = 0
        for k, (i1, i2, io, mode, hw) in enumerate(self.instructions):
            m1, ir1 = self.irreps_in1[i1]
            m2, ir2 = self.irreps_in2[i2]
            mo, iro = self.irreps_out[io]
            a = x1[:, s1[i1]].reshape(Z, m1, ir1.dim)
            b = x2[:, s2[i2]].reshape(Z, m2, ir2.dim)
            w3 = wigner_3j(ir1.l, ir2.l, iro.l, dtype=x1.dtype).to(x1.device)
            w = None
            if hw:
                n = math.prod(self.shapes[k])
                if weight.dim() == 1:
                    w = weight[off:off + n].view(self.shapes[k])
                else:
                    w = weight[:, off:off + n].reshape((Z,) + self.shapes[k])
                off += n
            if mode == "uvw" and m2 == 1 and w.dim() == 3:
                # same contraction order opt_einsum picks for e3nn's 'uvw,ijk,zuvij->zwk' (dense CG first,
                # then the shared-weight GEMM), written as two matmuls
                xx = (a[:, :, :, None] * b[:, None, 0, None, :]).reshape(Z * m1, ir1.dim * ir2.dim)
                t = (xx @ w3.reshape(ir1.dim * ir2.dim, iro.dim)).reshape(Z, m1, iro.dim)
                y = torch.matmul(t.transpose(1, 2), w[:, 0, :]).transpose(1, 2)
            else:
                xx = torch.einsum("zui,zvj->zuvij", a, b)
                if mode == "uvw":
                    if w.dim() == 3:
                        y = torch.einsum("uvw,ijk,zuvij->zwk", w, w3, xx)
                    else:
                        y = torch.einsum("zuvw,ijk,zuvij->zwk", w, w3, xx)
                else:  # uvu
                    if w is None:
                        y = torch.einsum("ijk,zuvij->zuk", w3, xx)
                    elif w.dim() == 2:
                        y = torch.einsum("uv,ijk,zuvij->zuk", w, w3, xx)
                    else:
                        y = torch.einsum("zuv,ijk,zuvij->zuk", w, w3, xx)
            y = self._coef(k) * y
            outs[io] = y if outs[io] is None else outs[io] + y
        res = []
        for io, (mo, iro) in enumerate(self.irreps_out):
            if outs[io] is None:
                res.append(x1.new_zeros(Z, mo * iro.dim))
            else:
                res.append(outs[io].reshape(Z, mo * iro.dim))
        return torch.cat(res, dim=1)


# --------------------------------------------------------------------------------------
# Gate
# --------------------------------------------------------------------------------------
class Gate(torch.nn.Module):
    """e3nn.nn.Gate(irreps_scalars, act_scalars, irreps_gates, act_gates, irreps_gated).

    Input layout = e3nn's `_Sortcut`: the concatenation (scalars + gates + gated), each simplified,
    is SORTED by (l, p, position) and simplified -- for HamGNN's ResidualBlock that is
    `64x0o + 199x0e(64 scalars | 135 gates) + gated...`.  Output = act(scalars) + gated * act(gates).
    [e3nn-recall; if e3nn keeps [scalars|gates|gated] unsorted instead, only the flat layout of the
    preceding o3.Linear weight changes, not the function class.]
    """

    def __init__(self, irreps_scalars, act_scalars, irreps_gates, act_gates, irreps_gated):
        super().__init__()
        self.irreps_scalars = Irreps(irreps_scalars).simplify()
        self.irreps_gates = Irreps(irreps_gates).simplify()
        self.irreps_gated = Irreps(irreps_gated).simplify()
        assert self.irreps_gates.num_irreps == self.irreps_gated.num_irreps
        self.act_scalars = [Normalize2Mom(a) for a in act_scalars]
        self.act_gates = [Normalize2Mom(a) for a in act_gates]
        cat = self.irreps_scalars + self.irreps_gates + self.irreps_gated
        sorted_irreps, p, inv = cat.sort()
        self._sorted = sorted_irreps
        self._p = p
        self.irreps_in = sorted_irreps.simplify()
        self.irreps_out = self.irreps_scalars + self.irreps_gated
        self._n = (len(self.irreps_scalars), len(self.irreps_gates), len(self.irreps_gated))

    def forward(self, x):
        sl = self._sorted.slices()
        ns, ng, nd = self._n
        pieces = [x[:, sl[self._p[i]]] for i in range(ns + ng + nd)]
        scal = [self.act_scalars[i](pieces[i]) for i in range(ns)]
        gates = torch.cat([self.act_gates[i](pieces[ns + i]) for i in range(ng)], dim=1)
        out = list(scal)
        g0 = 0
        for i, (mul, ir) in enumerate(self.irreps_gated):
            v = pieces[ns + ng + i].reshape(-1, mul, ir.dim)
            out.append((v * gates[:, g0:g0 + mul, None]).reshape(x.shape[0], -1))
            g0 += mul
        return torch.cat(out, dim=1)
