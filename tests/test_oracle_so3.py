"""Substitute pins for the (unpinned) oracle: closed-form values and group-theoretic properties that any
correct restatement of e3nn's conventions must satisfy."""
import math

import numpy as np
import pytest
import torch

from hamgnn_b200 import so3
from hamgnn_b200.irreps import Irreps as PIrreps
from oracle import e3lite as E
from oracle import hamgnn_ref as R


def rand_rot(seed):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(4, generator=g, dtype=torch.float64)
    a, b, c, d = (q / q.norm()).tolist()
    return torch.tensor([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                         [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                         [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]], dtype=torch.float64)


def wigner_D_from_sh(l, Rm):
    """D^l(R) in the reference's SH convention: Y(R v) = D Y(v), Y evaluated on v[:, [1,2,0]]."""
    g = torch.Generator().manual_seed(123)
    v = torch.randn(400, 3, generator=g, dtype=torch.float64)
    Y0 = E.spherical_harmonics([l], v[:, [1, 2, 0]])
    Y1 = E.spherical_harmonics([l], (v @ Rm.T)[:, [1, 2, 0]])
    return torch.linalg.lstsq(Y0, Y1).solution.T


def test_w3j_closed_forms():
    w = E.wigner_3j(1, 1, 1, dtype=torch.float64)
    assert abs(w[0, 1, 2] - 1 / math.sqrt(6)) < 1e-14 and abs(w[1, 0, 2] + 1 / math.sqrt(6)) < 1e-14
    assert torch.allclose(E.wigner_3j(1, 1, 0, dtype=torch.float64)[:, :, 0], torch.eye(3, dtype=torch.float64) / math.sqrt(3))
    assert torch.allclose(E.wigner_3j(1, 1, 2, dtype=torch.float64)[:, :, 2] * math.sqrt(30),
                          torch.diag(torch.tensor([-1.0, 2.0, -1.0], dtype=torch.float64)), atol=1e-13)
    for l in range(7):
        d = E.wigner_3j(l, 0, l, dtype=torch.float64)[:, 0, :]
        assert torch.allclose(d, torch.eye(2 * l + 1, dtype=torch.float64) / math.sqrt(2 * l + 1), atol=1e-13)


def test_w3j_norm_sparsity_and_product_tables_agree():
    nnz = tot = 0
    for l1 in range(7):
        for l2 in range(6):
            for l3 in range(abs(l1 - l2), min(l1 + l2, 6) + 1):
                a = E.wigner_3j(l1, l2, l3, dtype=torch.float64)
                assert abs(float(a.norm()) - 1) < 1e-12
                b = so3.wigner_3j(l1, l2, l3)
                assert np.abs(a.numpy() - b).max() < 1e-13      # product tables == oracle tables
                i, j, k, v = so3.cg_nnz(l1, l2, l3)
                dense = np.zeros_like(b)
                dense[i, j, k] = v
                assert np.abs(dense - b).max() < 1e-13 and (np.diff(k) >= 0).all()
                nnz += len(v)
                tot += b.size
    assert 0.10 < nnz / tot < 0.15                               # ~12.5 % non-zeros


def test_sh_normalisation_axes_and_recursion():
    g = torch.Generator().manual_seed(0)
    v = torch.randn(64, 3, generator=g, dtype=torch.float64)
    u = torch.nn.functional.normalize(v, dim=-1)
    Y = E.spherical_harmonics(list(range(8)), v)
    assert torch.allclose(Y[:, 1:4], math.sqrt(3) * u, atol=1e-13)            # Y_1 = sqrt3 (x,y,z) of the e3nn input
    off = 0
    for l in range(8):
        assert torch.allclose((Y[:, off:off + 2 * l + 1] ** 2).sum(-1), torch.full((64,), 2 * l + 1.0, dtype=torch.float64))
        off += 2 * l + 1
    sh = lambda l: E.spherical_harmonics([l], v)
    for l1 in range(4):
        for l2 in range(4):
            for l3 in range(abs(l1 - l2), l1 + l2 + 1):
                if (l1 + l2 + l3) % 2:
                    continue
                t = torch.einsum("ijk,zi,zj->zk", E.wigner_3j(l1, l2, l3, dtype=torch.float64), sh(l1), sh(l2))
                c = (t * sh(l3)).sum(-1) / (sh(l3) ** 2).sum(-1)
                assert (c > 0).all() and (t - c[:, None] * sh(l3)).abs().max() < 1e-12


def test_w3j_is_equivariant_under_wigner_D():
    Rm = rand_rot(7)
    for (l1, l2, l3) in [(1, 1, 1), (2, 1, 3), (2, 2, 2), (3, 5, 4), (6, 5, 1), (4, 3, 6)]:
        D1, D2, D3 = (wigner_D_from_sh(l, Rm) for l in (l1, l2, l3))
        w = E.wigner_3j(l1, l2, l3, dtype=torch.float64)
        w_rot = torch.einsum("ia,jb,kc,abc->ijk", D1, D2, D3, w)
        assert (w_rot - w).abs().max() < 1e-10


def test_normalize2mom_constants():
    assert abs(so3.normalize2mom_const("silu") - 1.6791767924) < 1e-6
    assert abs(so3.normalize2mom_const("ssp") - 1.8782046685) < 1e-6
    assert abs(so3.normalize2mom_const("tanh") - 1.5937334473) < 1e-6
    assert abs(E._second_moment_const("silu") - so3.normalize2mom_const("silu")) < 1e-12


def test_irreps_sort_simplify_semantics():
    a = E.Irreps("64x0e+64x0o+32x1o+16x1e")
    s, p, inv = a.sort()
    assert str(s) == "64x0o+64x0e+32x1o+16x1e" and p == (1, 0, 2, 3)
    assert str(E.Irreps("2x0e+3x0e+1x1o+1x0e").simplify()) == "5x0e+1x1o+1x0e"
    b = PIrreps("64x0e+64x0o+32x1o+16x1e")
    s2, p2, _ = b.sort()
    assert str(s2) == str(s) and tuple(p2) == p
    assert str(PIrreps("0e + 1o + 2e")) == "1x0e+1x1o+1x2e"


def test_full_model_equivariance_and_hermiticity():
    """Rotating the crystal rotates node features by D(R) and leaves scalar channels / Hermiticity intact."""
    from hamgnn_b200 import graph_data as gd
    torch.manual_seed(0)
    cfg = dict(irreps_node_features="6x0e+6x0o+4x1o+3x1e+2x2o+3x2e+2x3o+1x3e+1x4e", num_layers=2, num_radial=8,
               radial_MLP=[8, 8], irreps_edge_sh="0e+1o+2e+3o")
    pre = R.HamGNNConvE3(cfg).double()
    out = R.HamGNNPlusPlusOut(pre.irreps_node_features, pre.irreps_node_features, nao_max=14, add_H0=False).double()
    g = gd.Batch.from_data_list([gd.bulk_silicon(nao_max=14)])

    def run(Rm=None, inversion=False):
        d = R.AttrDict({k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in g.to_dict().items()})
        if Rm is not None:
            d["pos"], d["nbr_shift"] = d["pos"] @ Rm.T, d["nbr_shift"] @ Rm.T
        if inversion:
            d["pos"], d["nbr_shift"] = -d["pos"], -d["nbr_shift"]
        with torch.no_grad():
            rep = pre(d)
            return rep, out(d, rep)["hamiltonian"]

    Rm = rand_rot(3)
    (rep0, H0), (rep1, H1), (rep2, H2) = run(), run(Rm), run(inversion=True)
    off = 0
    for mul, ir in pre.irreps_node_features:
        D = wigner_D_from_sh(ir.l, Rm)
        a = rep0["node_attr"][:, off:off + mul * ir.dim].reshape(-1, mul, ir.dim)
        b = rep1["node_attr"][:, off:off + mul * ir.dim].reshape(-1, mul, ir.dim)
        assert (torch.einsum("ij,zuj->zui", D, a) - b).abs().max() < 1e-10 * max(1.0, float(a.abs().max()))
        c = rep2["node_attr"][:, off:off + mul * ir.dim].reshape(-1, mul, ir.dim)
        assert (c - ir.p * a).abs().max() < 1e-10 * max(1.0, float(a.abs().max()))   # parity
        off += mul * ir.dim
    n = g.num_nodes
    Hoff = H0[n:].view(-1, 14, 14)
    assert (Hoff - Hoff[g.inv_edge_idx].transpose(1, 2)).abs().max() == 0
    # s-s blocks (orbitals 0..2 are l=0) are rotation invariants
    assert (H0.view(-1, 14, 14)[:, :3, :3] - H1.view(-1, 14, 14)[:, :3, :3]).abs().max() < 1e-10
