"""CPU tests of the spin-orbit (a16) path: the oracle's su2 recoupling is pinned by SU(2) x SO(3) equivariance and
Hermiticity (the reference ships no golden vectors), and the host planners of the product
(SortedHeadOp, SocSU2Assembly CSR, shell tables) are checked against the oracle through a CPU emulation of the
kernels' arithmetic."""
import math

import numpy as np
import pytest
import torch

import hgb_kernel_emulator as EM
from hamgnn_b200 import graph_data as gd
from hamgnn_b200.hamgnn_output import HamGNNPlusPlusOut
from hgb_testlib import rel_err
from oracle import hamgnn_ref as R
from test_oracle_so3 import wigner_D_from_sh

D_SMALL = "8x0e+8x0o+4x1o+4x1e+3x2o+5x2e+2x3o+2x3e+2x4e+1x4o+1x5o+1x5e"


def _axis_angle(axis, th):
    axis = torch.tensor(axis, dtype=torch.float64)
    axis = axis / axis.norm()
    K = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]], dtype=torch.float64)
    Rm = torch.eye(3, dtype=torch.float64) + math.sin(th) * K + (1 - math.cos(th)) * K @ K
    sx = torch.tensor([[0, 1], [1, 0]], dtype=torch.complex128)
    sy = torch.tensor([[0, -1j], [1j, 0]], dtype=torch.complex128)
    sz = torch.tensor([[1, 0], [0, -1]], dtype=torch.complex128)
    U = math.cos(th / 2) * torch.eye(2, dtype=torch.complex128) - 1j * math.sin(th / 2) * (axis[0] * sx + axis[1] * sy + axis[2] * sz)
    return Rm, U


@pytest.mark.parametrize("ls", [[0, 1, 2], [0, 0, 1, 2, 2]])
def test_su2_decomposition_is_equivariant_under_spin_and_orbital_rotation(ls):
    """H(D c) = (U x D_orb) H(c) (U x D_orb)^dagger with U = exp(-i theta n.sigma / 2): fixes the w3j(L,1,L')
    recoupling, the w3j(l1,l2,L) coupling and oyzx2spin conventions of get_H up to an overall constant."""
    torch.manual_seed(0)
    js = [(a, b) for a in ls for b in ls]
    nao = sum(2 * l + 1 for l in ls)
    dec = R.SU2Decomposition(js, nao)
    c = torch.randn(4, 2 * dec.base_irreps.dim, dtype=torch.float64)
    Rm, U = _axis_angle([0.3, -0.5, 0.8], 0.9)
    Ds = {l: wigner_D_from_sh(l, Rm) for l in range(dec.base_irreps.lmax + 1)}
    half = dec.base_irreps.dim
    cr = torch.zeros_like(c)
    for h in range(2):
        off = h * half
        for _, ir in dec.base_irreps:
            cr[:, off:off + ir.dim] = c[:, off:off + ir.dim] @ Ds[ir.l].T
            off += ir.dim
    mat = lambda x: dec.get_H(x).reshape(-1, 2, 2, nao, nao).transpose(2, 3).reshape(-1, 2 * nao, 2 * nao)
    H0, H1 = mat(c), mat(cr)
    T = torch.kron(U, torch.block_diag(*[Ds[l] for l in ls]).to(torch.complex128))
    assert (H1 - T @ H0 @ T.conj().T).abs().max() < 1e-12 * H0.abs().max()
    # a pure l=0 "scalar" coefficient only feeds the spin-diagonal blocks with equal weight
    c0 = torch.zeros(1, 2 * half, dtype=torch.float64)
    c0[0, 0] = 1.0
    Hs = mat(c0)[0]
    assert abs(Hs[0, 0] - Hs[nao, nao]) < 1e-15 and Hs[0, nao].abs() < 1e-15


@pytest.fixture(scope="module")
def soc_setup():
    torch.manual_seed(0)
    g = gd.Batch.from_data_list([gd.mos2_monolayer(seed=2, soc=True), gd.bulk_silicon(seed=1, soc=True)])
    x_n = torch.randn(g.num_nodes, R.Irreps(D_SMALL).dim, dtype=torch.float64)
    x_e = torch.randn(g.edge_index.shape[1], R.Irreps(D_SMALL).dim, dtype=torch.float64)
    return g, x_n, x_e


def _pair(basis, **kw):
    torch.manual_seed(1)
    out = HamGNNPlusPlusOut(D_SMALL, D_SMALL, nao_max=19, soc_switch=True, soc_basis=basis, ham_only=True, **kw)
    oout = R.HamGNNPlusPlusOut(D_SMALL, D_SMALL, nao_max=19, soc_switch=True, soc_basis=basis, **kw)
    missing, unexpected = oout.load_state_dict(out.state_dict(), strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return out, oout.double()


def _data64(g):
    return R.AttrDict({k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in g.to_dict().items()})


def test_su2_planner_matches_oracle(soc_setup):
    g, x_n, x_e = soc_setup
    out, oout = _pair("su2", add_H0=True)
    d = _data64(g)
    with torch.no_grad():
        ref = oout(d, {"node_attr": x_n, "edge_attr": x_e})
    asm = out.soc_assembly
    assert asm.base_dim == 1444 and asm.head_irreps.dim == 5776 and asm.base_irreps.lmax == 5
    on_row, off_row, inv = out._row_maps(g)
    s, r = g.edge_index
    N, E = g.num_nodes, s.shape[0]
    H_re = torch.zeros(N + E, (2 * 19) ** 2, dtype=torch.float64)
    H_im = torch.zeros_like(H_re)
    for net, x, partner, na, nb, rows, kre, kim in ((out.onsite_hamiltonian_network, x_n, None, None, None, on_row, "Hon0", "iHon0"),
                                                     (out.offsite_hamiltonian_network, x_e, inv, s, r, off_row, "Hoff0", "iHoff0")):
        y = EM.emulate_resblock(net.residual_block, x)
        coef = EM.emulate_sorted_head(net.head, net.linear_transform.weight, y)
        raw = EM.emulate_csr_rows(asm.row_ptr, asm.col, asm.val, asm.n_out, coef)
        re, im = EM.emulate_finalize_su2(19, asm.mask, raw, partner, d[kre], d[kim], g.z, na, nb)
        H_re[rows], H_im[rows] = re, im
    assert rel_err(H_re, ref["hamiltonian_real"]) < 2e-6 and rel_err(H_im, ref["hamiltonian_imag"]) < 2e-6
    assert ref["hamiltonian"].shape == (2 * (N + E), 1444)
    # Hermiticity of the predicted spin-orbital blocks (H0 is Hermitian by construction)
    Hc = torch.complex(H_re, H_im)[off_row].view(-1, 38, 38)
    assert (Hc - Hc[inv].conj().transpose(1, 2)).abs().max() < 1e-12


def test_so3_planner_matches_oracle(soc_setup):
    g, x_n, x_e = soc_setup
    for kw in (dict(add_H0=True), dict(add_H0=True, add_H_nonsoc=True), dict(add_H0=False, symmetrize=False)):
        out, oout = _pair("so3", **kw)
        d = _data64(g)
        with torch.no_grad():
            ref = oout(d, {"node_attr": x_n, "edge_attr": x_e})
        d = _data64(g)     # the oracle (like the reference) rewrites Hon0/Hoff0 under add_H_nonsoc
        on_row, off_row, inv = out._row_maps(g)
        s, r = g.edge_index
        N, E = g.num_nodes, s.shape[0]
        sym = kw.get("symmetrize", True)
        shells = [(out._shell_lo[i], out._shell_hi[i]) for i in range(out._n_shells)]
        assert shells == [(3, 6), (6, 9), (9, 14), (14, 19)]
        H_re = torch.zeros(N + E, 1444, dtype=torch.float64)
        H_im = torch.zeros_like(H_re)
        for hnet, knet, x, partner, na, nb, rows, kre, kim, lk, nsk in (
                (out.onsite_hamiltonian_network, out.onsite_ksi_network, x_n, None, None, None, on_row, "Hon0", "iHon0", "Lon", "Hon_nonsoc"),
                (out.offsite_hamiltonian_network, out.offsite_ksi_network, x_e, inv, s, r, off_row, "Hoff0", "iHoff0", "Loff", "Hoff_nonsoc")):
            if kw.get("add_H_nonsoc"):
                hns = d[nsk]
            else:
                coef = EM.emulate_resblock(hnet.residual_block, x, post=hnet.op, post_w=hnet.linear_transform.weight)
                hns = EM.emulate_ham(out.assembly, coef, partner, None, g.z, na, nb, symmetrize=sym)
            ksi = EM.emulate_resblock(knet.residual_block, x, post=knet.op, post_w=knet.linear_transform.weight)
            ksi = EM.emulate_ksi_shell_average(19, shells, ksi)
            re, im = EM.emulate_finalize_so3(19, hns, ksi, d[lk], partner, d[kre] if kw["add_H0"] else None,
                                             d[kim] if kw["add_H0"] else None, symmetrize=sym,
                                             h0_offdiag_only=bool(kw.get("add_H_nonsoc")))
            H_re[rows], H_im[rows] = re, im
        assert rel_err(H_re, ref["hamiltonian_real"]) < 2e-6 and rel_err(H_im, ref["hamiltonian_imag"]) < 2e-6, kw


def test_soc_ctor_contract():
    with pytest.raises(NotImplementedError):
        HamGNNPlusPlusOut(D_SMALL, D_SMALL, nao_max=19, soc_switch=True, soc_basis="u1", ham_only=True)
    with pytest.raises(NotImplementedError):
        HamGNNPlusPlusOut(D_SMALL, D_SMALL, nao_max=19, soc_switch=False, spin_constrained=True, ham_only=True)
    out = HamGNNPlusPlusOut(D_SMALL, D_SMALL, nao_max=14, soc_switch=True, soc_basis="su2", ham_only=False)
    assert out.hamiltonian_irreps_su2.dim == 2 * 784
    assert {"onsite_overlap_network.linear_transform.weight", "offsite_hamiltonian_network.linear_transform.weight"} <= set(out.state_dict())
