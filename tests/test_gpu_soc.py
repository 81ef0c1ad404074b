"""GPU parity tests of the spin-orbit heads and the overlap head, through the C ABI:
hgb_linear_forward_ld, hgb_csr_rows, hgb_ham_finalize_su2, hgb_ksi_shell_average, hgb_ham_finalize_so3.

Tolerance: 1e-5 relative (max|a-b| / max|b| per tensor) against the fp64 oracle on the same fp32 weights; Hermiticity
with the inverse edge and the zeros of masked orbitals are checked bit-exactly."""
import pytest
import torch

import hgb_kernel_emulator as EM
from hamgnn_b200 import graph_data as gd
from hamgnn_b200.hamgnn_conv import HamGNNConvE3
from hamgnn_b200.hamgnn_output import HamGNNPlusPlusOut
from hamgnn_b200.irreps import Irreps
from hamgnn_b200.plan import SortedHeadOp
from hgb_testlib import rel_err
from oracle import hamgnn_ref as R

pytestmark = pytest.mark.gpu
TOL = 1e-5
CFG = dict(irreps_node_features="8x0e+8x0o+4x1o+4x1e+3x2o+5x2e+2x3o+2x3e+2x4e+1x4o+1x5o+1x5e", num_layers=2, num_radial=16,
           radial_MLP=[16, 16], irreps_edge_sh="0e+1o+2e+3o+4e", cutoff=26.0)


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _pair(cfg, nao_max=19, seed=0, **kw):
    torch.manual_seed(seed)
    pre = HamGNNConvE3(cfg)
    D = str(pre.irreps_node_features)
    out = HamGNNPlusPlusOut(D, D, nao_max=nao_max, **kw)
    opre = R.HamGNNConvE3(cfg)
    okw = {k: v for k, v in kw.items()}
    oout = R.HamGNNPlusPlusOut(D, D, nao_max=nao_max, **okw)
    assert not opre.load_state_dict(pre.state_dict(), strict=False).missing_keys
    res = oout.load_state_dict(out.state_dict(), strict=False)
    assert not res.missing_keys and not res.unexpected_keys, res
    return pre, out, opre.double(), oout.double()


def _oracle(opre, oout, batch):
    d = R.AttrDict({k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.to_dict().items()})
    with torch.no_grad():
        rep = opre(d)
        return d, rep, oout(d, rep)


def _run(pre, out, batch, dev):
    pre.to(dev)
    out.to(dev)
    b = gd.Batch(**batch.to_dict()).to(dev)
    with torch.no_grad():
        o = out(b, pre(b))
    torch.cuda.synchronize()
    return b, o


def _batch(nao_max=19):
    return gd.Batch.from_data_list([gd.mos2_monolayer(seed=2, soc=True, nao_max=nao_max),
                                    gd.bulk_silicon(seed=1, soc=True, nao_max=nao_max),
                                    gd.graphene(rep=(2, 2, 1), seed=3, soc=True, nao_max=nao_max)])


@pytest.mark.parametrize("add_H0", [True, False])
def test_su2_forward_matches_oracle(add_H0):
    dev = _dev()
    pre, out, opre, oout = _pair(CFG, soc_switch=True, soc_basis="su2", ham_only=True, add_H0=add_H0)
    batch = _batch()
    d, rep, ref = _oracle(opre, oout, batch)
    b, o = _run(pre, out, batch, dev)
    e_re = rel_err(o["hamiltonian_real"].cpu(), ref["hamiltonian_real"])
    e_im = rel_err(o["hamiltonian_imag"].cpu(), ref["hamiltonian_imag"])
    print(f"su2 (add_H0={add_H0}) rel err real {e_re:.2e} imag {e_im:.2e}")
    assert e_re < TOL and e_im < TOL
    assert o["hamiltonian"].shape == ref["hamiltonian"].shape and rel_err(o["hamiltonian"].cpu(), ref["hamiltonian"]) < TOL
    assert rel_err(b["hamiltonian"].cpu(), d["hamiltonian"]) == 0          # targets stacked real;imag like the reference
    # Hermiticity with the inverse edge, bit for bit; masked orbitals exactly zero (prediction only)
    if not add_H0:
        on_row, off_row, inv = out._row_maps(b)
        M = 2 * out.nao_max
        Hc = torch.complex(o["hamiltonian_real"], o["hamiltonian_imag"])
        Hoff, Hon = Hc[off_row].view(-1, M, M), Hc[on_row].view(-1, M, M)
        assert torch.equal(Hoff, Hoff[inv].conj().transpose(1, 2)) and torch.equal(Hon, Hon.conj().transpose(1, 2))
        mask = torch.from_numpy(out.soc_assembly.mask).to(dev).bool()[b["z"]].repeat(1, 2)
        assert (Hon[~(mask[:, :, None] & mask[:, None, :])] == 0).all()


@pytest.mark.parametrize("kw", [dict(add_H0=True), dict(add_H0=True, add_H_nonsoc=True), dict(add_H0=False, symmetrize=False)])
def test_so3_forward_matches_oracle(kw):
    dev = _dev()
    pre, out, opre, oout = _pair(CFG, soc_switch=True, soc_basis="so3", ham_only=True, **kw)
    batch = _batch()
    d, rep, ref = _oracle(opre, oout, batch)
    b, o = _run(pre, out, batch, dev)
    e_re = rel_err(o["hamiltonian_real"].cpu(), ref["hamiltonian_real"])
    e_im = rel_err(o["hamiltonian_imag"].cpu(), ref["hamiltonian_imag"])
    print(f"so3 {kw} rel err real {e_re:.2e} imag {e_im:.2e}")
    assert e_re < TOL and e_im < TOL


def test_overlap_head_and_zero_point_shift():
    dev = _dev()
    pre, out, opre, oout = _pair(CFG, soc_switch=False, ham_only=False, add_H0=True, zero_point_shift=True)
    batch = gd.Batch.from_data_list([gd.bulk_silicon(seed=1), gd.graphene(rep=(2, 2, 1), seed=3)])
    d, rep, ref = _oracle(opre, oout, batch)
    b, o = _run(pre, out, batch, dev)
    assert rel_err(o["overlap"].cpu(), ref["overlap"]) < TOL
    assert rel_err(o["hamiltonian"].cpu(), ref["hamiltonian"]) < TOL


def test_su2_zero_point_shift_and_nao26_chunked_head():
    dev = _dev()
    pre, out, opre, oout = _pair(CFG, nao_max=26, soc_switch=True, soc_basis="su2", ham_only=True, add_H0=True,
                                 zero_point_shift=True)
    assert len(out.offsite_hamiltonian_network.head.chunks) >= 2           # 5274 sorted columns: evaluated in column chunks
    batch = gd.Batch.from_data_list([gd.bulk_silicon(seed=1, soc=True, nao_max=26)])
    d, rep, ref = _oracle(opre, oout, batch)
    b, o = _run(pre, out, batch, dev)
    assert rel_err(o["hamiltonian_real"].cpu(), ref["hamiltonian_real"]) < TOL
    assert rel_err(o["hamiltonian_imag"].cpu(), ref["hamiltonian_imag"]) < TOL


def test_sorted_head_chunks_and_strided_linear():
    """hgb_linear_forward_ld writes column windows of a wider row; chunking must not change the result."""
    dev = _dev()
    torch.manual_seed(3)
    D = Irreps(CFG["irreps_node_features"])
    outs = Irreps("+".join(["0e", "1o", "2e", "1e", "0o", "3o", "2e", "5o", "4e", "6e"] * 7))
    used = list(range(0, len(outs), 2))
    w = torch.randn(SortedHeadOp(D, outs, used).weight_numel)
    x = torch.randn(37, D.dim)
    ref = EM.emulate_sorted_head(SortedHeadOp(D, outs, used), w.double(), x.double())
    for max_cols in (10 ** 6, 64, 24):
        head = SortedHeadOp(D, outs, used, max_cols=max_cols)
        y = head.forward(w.to(dev), x.to(dev))
        torch.cuda.synchronize()
        assert rel_err(y.cpu(), ref) < 2e-6, max_cols
    assert (SortedHeadOp(D, outs, used).pos[[outs[o].ir.l == 6 for o in used]] == -1).all()   # no 6e input -> zero output
