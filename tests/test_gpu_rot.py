"""GPU parity tests of the edge-aligned ('rot') message path, stage by stage through the C ABI:
hgb_wigner vs the fp64 emulation, single fused-message calls (scatter form, edge-update form with the direct
skip path, embedding form) vs the oracle, edge chunking, and the full forward vs the oracle.  Tolerance 1e-5
relative (north-star bar), the oracle evaluated in fp64 on the same fp32 weights."""
import numpy as np
import pytest
import torch

from hamgnn_b200 import graph_data as gd
from hamgnn_b200 import plan as P
import hgb_kernel_emulator as EM
from hgb_testlib import DEFAULT_CFG, SMALL_CFG, build_pair, oracle_forward, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _graphs(which):
    if which == "si":
        return [gd.bulk_silicon()]
    return [gd.bulk_silicon(), gd.graphene(rep=(2, 2, 1), seed=1), gd.mos2_monolayer(seed=2)]


@pytest.fixture(scope="module", params=[("small", "mixed"), ("default", "si")])
def setup(request):
    cfg_name, gname = request.param
    cfg = SMALL_CFG if cfg_name == "small" else DEFAULT_CFG
    pre, out, opre, oout = build_pair(cfg, nao_max=19, add_H0=False)
    batch = gd.Batch.from_data_list(_graphs(gname))
    d, rep, res = oracle_forward(opre, oout, batch)
    dev = torch.device("cuda:0")
    pre.to(dev)
    out.to(dev)
    return cfg_name, pre, out, opre, oout, batch, d, rep, res, dev


def test_wigner_kernel_matches_emulation(setup):
    cfg_name, pre, out, opre, oout, batch, d, rep, res, dev = setup
    op = pre.convolutions[0].conv_tp.op
    vec = d["edge_vectors"].float()
    # include the two poles and a vector a hair off the pole
    extra = torch.tensor([[0, 0, 1.0], [0, 0, -1.0], [1e-7, 0, 1.0], [0.6, 0.0, 0.8], [0, -1.0, 0]])
    vec = torch.cat([vec, extra / extra.norm(dim=1, keepdim=True)]).contiguous()
    dw = P.wigner_for(op, vec.to(dev)).cpu().double().numpy()
    ref = EM.emulate_wigner(vec.double().numpy(), op)
    for l in range(op.rot_lmax + 1):
        dl = 2 * l + 1
        a = dw[:, op.rot_doff[l]:op.rot_doff[l] + dl * dl]
        b = ref[:, op.rot_doff[l]:op.rot_doff[l] + dl * dl]
        assert np.abs(a - b).max() < 2e-7, (l, np.abs(a - b).max())


@pytest.mark.parametrize("chunk", [None, 128])
def test_single_message_calls(setup, chunk):
    """Each fused-message form on random inputs: rot kernel vs the oracle module, and vs the tcg kernel."""
    cfg_name, pre, out, opre, oout, batch, d, rep, res, dev = setup
    if chunk is not None and cfg_name == "default":
        pytest.skip("chunking is exercised on the small model")
    torch.manual_seed(11)
    E, N, D = batch.edge_index.shape[1], batch.num_nodes, pre.irreps_node_features.dim
    x, e = torch.randn(N, D), torch.randn(E, D)
    s, r = batch.edge_index
    dd = {"edge_index": batch.edge_index, "node_features": x.double(), "edge_features": e.double(),
          "edge_attrs": d["edge_attrs"], "edge_embedding": d["edge_embedding"]}
    with torch.no_grad():
        ref_pair = opre.pair_interactions[1](dict(dd))
        ref_msg = opre.convolutions[0].conv_tp(x.double()[s], x.double()[r], e.double(), d["edge_attrs"], d["edge_embedding"])
        ref_agg = torch.zeros(N, D, dtype=torch.float64).index_add_(0, r, ref_msg)
    sh, rbf, vec = d["edge_attrs"].float().to(dev), d["edge_embedding"].float().to(dev), d["edge_vectors"].float().to(dev)
    xd, ed, sd, rd = x.to(dev), e.to(dev), s.to(dev), r.to(dev)
    old = (P.BACKEND, P.ROT_CHUNK_EDGES)
    got = {}
    try:
        if chunk is not None:
            P.ROT_CHUNK_EDGES = chunk
        for backend in ("rot", "tcg"):
            P.BACKEND = backend
            cb = pre.convolutions[0].conv_tp
            msg = torch.empty(E, D, device=dev)
            cb.op.forward(cb.weights(), [xd, xd, ed], [sd, rd, None], sh, rbf, E, msg, edge_vec=vec)
            agg = torch.zeros(N, D, device=dev)
            cb.op.forward(cb.weights(), [xd, xd, ed], [sd, rd, None], sh, rbf, E, agg, out_index=rd, edge_vec=vec)
            pb = pre.pair_interactions[1]
            b = gd.Batch(**batch.to_dict()).to(dev)
            b["node_features"], b["edge_features"], b["edge_attrs"], b["edge_embedding"], b["edge_vectors"] = xd, ed, sh, rbf, vec
            pair = pb(b)
            torch.cuda.synchronize()
            got[backend] = (msg.cpu(), agg.cpu(), pair.cpu())
    finally:
        P.BACKEND, P.ROT_CHUNK_EDGES = old
    # the receiver reduction of the rot path is a serial segment sum (hgb_segment_sum): bit-reproducible
    P.BACKEND = "rot"
    try:
        if chunk is not None:
            P.ROT_CHUNK_EDGES = chunk
        cb = pre.convolutions[0].conv_tp
        agg2 = torch.full((N, D), float("nan"), device=dev)
        cb.op.forward(cb.weights(), [xd, xd, ed], [sd, rd, None], sh, rbf, E, agg2, out_index=rd, edge_vec=vec)
        torch.cuda.synchronize()
    finally:
        P.BACKEND, P.ROT_CHUNK_EDGES = old
    assert torch.equal(agg2.cpu(), got["rot"][1]), "rot aggregate must be bit-reproducible run to run"
    for backend in ("tcg", "rot"):
        em, ea, ep = (rel_err(got[backend][0], ref_msg), rel_err(got[backend][1], ref_agg), rel_err(got[backend][2], ref_pair))
        print(f"[{cfg_name} {backend} chunk={chunk}] rel err message {em:.2e} scatter {ea:.2e} edge update {ep:.2e}")
    em, ea, ep = (rel_err(got["rot"][0], ref_msg), rel_err(got["rot"][1], ref_agg), rel_err(got["rot"][2], ref_pair))
    assert em < TOL and ea < TOL and ep < TOL


@pytest.mark.parametrize("gate", ["simt", "tc"])
@pytest.mark.parametrize("cfg_name,gname", [("small", "mixed"), ("default", "si")])
def test_full_forward_rot_backend(cfg_name, gname, gate):
    cfg = SMALL_CFG if cfg_name == "small" else DEFAULT_CFG
    pre, out, opre, oout = build_pair(cfg, nao_max=19, add_H0=False)
    batch = gd.Batch.from_data_list(_graphs(gname))
    d, rep, res = oracle_forward(opre, oout, batch)
    dev = torch.device("cuda:0")
    pre.to(dev)
    out.to(dev)
    errs = {}
    old = (P.BACKEND, P.GATE_BACKEND)
    try:
        for backend in ("tcg", "rot"):
            P.BACKEND, P.GATE_BACKEND = backend, gate
            b = gd.Batch(**batch.to_dict()).to(dev)
            with torch.no_grad():
                r = pre(b)
                o = out(b, r)
            torch.cuda.synchronize()
            errs[backend] = (rel_err(r["node_attr"].cpu(), rep["node_attr"]), rel_err(r["edge_attr"].cpu(), rep["edge_attr"]),
                             rel_err(o["hamiltonian"].cpu(), res["hamiltonian"]))
            print(f"[{cfg_name} {gname} {backend} gate={gate}] rel err node {errs[backend][0]:.2e} edge {errs[backend][1]:.2e} H {errs[backend][2]:.2e}")
    finally:
        P.BACKEND, P.GATE_BACKEND = old
    assert max(errs["rot"]) < TOL, errs
