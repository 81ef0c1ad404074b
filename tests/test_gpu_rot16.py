"""GPU parity tests of the fp16 x 2 split variant of the edge-aligned message path (`HGB_MSGPACK=rot16`,
csrc/msgpack_rot16_kernel.cuh) through the C ABI (hgb_msgpack_rot16_forward): the three fused-message forms vs the oracle
module (hamgnn/nn/message_passing.py:26-231 restated in oracle/), edge chunking, bit-reproducibility of the receiver
reduction, and the full forward vs the oracle.  Tolerance 1e-5 relative (north-star bar), oracle in fp64 on the same fp32
weights."""
import os

import pytest
import torch

from hamgnn_b200 import graph_data as gd
from hamgnn_b200 import plan as P
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


def _single_calls(pre, batch, d, dev, x, e, backend, flags=0, chunk=None):
    E, N, D = batch.edge_index.shape[1], batch.num_nodes, pre.irreps_node_features.dim
    s, r = batch.edge_index
    sh, rbf, vec = d["edge_attrs"].float().to(dev), d["edge_embedding"].float().to(dev), d["edge_vectors"].float().to(dev)
    xd, ed, sd, rd = x.to(dev), e.to(dev), s.to(dev), r.to(dev)
    old = (P.BACKEND, P.ROT_CHUNK_EDGES, P.ROT16_FLAGS)
    try:
        P.BACKEND, P.ROT16_FLAGS = backend, flags
        if chunk is not None:
            P.ROT_CHUNK_EDGES = chunk
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
    finally:
        P.BACKEND, P.ROT_CHUNK_EDGES, P.ROT16_FLAGS = old
    return msg.cpu(), agg.cpu(), pair.cpu()


@pytest.mark.parametrize("chunk", [None, 128])
def test_single_message_calls_rot16(setup, chunk):
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
    got = _single_calls(pre, batch, d, dev, x, e, "rot16", 0, chunk)
    errs = (rel_err(got[0], ref_msg), rel_err(got[1], ref_agg), rel_err(got[2], ref_pair))
    print(f"[{cfg_name} rot16 chunk={chunk}] rel err message {errs[0]:.2e} scatter {errs[1]:.2e} edge update {errs[2]:.2e}")
    if max(errs) >= TOL:
        # diagnostics for a failing run: the other half-word order of the packed TMEM operand, and the tf32 kernel
        alt = _single_calls(pre, batch, d, dev, x, e, "rot16", 1, chunk)
        ref32 = _single_calls(pre, batch, d, dev, x, e, "rot", 0, chunk)
        print("swapped halves:", rel_err(alt[0], ref_msg), rel_err(alt[1], ref_agg), rel_err(alt[2], ref_pair))
        print("rot (tf32)    :", rel_err(ref32[0], ref_msg), rel_err(ref32[1], ref_agg), rel_err(ref32[2], ref_pair))
        D_ = pre.irreps_node_features
        off = 0
        for m in D_:
            w = m.mul * m.ir.dim
            print(f"  slot {m}: msg err {rel_err(got[0][:, off:off + w], ref_msg[:, off:off + w]):.2e}"
                  f" swapped {rel_err(alt[0][:, off:off + w], ref_msg[:, off:off + w]):.2e}")
            off += w
    assert max(errs) < TOL, errs
    again = _single_calls(pre, batch, d, dev, x, e, "rot16", 0, chunk)
    assert torch.equal(again[1], got[1]), "rot16 aggregate must be bit-reproducible run to run"


@pytest.mark.parametrize("cfg_name,gname", [("small", "mixed"), ("default", "si")])
def test_full_forward_rot16_backend(cfg_name, gname):
    cfg = SMALL_CFG if cfg_name == "small" else DEFAULT_CFG
    pre, out, opre, oout = build_pair(cfg, nao_max=19, add_H0=False)
    batch = gd.Batch.from_data_list(_graphs(gname))
    d, rep, res = oracle_forward(opre, oout, batch)
    dev = torch.device("cuda:0")
    pre.to(dev)
    out.to(dev)
    old = (P.BACKEND, P.GATE_BACKEND)
    try:
        P.BACKEND, P.GATE_BACKEND = "rot16", "tc"
        b = gd.Batch(**batch.to_dict()).to(dev)
        with torch.no_grad():
            r = pre(b)
            o = out(b, r)
        torch.cuda.synchronize()
    finally:
        P.BACKEND, P.GATE_BACKEND = old
    errs = (rel_err(r["node_attr"].cpu(), rep["node_attr"]), rel_err(r["edge_attr"].cpu(), rep["edge_attr"]),
            rel_err(o["hamiltonian"].cpu(), res["hamiltonian"]))
    print(f"[{cfg_name} {gname} rot16] rel err node {errs[0]:.2e} edge {errs[1]:.2e} H {errs[2]:.2e}")
    assert max(errs) < TOL, errs


@pytest.mark.parametrize("fma", ["0", "1", "5", "13", "23"])
def test_rot_path_variants(setup, fma):
    """Every variant of the default backend that ships in the library (HGB_ROT_FMA bit mask, read per call by
    hgb_msgpack_rot_forward): 0 = msgpack_rot_kernel for every slot class, 1 = class 16 on msgpack_rotf_kernel with one gate
    warpgroup, 5 = two gate warpgroups, 13 = + two GEMM1 issuer warps, 23 = class 32 on the FMA pipes as well (both warpgroups on
    every step); the default (21 = two ring stages) is what every other test runs.  Same bar as the default: 1e-5."""
    cfg_name, pre, out, opre, oout, batch, d, rep, res, dev = setup
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
    old = os.environ.get("HGB_ROT_FMA")
    try:
        os.environ["HGB_ROT_FMA"] = fma
        got = _single_calls(pre, batch, d, dev, x, e, "rot", 0, None)
    finally:
        if old is None:
            os.environ.pop("HGB_ROT_FMA", None)
        else:
            os.environ["HGB_ROT_FMA"] = old
    errs = (rel_err(got[0], ref_msg), rel_err(got[1], ref_agg), rel_err(got[2], ref_pair))
    print(f"[{cfg_name} rot HGB_ROT_FMA={fma}] rel err message {errs[0]:.2e} scatter {errs[1]:.2e} edge update {errs[2]:.2e}")
    assert max(errs) < TOL, errs
