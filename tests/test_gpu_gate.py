"""GPU parity tests of the radial gate pre-pass (a7: FullyConnectedNet [R, h1, h2, n_channels]) through hgb_radial_gate:
the tcgen05 kernel (radial_gate_tc_kernel, 3xTF32) and the fp32-FMA kernel against the fp64 oracle, tolerance 1e-5
relative; ragged edge counts around the 128-row tile; full forward with the tensor-core gate selected."""
import pytest
import torch

from hamgnn_b200 import graph_data as gd
from hamgnn_b200 import plan as P
from hgb_testlib import DEFAULT_CFG, SMALL_CFG, build_pair, oracle_forward, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("backend", ["simt", "tc"])
@pytest.mark.parametrize("cfg_name,n_edges", [("small", 1), ("small", 127), ("small", 300), ("default", 129), ("default", 1000)])
def test_radial_gate_matches_oracle(cfg_name, n_edges, backend):
    dev = torch.device("cuda:0")
    pre, out, opre, oout = build_pair(SMALL_CFG if cfg_name == "small" else DEFAULT_CFG)
    torch.manual_seed(n_edges)
    rbf = torch.randn(n_edges, pre.num_radial) * 0.5
    for blk, oblk in ((pre.convolutions[0].conv_tp, opre.convolutions[0].conv_tp),
                      (pre.pair_interactions[1].conv_tp, opre.pair_interactions[1].conv_tp)):
        with torch.no_grad():
            ref = [oblk.node_weight_generator.double()(rbf.double()), oblk.edge_weight_generator.double()(rbf.double())]
        blk.to(dev)
        direct = None if blk.op.direct_src is None else torch.zeros(blk.op.direct_blocks[1], device=dev)
        g = blk.op.radial_gate(blk.weights(direct), rbf.to(dev), backend=backend)
        torch.cuda.synchronize()
        for b in range(2):
            err = rel_err(g[b, :, :ref[b].shape[1]].cpu(), ref[b])
            assert err < TOL, (cfg_name, n_edges, backend, b, err)
            assert float(g[b, :, ref[b].shape[1]:].abs().max()) == 0 if g.shape[2] > ref[b].shape[1] else True   # padding untouched


@pytest.mark.parametrize("cfg_name,gname", [("small", "mixed"), ("default", "si")])
def test_full_forward_with_tensor_core_gate(cfg_name, gname):
    dev = torch.device("cuda:0")
    cfg = SMALL_CFG if cfg_name == "small" else DEFAULT_CFG
    pre, out, opre, oout = build_pair(cfg, nao_max=19, add_H0=False)
    graphs = [gd.bulk_silicon()] if gname == "si" else [gd.bulk_silicon(), gd.graphene(rep=(2, 2, 1), seed=1), gd.mos2_monolayer(seed=2)]
    batch = gd.Batch.from_data_list(graphs)
    d, rep, res = oracle_forward(opre, oout, batch)
    pre.to(dev)
    out.to(dev)
    old = (P.BACKEND, P.GATE_BACKEND)
    P.BACKEND, P.GATE_BACKEND = "tcg", "tc"
    try:
        b = gd.Batch(**batch.to_dict()).to(dev)
        with torch.no_grad():
            r = pre(b)
            o = out(b, r)
        torch.cuda.synchronize()
    finally:
        P.BACKEND, P.GATE_BACKEND = old
    e_node, e_edge = rel_err(r["node_attr"].cpu(), rep["node_attr"]), rel_err(r["edge_attr"].cpu(), rep["edge_attr"])
    e_h = rel_err(o["hamiltonian"].cpu(), res["hamiltonian"])
    print(f"[tcg + tc gate] rel err node {e_node:.2e} edge {e_edge:.2e} H {e_h:.2e}")
    # 'tcg' carries the truncating TMEM accumulation over all paths (see test_gpu_parity.py): bound 2e-5 for this backend
    assert e_node < 2 * TOL and e_edge < 2 * TOL and e_h < 2 * TOL
