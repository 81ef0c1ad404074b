"""GPU tests of the on-device graph construction (hgb_neighbor_list, hgb_edge_lookup) against the
host builder of the synthetic crystals (hamgnn_b200.graph_data.build_graph: scipy cKDTree, fp64, same neighbour rule
0 < d < rc_i + rc_j as /root/reference/hamgnn/models/base_model.py:87-178): identical edge lists (integers: bit-exact),
inverse-edge index, the DFT-edge matching of generate_graph (base_model.py:237-288) and `build_internal_graph=True`
through HamGNNConvE3 vs the oracle evaluated on the host-built graph of the same radius_scale (1e-5 relative)."""
import numpy as np
import pytest
import torch

from hamgnn_b200 import graph_build as gb
from hamgnn_b200 import graph_data as gd
from hgb_testlib import SMALL_CFG, build_pair, oracle_forward, rel_err

pytestmark = pytest.mark.gpu


def _crystals():
    return [("Si", gd.bulk_silicon(seed=0)), ("graphene3", gd.graphene(rep=(3, 3, 1), seed=1)), ("MoS2", gd.mos2_monolayer(seed=2)),
            ("mixed", gd.random_mixed_cell(n_atoms=12, species=(1, 6, 8, 14, 42, 16), seed=5)), ("tbg2", gd.twisted_bilayer_graphene(m=2, seed=0))]


def _host_graph(g, scale, pbc):
    """host builder on the fp32-rounded positions / cell that the device sees"""
    pos = g.pos.numpy().astype(np.float64)
    cell = g.cell.numpy().astype(np.float64).reshape(3, 3)
    return gd.build_graph(g.z.numpy(), pos, cell, pbc=pbc, radius_scale=scale, with_targets=False)


@pytest.mark.parametrize("scale", [1.0, 1.15])
def test_neighbor_list_matches_host_builder(scale):
    dev = torch.device("cuda:0")
    for name, g in _crystals():
        pbc = (True, True, True)      # the reference builds its internal graph with pbc=True in all directions (base_model.py:262)
        ref = _host_graph(g, scale, pbc)
        out = gb.neighbor_list(g.z.to(dev), g.pos.to(dev), g.cell.to(dev), radius_scale=scale, pbc=pbc)
        E = ref.edge_index.shape[1]
        assert out["edge_index"].shape[1] == E, (name, out["edge_index"].shape[1], E)
        assert torch.equal(out["edge_index"].cpu(), ref.edge_index), name
        assert torch.equal(out["cell_shift"].cpu(), ref.cell_shift), name
        assert rel_err(out["nbr_shift"].cpu(), ref.nbr_shift) < 1e-6
        assert torch.equal(out["offset"].cpu()[1:], torch.cumsum(torch.bincount(ref.edge_index[0], minlength=g.num_nodes), 0))
        inv = gb.edge_lookup(out["edge_index"], out["cell_shift"], out["edge_index"], out["cell_shift"], out["offset"], inverse=True)
        assert torch.equal(inv.cpu(), ref.inv_edge_idx), name
        print(f"{name} scale {scale}: N={g.num_nodes} E={E} identical")


def test_generate_graph_matches_dft_edges_of_a_batch():
    dev = torch.device("cuda:0")
    gs = [gd.bulk_silicon(seed=0), gd.mos2_monolayer(seed=2), gd.random_mixed_cell(n_atoms=9, species=(1, 6, 8, 14), seed=3)]
    batch = gd.Batch.from_data_list(gs).to(dev)
    g = gb.generate_graph(batch, radius_scale=1.2)
    m = g["matching_edges"]
    assert m.shape[0] == batch.edge_index.shape[1] and int(m.min()) >= 0
    assert torch.equal(g["edge_index"][:, m], batch.edge_index)            # cumulative node offsets (three crystals)
    assert torch.equal(g["cell_shift"][m], batch.cell_shift.to(torch.int64))
    assert g["edge_index"].shape[1] > batch.edge_index.shape[1]
    inv = g["inv_edge_idx_global"]
    assert torch.equal(inv[inv], torch.arange(inv.shape[0], device=dev))
    with pytest.raises(AssertionError):
        small = gd.Batch.from_data_list([gd.bulk_silicon(seed=0, radius_scale=1.3)]).to(dev)
        gb.generate_graph(small, radius_scale=1.05)                          # "Please increase radius_scale factor!"


def test_build_internal_graph_forward_matches_oracle():
    dev = torch.device("cuda:0")
    scale = 1.2
    cfg = dict(SMALL_CFG, build_internal_graph=True, radius_scale=scale)
    pre, out, opre, oout = build_pair(dict(SMALL_CFG), nao_max=19, add_H0=True)
    from hamgnn_b200.hamgnn_conv import HamGNNConvE3
    pre_int = HamGNNConvE3(cfg)
    pre_int.load_state_dict(pre.state_dict())
    g = gd.mos2_monolayer(seed=2)
    batch = gd.Batch.from_data_list([g])
    big = _host_graph(g, scale, (True, True, True))
    big_b = gd.Batch.from_data_list([big])
    from oracle import hamgnn_ref as R
    dd = R.AttrDict({k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in big_b.to_dict().items()})
    with torch.no_grad():
        rep_big = opre.double()(dd)
    # rows of the big graph that are the DFT edges of `g`
    key = lambda ei, cs: {(int(a), int(b), int(s[0]), int(s[1]), int(s[2])): k for k, (a, b, s) in enumerate(zip(ei[0].tolist(), ei[1].tolist(), cs.tolist()))}
    idx_big = key(big.edge_index, big.cell_shift)
    sel = torch.tensor([idx_big[(int(a), int(b), int(s[0]), int(s[1]), int(s[2]))]
                        for a, b, s in zip(g.edge_index[0].tolist(), g.edge_index[1].tolist(), g.cell_shift.tolist())])
    pre_int.to(dev)
    b = gd.Batch(**batch.to_dict()).to(dev)
    with torch.no_grad():
        rep = pre_int(b)
    torch.cuda.synchronize()
    e_n, e_e = rel_err(rep["node_attr"].cpu(), rep_big["node_attr"]), rel_err(rep["edge_attr"].cpu(), rep_big["edge_attr"][sel])
    print(f"build_internal_graph (radius_scale {scale}: {big.edge_index.shape[1]} internal edges for {g.edge_index.shape[1]} DFT edges): node {e_n:.2e} edge {e_e:.2e}")
    assert rep["edge_attr"].shape[0] == g.edge_index.shape[1]
    assert e_n < 1e-5 and e_e < 1e-5
