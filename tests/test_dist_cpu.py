"""Host-side logic of the edge-sharded multi-GPU path, on CPU with the gloo backend (world_size 2)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hamgnn_b200 import graph_data as gd
from hamgnn_b200.dist import AllReduceAggregates, shard_edges


def test_shards_partition_edges_and_are_closed_under_inversion():
    g = gd.graphene(rep=(3, 3, 1), seed=2)
    E = g.edge_index.shape[1]
    for world in (2, 3, 8):
        seen = torch.zeros(E, dtype=torch.long)
        sizes = []
        for r in range(world):
            s = shard_edges(g, r, world)
            gi = s["edge_global_idx"]
            seen[gi] += 1
            sizes.append(len(gi))
            inv = s["inv_edge_idx"]
            ei = s["edge_index"]
            assert (ei[0][inv] == ei[1]).all() and (ei[1][inv] == ei[0]).all()
            assert torch.equal(s["nbr_shift"][inv], -s["nbr_shift"])
            assert torch.equal(gi, gi.sort().values)                 # global order preserved
            assert torch.equal(g["Hoff0"][gi], s["Hoff0"])
            assert s["pos"].shape == g["pos"].shape                   # nodes replicated
        assert (seen == 1).all()
        assert max(sizes) - min(sizes) <= 2 + E // (50 * world)       # balanced by edge count


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hgb_testlib import SMALL_CFG, build_pair
        torch.set_num_threads(1)
        pre, out, opre, oout = build_pair(SMALL_CFG)
        g = gd.graphene(rep=(2, 2, 1), seed=1)
        D = pre.irreps_node_features.dim
        torch.manual_seed(1)
        x = torch.randn(g.num_nodes, D, dtype=torch.float64)
        efull = torch.randn(g.edge_index.shape[1], D, dtype=torch.float64)
        opre.double()
        conv = opre.convolutions[0]

        def partial(graph, e):
            from oracle import hamgnn_ref as R
            d = R.AttrDict({k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in graph.to_dict().items()})
            opre.edge_geometry(d)
            s, r = d["edge_index"]
            m = conv.conv_tp(x[s], x[r], e, d["edge_attrs"], d["edge_embedding"])
            return torch.zeros_like(x).index_add_(0, r, m)

        with torch.no_grad():
            full = partial(g, efull)
            sh = shard_edges(g, rank, world)
            mine = partial(sh, efull[sh["edge_global_idx"]])
            red = AllReduceAggregates()
            total = red(mine.clone())
        err = float((total - full).abs().max() / full.abs().max())
        ret[rank] = (err, red.calls, red.bytes)
    finally:
        dist.destroy_process_group()


def test_sharded_aggregate_allreduce_matches_single_process():
    world = 2
    port = 29000 + (os.getpid() % 2000)
    with mp.Manager() as man:
        ret = man.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        for r in range(world):
            err, calls, nbytes = ret[r]
            assert err < 1e-12, err
            assert calls == 1 and nbytes > 0
