"""Edge-sharded multi-GPU inference of one large crystal graph.

The reference only knows data-parallel replicas (PL DDP, /root/reference/hamgnn/main.py:296-323); a single
graph always lives on one device.  Here the directed edge list of ONE graph is split across the ranks of a
`torch.distributed` process group:

  * node tensors (pos, z, node features) are replicated, every per-edge tensor (edge features, SH, radial
    embedding, H0, predicted hopping blocks) exists only on the owning rank;
  * each undirected pair {e, inv(e)} is owned by one rank, so the inverse-edge symmetrisation of the hopping
    blocks (hamgnn_output.py:1231-1285) needs no communication;
  * the only exchange is ONE all-reduce (sum, fp32, N x D) of the partial receiver aggregates per
    ConvBlockE3 (convolution.py:147-149), issued on the compute stream right after the fused
    message+scatter kernel;
  * node-wise work (skip Linear, ResidualBlock, on-site head) is N-sized and runs redundantly on every rank.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .graph_data import Data


def shard_edges(g: Data, rank: int, world: int) -> Data:
    """Local view of graph `g` for `rank`: all nodes, the owned directed edges (global order preserved,
    closed under inversion), `inv_edge_idx` re-indexed locally, `edge_global_idx` for gathering results."""
    ei = g["edge_index"].cpu().numpy()
    inv = g["inv_edge_idx"].cpu().numpy()
    E = ei.shape[1]
    idx = np.arange(E)
    primary = idx <= inv                      # one representative per undirected pair (self-inverse edges included)
    prim_idx = idx[primary]
    bounds = np.linspace(0, len(prim_idx), world + 1).astype(np.int64)
    mine = prim_idx[bounds[rank]:bounds[rank + 1]]
    own = np.zeros(E, dtype=bool)
    own[mine] = True
    own[inv[mine]] = True
    sel = idx[own]
    pos_in_local = np.full(E, -1, dtype=np.int64)
    pos_in_local[sel] = np.arange(len(sel))
    out = {}
    for k, v in g.to_dict().items():
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == E and k not in ("edge_index",):
            out[k] = v[torch.from_numpy(sel)]
        else:
            out[k] = v
    out["edge_index"] = g["edge_index"][:, torch.from_numpy(sel)].contiguous()
    out["inv_edge_idx"] = torch.from_numpy(pos_in_local[inv[sel]])
    assert (out["inv_edge_idx"] >= 0).all()
    out["edge_global_idx"] = torch.from_numpy(sel)
    return Data(**out)


class AllReduceAggregates:
    """Callable installed on every ConvBlockE3 (`conv.reduce_fn`): sums the partial aggregates over ranks."""

    def __init__(self, group=None):
        self.group = group
        self.calls = 0
        self.bytes = 0

    def __call__(self, agg: torch.Tensor) -> torch.Tensor:
        import torch.distributed as dist
        dist.all_reduce(agg, op=dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
        self.bytes += agg.numel() * agg.element_size()
        return agg


def install_edge_sharding(pre_module, group=None) -> AllReduceAggregates:
    red = AllReduceAggregates(group)
    for conv in pre_module.convolutions:
        conv.reduce_fn = red
    return red
