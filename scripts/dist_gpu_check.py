#!/usr/bin/env python
"""Multi-GPU correctness check of the edge-sharded forward (run under torchrun, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_gpu_check.py
Every rank runs the sharded forward of the same graph; it is compared with the unsharded forward computed
on the same GPU (node features everywhere, hopping blocks on the edges the rank owns)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from hamgnn_b200 import graph_data as gd  # noqa: E402
from hamgnn_b200.dist import install_edge_sharding, shard_edges  # noqa: E402
from hgb_testlib import SMALL_CFG, build_pair  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = {} if "--default" in sys.argv else SMALL_CFG
    pre, out, _, _ = build_pair(cfg, nao_max=19, add_H0=True, seed=0)
    pre.to(dev)
    out.to(dev)
    g = gd.twisted_bilayer_graphene(m=3, seed=0)
    full = gd.Batch.from_data_list([g]).to(dev)
    with torch.no_grad():
        rep_f = pre(full)
        H_f = out(full, rep_f)["hamiltonian"]
    sh = shard_edges(gd.Batch.from_data_list([g]), rank, world)
    gidx = sh["edge_global_idx"].to(dev)
    local_b = gd.Batch(**sh.to_dict()).to(dev)
    red = install_edge_sharding(pre)
    with torch.no_grad():
        rep_s = pre(local_b)
        H_s = out(local_b, rep_s)["hamiltonian"]
    for conv in pre.convolutions:
        conv.reduce_fn = None
    N = g.num_nodes
    e_node = float((rep_s["node_attr"] - rep_f["node_attr"]).abs().max() / rep_f["node_attr"].abs().max())
    e_edge = float((rep_s["edge_attr"] - rep_f["edge_attr"][gidx]).abs().max() / rep_f["edge_attr"].abs().max())
    e_on = float((H_s[:N] - H_f[:N]).abs().max() / H_f.abs().max())
    e_off = float((H_s[N:] - H_f[N:][gidx]).abs().max() / H_f.abs().max())
    worst = torch.tensor([max(e_node, e_edge, e_on, e_off)], device=dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    print(f"rank {rank}/{world}: edges {len(gidx)}/{g.edge_index.shape[1]} all_reduce calls {red.calls} "
          f"rel err node {e_node:.2e} edge {e_edge:.2e} H_on {e_on:.2e} H_off {e_off:.2e}", flush=True)
    dist.barrier()
    if rank == 0:
        print("DIST_CHECK", "PASS" if float(worst) < 1e-5 else "FAIL", float(worst), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if float(worst) < 1e-5 else 1)


if __name__ == "__main__":
    main()
