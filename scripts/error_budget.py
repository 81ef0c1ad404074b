#!/usr/bin/env python
"""Error budget of the full forward (default model, bulk Si) against the fp64 oracle, per message backend, next to the
fp32-CPU-oracle drift.  Run on the GPU box:  python scripts/error_budget.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

from hamgnn_b200 import graph_data as gd  # noqa: E402
from hamgnn_b200 import plan as P  # noqa: E402
from hgb_testlib import DEFAULT_CFG, build_pair, oracle_forward, rel_err  # noqa: E402

torch.set_num_threads(16)
pre, out, opre, oout = build_pair(DEFAULT_CFG, nao_max=19, add_H0=False)
graphs = {"si": [gd.bulk_silicon()], "mixed": [gd.bulk_silicon(), gd.graphene(rep=(2, 2, 1), seed=1), gd.mos2_monolayer(seed=2)]}
dev = torch.device("cuda:0")
pre.to(dev)
out.to(dev)
for gname, gl in graphs.items():
    batch = gd.Batch.from_data_list(gl)
    d, rep, res = oracle_forward(opre, oout, batch)
    ref = {k: d[k].clone() for k in ("edge_attrs", "edge_embedding")}
    ref.update(node=rep["node_attr"].clone(), edge=rep["edge_attr"].clone(), H=res["hamiltonian"].clone())
    d32, rep32, res32 = oracle_forward(opre, oout, batch, dtype=torch.float32)
    print(f"[{gname}] fp32 CPU oracle drift: rbf {rel_err(d32['edge_embedding'], ref['edge_embedding']):.2e} node {rel_err(rep32['node_attr'], ref['node']):.2e} "
          f"edge {rel_err(rep32['edge_attr'], ref['edge']):.2e} H {rel_err(res32['hamiltonian'], ref['H']):.2e}")
    opre.double(); oout.double()
    for backend in ("simt", "tcg", "rot"):
        for gate in ("simt", "tc"):
            if backend == "simt" and gate == "tc":
                continue
            P.BACKEND, P.GATE_BACKEND = backend, gate
            b = gd.Batch(**batch.to_dict()).to(dev)
            with torch.no_grad():
                r = pre(b)
                o = out(b, r)
            torch.cuda.synchronize()
            print(f"[{gname}] {backend:4s} gate {gate:4s}: sh {rel_err(b['edge_attrs'].cpu(), ref['edge_attrs']):.2e} rbf {rel_err(b['edge_embedding'].cpu(), ref['edge_embedding']):.2e} "
                  f"node {rel_err(r['node_attr'].cpu(), ref['node']):.2e} edge {rel_err(r['edge_attr'].cpu(), ref['edge']):.2e} "
                  f"H {rel_err(o['hamiltonian'].cpu(), ref['H']):.2e}")
