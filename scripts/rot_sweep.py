"""In-process sweep of run-time switches of the message kernels: per-kernel device times (CUDA events around every launch of
the library, hgb_timing_*) of full forwards of one workload for each value of an environment variable the C side reads per call.

    python scripts/rot_sweep.py --workload tbg_m8 --var HGB_ROT_GD --values 111,211,311,411,221,321 [--backend rot]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench as B
from hamgnn_b200 import graph_data as gd, lib as L, plan as P

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="tbg_m8")
ap.add_argument("--var", default="HGB_ROT_GD")
ap.add_argument("--values", default="111,211,311,411,221,321")
ap.add_argument("--backend", default="rot")
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
graphs, desc, kw = B.build_workload(args.workload)
pre, out = B.build_models(kw)
dev = torch.device("cuda:0")
pre.to(dev); out.to(dev)
batch = gd.Batch.from_data_list(graphs).to(dev)
P.BACKEND = args.backend
E = batch.edge_index.shape[1]
ref = None
for v in args.values.split(","):
    os.environ[args.var] = v
    with torch.no_grad():
        for _ in range(2):
            o = out(batch, pre(batch))
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            o = out(batch, pre(batch))
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / args.steps
        L.timing_enable(True)
        for _ in range(args.steps):
            o = out(batch, pre(batch))
        t = L.timing_collect()
        L.timing_enable(False)
    h = o["hamiltonian"].float()
    if ref is None:
        ref = h.clone()
    dmax = float((h - ref).abs().max() / ref.abs().max())
    print(json.dumps({"var": args.var, "value": v, "workload": args.workload, "E": E, "ms_per_forward": round(ms, 3),
                      "per_kernel_ms": {k: round(x[0] / args.steps, 3) for k, x in t.items()}, "H_vs_first": dmax}), flush=True)
