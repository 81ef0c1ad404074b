#!/usr/bin/env python
"""Reference pin, one command away (VERDICT r1 item 1d / SURVEY.md section 8c item 5).

Run this in an environment that HAS the reference's dependencies (python 3.9, torch 2.5, e3nn==0.5.0, torch_scatter,
torch_geometric, easydict, ase, pymatgen, opt_einsum(_fx), torch_runstats -- HamGNN.yaml):

    python scripts/dump_reference_golden.py --reference /path/to/HamGNN [--out tests/golden]

It imports the REAL reference modules (`hamgnn.models.hamgnn_conv.HamGNNConvE3`, `hamgnn.models.hamgnn_output.
HamGNNPlusPlusOut`), loads the seeded weights of this repository's modules into them (same parameter names and e3nn
flat layouts; the e3nn constant buffers keep their own values), runs them on the seeded synthetic crystals of
SURVEY.md section 8d and writes tests/golden/ref_<case>.npz with inputs, weights and per-stage outputs.  When those
files exist, tests/test_golden.py checks the oracle (CPU, fp64) and the CUDA path (GPU) against them at 1e-5 -- that
turns "parity unpinned" into a pin against e3nn itself.  None of this can run in the build image (no e3nn, no network).
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    # name: (model config overrides, nao_max, output kwargs, graph builders)
    "si_default": ({}, 19, dict(soc_switch=False, ham_only=True, add_H0=True), ["bulk_silicon"]),
    "mixed_small": (dict(irreps_node_features="8x0e+8x0o+4x1o+4x1e+3x2o+5x2e+2x3o+2x3e+2x4e", num_layers=2, num_radial=16,
                         radial_MLP=[16, 16], irreps_edge_sh="0e+1o+2e+3o+4e"), 19,
                    dict(soc_switch=False, ham_only=True, add_H0=True), ["bulk_silicon", "graphene2", "mos2"]),
    "mos2_su2": ({}, 19, dict(soc_switch=True, soc_basis="su2", ham_only=True, add_H0=True), ["mos2_soc"]),
    "uni_nao26": (dict(legacy_edge_update=True, use_corr_prod=False), 26, dict(soc_switch=False, ham_only=True, add_H0=True), ["mixed26"]),
}


def graphs_for(names, nao_max):
    from hamgnn_b200 import graph_data as gd
    mk = {"bulk_silicon": lambda: gd.bulk_silicon(seed=0, nao_max=nao_max),
          "graphene2": lambda: gd.graphene(rep=(2, 2, 1), seed=1, nao_max=nao_max),
          "mos2": lambda: gd.mos2_monolayer(seed=2, nao_max=nao_max),
          "mos2_soc": lambda: gd.mos2_monolayer(seed=2, soc=True, nao_max=nao_max),
          "mixed26": lambda: gd.random_mixed_cell(n_atoms=10, species=(1, 6, 8, 14, 42, 16), seed=11, nao_max=nao_max)}
    return [mk[n]() for n in names]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of QuantumLab-ZY/HamGNN (the directory that contains hamgnn/)")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()
    # our own (dependency-free) modules first: they provide the seeded weights and the synthetic graphs
    from hamgnn_b200.hamgnn_conv import DEFAULTS, HamGNNConvE3 as OurPre
    from hamgnn_b200.hamgnn_output import HamGNNPlusPlusOut as OurOut
    ours = {}
    for name, (cfg, nao, okw, gnames) in CASES.items():
        torch.manual_seed(0)
        pre = OurPre(dict(cfg))
        out = OurOut(pre.irreps_node_features, pre.irreps_node_features, nao_max=nao, **okw)
        ours[name] = (pre.state_dict(), out.state_dict(), str(pre.irreps_node_features))
    # now the real reference.  Our repository also ships a `hamgnn` import shim: make sure the reference's package wins.
    for m in [k for k in sys.modules if k == "hamgnn" or k.startswith("hamgnn.")]:
        del sys.modules[m]
    sys.path.insert(0, os.path.abspath(args.reference))
    from easydict import EasyDict
    from torch_geometric.data import Batch, Data
    from hamgnn.models.hamgnn_conv import HamGNNConvE3
    from hamgnn.models.hamgnn_output import HamGNNPlusPlusOut
    assert "hamgnn_b200" not in HamGNNConvE3.__module__
    os.makedirs(args.out, exist_ok=True)
    for name, (cfg, nao, okw, gnames) in CASES.items():
        c = dict(DEFAULTS)
        c.update(cfg)
        c.update(radius_type="openmx", cutoff_func="cos", set_features=True)
        rep_cfg = EasyDict({"HamGNN_pre": EasyDict(c)})
        ref_pre = HamGNNConvE3(rep_cfg)
        ref_out = HamGNNPlusPlusOut(irreps_in_node=ref_pre.irreps_node_features, irreps_in_edge=ref_pre.irreps_node_features,
                                    nao_max=nao, ham_type="openmx", symmetrize=True, calculate_band_energy=False,
                                    nonlinearity_type="gate", zero_point_shift=False, **okw)
        sd_pre, sd_out, irreps = ours[name]
        for mod, sd in ((ref_pre, sd_pre), (ref_out, sd_out)):
            own = mod.state_dict()
            missing = [k for k in sd if k not in own]
            bad = [k for k in sd if k in own and tuple(own[k].shape) != tuple(sd[k].shape)]
            assert not missing and not bad, f"{name}: parameter names / shapes differ from the reference: {missing[:5]} {bad[:5]}"
            res = mod.load_state_dict(sd, strict=False)
            assert not res.unexpected_keys
        gs = graphs_for(gnames, nao)
        batch = Batch.from_data_list([Data(**{k: v for k, v in g.to_dict().items()}) for g in gs])
        with torch.no_grad():
            rep = ref_pre(batch)
            res = ref_out(batch, rep)
        payload = {}
        for gi, g in enumerate(gs):
            payload.update({f"in{gi}_{k}": v.numpy() for k, v in g.to_dict().items() if torch.is_tensor(v)})
        payload["n_graphs"] = np.array(len(gs))
        payload.update({f"pre_{k}": v.numpy() for k, v in sd_pre.items()})
        payload.update({f"out_{k}": v.numpy() for k, v in sd_out.items()})
        for k in ("edge_attrs", "edge_embedding", "edge_vectors", "edge_lengths"):
            payload[f"ref_{k}"] = batch[k].numpy()
        payload["ref_node_attr"] = rep["node_attr"].numpy()
        payload["ref_edge_attr"] = rep["edge_attr"].numpy()
        for k in ("hamiltonian", "hamiltonian_real", "hamiltonian_imag"):
            if k in res and res[k] is not None:
                payload[f"ref_{k}"] = res[k].numpy()
        payload["ref_sparsity_ratio"] = np.array(float(res["sparsity_ratio"]))
        path = os.path.join(args.out, f"ref_{name}.npz")
        np.savez_compressed(path, **payload)
        print("wrote", path, {k: v.shape for k, v in payload.items() if k.startswith("ref_")})


if __name__ == "__main__":
    main()
