"""Generates tests/golden/si_small.npz: seeded inputs, weights and ORACLE outputs (fp64 oracle evaluated on
fp32 weights) for bulk Si + the small test model.  The reference itself cannot be imported here (e3nn et al.
are absent), so these vectors pin the oracle/product pair against regressions; they are not
reference outputs.      python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from hamgnn_b200 import graph_data as gd  # noqa: E402
from hgb_testlib import SMALL_CFG, build_pair, oracle_forward  # noqa: E402


def main():
    pre, out, opre, oout = build_pair(SMALL_CFG, nao_max=19, add_H0=True, seed=0)
    g = gd.bulk_silicon(seed=0)
    batch = gd.Batch.from_data_list([g])
    d, rep, res = oracle_forward(opre, oout, batch)
    payload = {f"in_{k}": v.numpy() for k, v in g.to_dict().items() if torch.is_tensor(v)}
    payload.update({f"pre_{k}": v.detach().numpy() for k, v in pre.state_dict().items()})
    payload.update({f"out_{k}": v.detach().numpy() for k, v in out.state_dict().items()})
    payload["ref_hamiltonian"] = res["hamiltonian"].numpy().astype(np.float64)
    payload["ref_node_attr"] = rep["node_attr"].numpy().astype(np.float32)
    payload["ref_edge_attrs"] = d["edge_attrs"].numpy().astype(np.float32)
    payload["ref_edge_embedding"] = d["edge_embedding"].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "si_small.npz"), **payload)
    print("wrote", os.path.join(HERE, "si_small.npz"))


if __name__ == "__main__":
    main()
