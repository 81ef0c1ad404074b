"""Committed golden vectors (tests/golden/si_small.npz, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from hamgnn_b200 import graph_data as gd
from hgb_testlib import SMALL_CFG, build_pair, oracle_forward, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden", "si_small.npz")


def _load():
    z = np.load(GOLD)
    g = gd.Data(**{k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in_")})
    pre, out, opre, oout = build_pair(SMALL_CFG, nao_max=19, add_H0=True, seed=123)  # different init, then overwrite
    pre.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("pre_")})
    out.load_state_dict({k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("out_")})
    opre.load_state_dict(pre.state_dict(), strict=False)
    oout.load_state_dict(out.state_dict(), strict=False)
    return z, g, pre, out, opre, oout


def test_oracle_reproduces_golden():
    z, g, pre, out, opre, oout = _load()
    d, rep, res = oracle_forward(opre, oout, gd.Batch.from_data_list([g]))
    assert rel_err(res["hamiltonian"], torch.from_numpy(z["ref_hamiltonian"])) < 1e-12
    assert rel_err(rep["node_attr"], torch.from_numpy(z["ref_node_attr"])) < 1e-6


@pytest.mark.gpu
def test_cuda_path_reproduces_golden():
    z, g, pre, out, opre, oout = _load()
    dev = torch.device("cuda:0")
    pre.to(dev)
    out.to(dev)
    b = gd.Batch.from_data_list([g]).to(dev)
    with torch.no_grad():
        rep = pre(b)
        res = out(b, rep)
    assert rel_err(b["edge_attrs"].cpu(), torch.from_numpy(z["ref_edge_attrs"])) < 1e-5
    assert rel_err(b["edge_embedding"].cpu(), torch.from_numpy(z["ref_edge_embedding"])) < 1e-5
    assert rel_err(res["hamiltonian"].cpu(), torch.from_numpy(z["ref_hamiltonian"])) < 1e-5
