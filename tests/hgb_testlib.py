"""Shared helpers for the test-suite: small model configs, oracle <-> product weight transfer."""
import torch

from hamgnn_b200 import graph_data as gd
from hamgnn_b200.hamgnn_conv import HamGNNConvE3
from hamgnn_b200.hamgnn_output import HamGNNPlusPlusOut
from oracle import hamgnn_ref as R

SMALL_CFG = dict(irreps_node_features="8x0e+8x0o+4x1o+4x1e+3x2o+5x2e+2x3o+2x3e+2x4e", num_layers=2, num_radial=16,
                 radial_MLP=[16, 16], irreps_edge_sh="0e+1o+2e+3o+4e", cutoff=26.0)
DEFAULT_CFG = dict()


def build_pair(cfg, nao_max=19, add_H0=True, seed=0, legacy=False):
    """Product modules (fp32, CPU) and oracle modules carrying the same weights."""
    torch.manual_seed(seed)
    cfg = dict(cfg)
    if legacy:
        cfg["legacy_edge_update"] = True
    pre = HamGNNConvE3(cfg)
    out = HamGNNPlusPlusOut(pre.irreps_node_features, pre.irreps_node_features, nao_max=nao_max, soc_switch=False,
                            ham_only=True, add_H0=add_H0)
    opre = R.HamGNNConvE3(cfg)
    oout = R.HamGNNPlusPlusOut(str(pre.irreps_node_features), str(pre.irreps_node_features), nao_max=nao_max, add_H0=add_H0)
    missing, unexpected = opre.load_state_dict(pre.state_dict(), strict=False)
    assert not missing, missing
    m2, u2 = oout.load_state_dict(out.state_dict(), strict=False)
    assert not m2, m2
    return pre, out, opre, oout


def oracle_forward(opre, oout, batch, dtype=torch.float64):
    d = R.AttrDict({k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.to_dict().items()})
    opre_d, oout_d = opre.to(dtype), oout.to(dtype)
    with torch.no_grad():
        rep = opre_d(d)
        res = oout_d(d, rep)
    return d, rep, res


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
