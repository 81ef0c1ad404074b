"""GPU parity on the BASELINE.json configurations that round 1 only covered with a reduced model:

  C3  monolayer MoS2 with spin-orbit coupling, DEFAULT irreps (D = 877, l <= 6), soc_basis su2 and so3, nao 19
  C4  Uni-HamGNN call path: nao 26, mixed-Z batch, legacy_edge_update=True, get_nonzero_mask_tensor=True
      (Uni-HamGNN/Uni-HamiltonianPredictor.py:40-76), default irreps
  multi-GPU  edge-sharded forward on the default ('rot') message backend == unsharded forward (2 ranks, NCCL)
  device affinity  a model on cuda:1 while cuda:0 is current (no device guard)

Tolerance 1e-5 relative (max|a-b| / max|b| per tensor) against the fp64 oracle on the same fp32 weights and inputs."""
import os
import subprocess
import sys

import pytest
import torch

from hamgnn_b200 import graph_data as gd
from hamgnn_b200.hamgnn_conv import HamGNNConvE3
from hamgnn_b200.hamgnn_output import HamGNNPlusPlusOut
from hgb_testlib import SMALL_CFG, build_pair, rel_err
from oracle import hamgnn_ref as R

pytestmark = pytest.mark.gpu
TOL = 1e-5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pair(cfg, nao_max, seed=0, **kw):
    torch.manual_seed(seed)
    pre = HamGNNConvE3(cfg)
    D = str(pre.irreps_node_features)
    out = HamGNNPlusPlusOut(D, D, nao_max=nao_max, **kw)
    opre = R.HamGNNConvE3(cfg)
    okw = {k: v for k, v in kw.items() if k != "get_nonzero_mask_tensor"}
    oout = R.HamGNNPlusPlusOut(D, D, nao_max=nao_max, **okw)
    assert not opre.load_state_dict(pre.state_dict(), strict=False).missing_keys
    res = oout.load_state_dict(out.state_dict(), strict=False)
    assert not res.missing_keys and not res.unexpected_keys, res
    return pre, out, opre.double(), oout.double()


def _oracle(opre, oout, batch):
    d = R.AttrDict({k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.to_dict().items()})
    with torch.no_grad():
        rep = opre(d)
        return d, rep, oout(d, rep)


def _run(pre, out, batch, dev):
    pre.to(dev)
    out.to(dev)
    b = gd.Batch(**batch.to_dict()).to(dev)
    with torch.no_grad():
        rep = pre(b)
        o = out(b, rep)
    torch.cuda.synchronize(dev)
    return b, rep, o


@pytest.mark.parametrize("basis", ["su2", "so3"])
def test_c3_mos2_soc_default_irreps(basis):
    dev = torch.device("cuda:0")
    pre, out, opre, oout = _pair({}, 19, soc_switch=True, soc_basis=basis, ham_only=True, add_H0=True)
    batch = gd.Batch.from_data_list([gd.mos2_monolayer(seed=2, soc=True, nao_max=19)])
    d, rep, ref = _oracle(opre, oout, batch)
    b, grep, o = _run(pre, out, batch, dev)
    e_n, e_e = rel_err(grep["node_attr"].cpu(), rep["node_attr"]), rel_err(grep["edge_attr"].cpu(), rep["edge_attr"])
    e_re = rel_err(o["hamiltonian_real"].cpu(), ref["hamiltonian_real"])
    e_im = rel_err(o["hamiltonian_imag"].cpu(), ref["hamiltonian_imag"])
    print(f"C3 MoS2 {basis} default irreps (E={batch.edge_index.shape[1]}): node {e_n:.2e} edge {e_e:.2e} H_re {e_re:.2e} H_im {e_im:.2e}")
    assert max(e_n, e_e, e_re, e_im) < TOL
    assert tuple(o["hamiltonian"].shape) == (2 * (batch.num_nodes + batch.edge_index.shape[1]), 38 * 38)


def test_c4_uni_path_nao26_mixed_z_default_irreps():
    dev = torch.device("cuda:0")
    cfg = dict(legacy_edge_update=True, use_corr_prod=False)
    pre, out, opre, oout = _pair(cfg, 26, soc_switch=False, ham_only=True, add_H0=True, zero_point_shift=True,
                                 get_nonzero_mask_tensor=True)
    gs = [gd.random_mixed_cell(n_atoms=10, species=(1, 6, 8, 14, 42, 16), seed=11, nao_max=26),
          gd.random_mixed_cell(n_atoms=8, species=(3, 7, 31, 33, 83), seed=12, nao_max=26)]
    batch = gd.Batch.from_data_list(gs)
    d, rep, ref = _oracle(opre, oout, batch)
    b, grep, o = _run(pre, out, batch, dev)
    e_n, e_e = rel_err(grep["node_attr"].cpu(), rep["node_attr"]), rel_err(grep["edge_attr"].cpu(), rep["edge_attr"])
    e_h = rel_err(o["hamiltonian"].cpu(), ref["hamiltonian"])
    print(f"C4 nao26 mixed-Z legacy (N={batch.num_nodes}, E={batch.edge_index.shape[1]}): node {e_n:.2e} edge {e_e:.2e} H {e_h:.2e}")
    assert max(e_n, e_e, e_h) < TOL
    m = o["mask"]
    assert m.dtype == torch.bool and tuple(m.shape) == tuple(o["hamiltonian"].shape)
    # single-crystal check of the mask semantics: predicted blocks vanish exactly where the mask is False
    one = gd.Batch.from_data_list(gs[:1])
    b1, _, o1 = _run(pre, HamGNNPlusPlusOut(str(pre.irreps_node_features), str(pre.irreps_node_features), nao_max=26,
                                           soc_switch=False, ham_only=True, add_H0=False, get_nonzero_mask_tensor=True).to(dev),
                     one, dev)
    assert (o1["hamiltonian"][~o1["mask"]] == 0).all() and (o1["hamiltonian"][o1["mask"]] != 0).any()


def test_model_on_second_device_while_first_is_current():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    pre, out, opre, oout = build_pair(SMALL_CFG, nao_max=19, add_H0=True)
    batch = gd.Batch.from_data_list([gd.bulk_silicon()])
    torch.cuda.set_device(0)
    b0, _, o0 = _run(pre, out, batch, torch.device("cuda:0"))
    h0 = o0["hamiltonian"].cpu()
    assert torch.cuda.current_device() == 0
    b1, _, o1 = _run(pre, out, batch, torch.device("cuda:1"))     # current device stays 0: the C ABI switches per call
    assert torch.cuda.current_device() == 0 and o1["hamiltonian"].device.index == 1
    assert torch.equal(o1["hamiltonian"].cpu(), h0)


def test_edge_sharded_forward_two_ranks_default_backend():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    env.pop("HGB_MSGPACK", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "dist_gpu_check.py"), "--default"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    print(res.stdout[-2000:])
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "DIST_CHECK PASS" in res.stdout
