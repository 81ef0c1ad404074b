"""GPU test of hamgnn_b200.pipeline.streamed_forward (the caller-side loop of the hot path; the reference's per-batch device
transfer is Lightning's transfer_batch_to_device around hamgnn/models/Model.py:128-179): results of the streamed pipeline --
host -> device and device -> host copies on their own streams -- are bit-identical to the plain forward of every batch, in order."""
import pytest
import torch

from hamgnn_b200 import graph_data as gd
from hamgnn_b200.pipeline import streamed_forward
from hgb_testlib import SMALL_CFG, build_pair

pytestmark = pytest.mark.gpu


def test_streamed_forward_matches_plain_forward():
    pre, out, _opre, _oout = build_pair(SMALL_CFG, nao_max=19, add_H0=False)
    dev = torch.device("cuda:0")
    pre.to(dev)
    out.to(dev)
    graphs = [[gd.bulk_silicon()], [gd.graphene(rep=(2, 2, 1), seed=1), gd.bulk_silicon()], [gd.mos2_monolayer(seed=2)],
              [gd.graphene(rep=(3, 3, 1), seed=3)], [gd.bulk_silicon()]]
    drop = ("Hon", "Hoff", "Son", "Soff", "cell_shift", "iHon", "iHoff", "edge_global_idx")
    hosts = []
    for gs in graphs:
        b = gd.Batch.from_data_list(gs)
        hosts.append(gd.Batch(**{k: v for k, v in b.to_dict().items() if k not in drop}).pin_memory())
    want = []
    with torch.no_grad():
        for h in hosts:
            b = gd.Batch(**h.to_dict()).to(dev)
            want.append(out(b, pre(b))["hamiltonian"].cpu())
    seen = []
    got = [h.clone() for h in streamed_forward(pre, out, hosts, device=dev, on_result=lambda i, t: seen.append(i))]
    assert seen == list(range(len(hosts)))
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.equal(a, b)
    # an empty iterator yields nothing
    assert list(streamed_forward(pre, out, [], device=dev)) == []
