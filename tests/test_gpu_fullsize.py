"""Parity at BASELINE.json's full size (twisted bilayer graphene m = 28, N = 9 748, E = 778 772) through
size-independent properties of the fused message call -- the oracle cannot run 7.8e5 edges, but

  * messages are per-edge independent, so a sample of edges (first / last tile, both sides of every chunk boundary,
    random interior) is compared with the oracle's MessagePackBlock evaluated on exactly those rows (1e-5, fp64 oracle);
  * the edge-update form has no atomics, so the result must not depend on the edge-chunk size: bit-identical;
  * the scatter form must equal an index_add of the per-edge messages (1e-5: atomics order).

This is the only test that drives tile / chunk addressing beyond 2^31 bytes."""
import pytest
import torch

from hamgnn_b200 import graph_data as gd
from hamgnn_b200 import plan as P
from hgb_testlib import DEFAULT_CFG, build_pair, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_message_call_at_full_size():
    dev = torch.device("cuda:0")
    pre, out, opre, oout = build_pair(DEFAULT_CFG, nao_max=19, add_H0=False)
    g = gd.twisted_bilayer_graphene(m=28, seed=0, nao_max=19)
    batch = gd.Batch.from_data_list([g])
    E, N, D = batch.edge_index.shape[1], batch.num_nodes, pre.irreps_node_features.dim
    assert E > 700_000 and N == 9748
    pre.to(dev)
    b = gd.Batch(**{k: v for k, v in batch.to_dict().items() if k not in ("Hon", "Hoff", "Son", "Soff", "cell_shift")}).to(dev)
    with torch.no_grad():
        pre.edge_embed(b)
    sh, rbf, vec = b["edge_attrs"], b["edge_embedding"], b["edge_vectors"]
    gen = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn(N, D, generator=gen).to(dev)
    e = torch.randn(E, D, generator=gen).to(dev)
    sd, rd = b["edge_index"][0].contiguous(), b["edge_index"][1].contiguous()
    cb = pre.convolutions[0].conv_tp
    old = (P.BACKEND, P.ROT_CHUNK_EDGES)
    try:
        P.BACKEND = "rot"
        assert cb.op.rot_supported()
        msg = torch.empty(E, D, device=dev)
        cb.op.forward(cb.weights(), [x, x, e], [sd, rd, None], sh, rbf, E, msg, edge_vec=vec)
        chunk_a = P.ROT_CHUNK_EDGES or P._AUTO_CHUNK[str(msg.device)]   # HGB_ROT_CHUNK, or the size derived from free memory
        agg = torch.zeros(N, D, device=dev)
        cb.op.forward(cb.weights(), [x, x, e], [sd, rd, None], sh, rbf, E, agg, out_index=rd, edge_vec=vec)
        chunk_b = 65536 if chunk_a != 65536 else 131072
        P.ROT_CHUNK_EDGES = chunk_b
        msg_b = torch.empty(E, D, device=dev)
        cb.op.forward(cb.weights(), [x, x, e], [sd, rd, None], sh, rbf, E, msg_b, edge_vec=vec)
        torch.cuda.synchronize()
    finally:
        P.BACKEND, P.ROT_CHUNK_EDGES = old
    assert bool(torch.isfinite(msg).all())
    # (2) chunk-size independence, bit for bit
    assert torch.equal(msg, msg_b), float((msg - msg_b).abs().max())
    # (3) scatter form == index_add of the messages
    ref_agg = torch.zeros(N, D, device=dev, dtype=torch.float64).index_add_(0, rd, msg.double())
    err_sc = rel_err(agg.cpu(), ref_agg.cpu())
    # (1) sampled edges against the oracle
    idx = set(range(0, 130)) | set(range(E - 130, E))
    for c in (chunk_a, chunk_b):
        for k in range(c, E, c):
            idx |= set(range(max(0, k - 3), min(E, k + 3)))
    idx |= set(torch.randint(0, E, (256,), generator=gen).tolist())
    sel = torch.tensor(sorted(idx), dtype=torch.long)
    sel_d = sel.to(dev)
    s_c, r_c = sd[sel_d].cpu(), rd[sel_d].cpu()
    xc, ec = x.cpu().double(), e[sel_d].cpu().double()
    with torch.no_grad():
        ref = opre.double().convolutions[0].conv_tp(xc[s_c], xc[r_c], ec, sh[sel_d].cpu().double(), rbf[sel_d].cpu().double())
    err_msg = rel_err(msg[sel_d].cpu(), ref)
    print(f"[tbg_m28 rot] E={E} sampled edges {len(sel)}: rel err message {err_msg:.2e}, scatter vs index_add {err_sc:.2e}, "
          f"chunks {chunk_a} / {chunk_b} bit-identical")
    assert err_msg < TOL and err_sc < TOL
