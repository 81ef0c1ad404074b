"""CPU tests of the host planners: the packed tables that drive the CUDA kernels, executed by the
arithmetic emulator (tests/kernel_emulator.py), must reproduce the oracle."""
import numpy as np
import pytest
import torch

from hamgnn_b200 import graph_data as gd
from hamgnn_b200.irreps import Irreps
from hamgnn_b200.plan import tp_paths
import hgb_kernel_emulator as EM
from hgb_testlib import SMALL_CFG, build_pair, oracle_forward, rel_err


def test_default_mid_irreps_match_survey():
    D = Irreps("64x0e+64x0o+32x1o+16x1e+12x2o+25x2e+18x3o+9x3e+4x4o+9x4e+4x5o+4x5e+2x6e")
    sh = Irreps("0e + 1o + 2e + 3o + 4e + 5o")
    mid, paths = tp_paths(D.scaled(2), sh, D)
    assert str(mid.simplify()) == ("384x0o+384x0e+512x1o+240x1e+264x2o+550x2e+468x3o+225x3e+104x4o+234x4e+96x5o+"
                                   "92x5e+36x6e")
    assert len(paths) == 255 and sum(p.mul_in_total * p.mul_out for p in paths) == 118482
    mid_e, paths_e = tp_paths(D, sh, D)
    assert sum(p.mul_in_total * p.mul_out for p in paths_e) == 59241 and mid_e.dim == 17523


@pytest.fixture(scope="module")
def small():
    pre, out, opre, oout = build_pair(SMALL_CFG)
    g = gd.Batch.from_data_list([gd.bulk_silicon(), gd.graphene(rep=(2, 2, 1), seed=1)])
    d, rep, res = oracle_forward(opre, oout, g)
    return pre, out, opre, oout, g, d, rep, res


def _dbl(t):
    return t.detach().double()


def test_conv_message_and_scatter(small):
    pre, out, opre, oout, g, d, rep, res = small
    # re-run the oracle up to the first conv to capture its inputs
    from oracle import hamgnn_ref as R
    dd = R.AttrDict({k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in g.to_dict().items()})
    with torch.no_grad():
        onehot = torch.nn.functional.one_hot(dd["z"], opre.num_types).double()
        dd["node_attrs"] = dd["node_features"] = onehot
        opre.edge_geometry(dd)
        opre.pair_embedding(dd)
        dd["node_features"] = opre.chemical_embedding.linear(onehot)
        x, e = dd["node_features"].clone(), dd["edge_features"].clone()
        conv = opre.convolutions[0]
        s, r = dd["edge_index"]
        m_ref = conv.conv_tp(x[s], x[r], e, dd["edge_attrs"], dd["edge_embedding"])
    blk = pre.convolutions[0].conv_tp
    st, wbuf = blk.op.pack(blk.weights())
    m = EM.emulate_msgpack(blk.op, wbuf.double(), [x, x, e], [s, r, None], dd["edge_attrs"], dd["edge_embedding"])
    assert rel_err(m, m_ref) < 2e-6
    agg = EM.emulate_msgpack(blk.op, wbuf.double(), [x, x, e], [s, r, None], dd["edge_attrs"], dd["edge_embedding"],
                             out_rows=r, n_out=x.shape[0])
    agg_ref = torch.zeros_like(x).index_add_(0, r, m_ref)
    assert rel_err(agg, agg_ref) < 2e-6


def test_embedding_block(small):
    pre, out, opre, oout, g, d, rep, res = small
    from oracle import hamgnn_ref as R
    dd = R.AttrDict({k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in g.to_dict().items()})
    with torch.no_grad():
        onehot = torch.nn.functional.one_hot(dd["z"], opre.num_types).double()
        dd["node_features"] = onehot
        opre.edge_geometry(dd)
        ref = opre.pair_embedding(dd)
    pe = pre.pair_embedding
    s, r = dd["edge_index"]
    h = EM.emulate_linear(pe.up_op, _dbl(pe.linear_up_src.weight), onehot[s]) + \
        EM.emulate_linear(pe.up_op, _dbl(pe.linear_up_dst.weight), onehot[r])
    st, wbuf = pe.conv_tp.op.pack(pe.conv_tp.weights())
    got = EM.emulate_msgpack(pe.conv_tp.op, wbuf.double(), [h], [None], dd["edge_attrs"], dd["edge_embedding"])
    assert rel_err(got, ref) < 2e-6


def test_pair_block_with_skip(small):
    pre, out, opre, oout, g, d, rep, res = small
    torch.manual_seed(3)
    E, N, D = g.edge_index.shape[1], g.num_nodes, pre.irreps_node_features.dim
    x, e = torch.randn(N, D).double(), torch.randn(E, D).double()
    dd = {"edge_index": g.edge_index, "node_features": x, "edge_features": e, "edge_attrs": d["edge_attrs"],
          "edge_embedding": d["edge_embedding"]}
    with torch.no_grad():
        ref = opre.pair_interactions[1](dict(dd))
    pb = pre.pair_interactions[1]
    xs = EM.emulate_linear(pb.up_op, _dbl(pb.linear_up_src.weight), x)
    xt = EM.emulate_linear(pb.up_op, _dbl(pb.linear_up_tar.weight), x)
    st, wbuf = pb.conv_tp.op.pack(pb.conv_tp.weights(pb.skip_linear.weight))
    s, r = g.edge_index
    got = EM.emulate_msgpack(pb.conv_tp.op, wbuf.double(), [xs, xt, e], [s, r, None], d["edge_attrs"], d["edge_embedding"])
    assert rel_err(got, ref) < 2e-6


def test_resblock_and_head(small):
    pre, out, opre, oout, g, d, rep, res = small
    torch.manual_seed(4)
    x = torch.randn(7, pre.irreps_node_features.dim).double()
    with torch.no_grad():
        ref = opre.convolutions[0].residual(x)
        ref_head = oout.offsite_hamiltonian_network(x)
    got = EM.emulate_resblock(pre.convolutions[0].residual, x)
    assert rel_err(got, ref) < 2e-6
    hl = out.offsite_hamiltonian_network
    got_head = EM.emulate_resblock(hl.residual_block, x, post=hl.op, post_w=hl.linear_transform.weight)
    assert rel_err(got_head, ref_head) < 2e-6


def test_ham_assembly(small):
    pre, out, opre, oout, g, d, rep, res = small
    torch.manual_seed(5)
    E, N = g.edge_index.shape[1], g.num_nodes
    coef_on = torch.randn(N, out.assembly.n_coef).double()
    coef_off = torch.randn(E, out.assembly.n_coef).double()
    on_row, off_row, inv = out._row_maps(g)
    with torch.no_grad():
        Hon = oout._sym(oout.reorder_matrix(oout.merge_tensor_components(torch.split(coef_on, oout.ham_dims, -1)))) + d["Hon0"]
        Hoff = oout._sym(oout.reorder_matrix(oout.merge_tensor_components(torch.split(coef_off, oout.ham_dims, -1))), inv) + d["Hoff0"]
        Hon, Hoff = oout._mask(Hon, Hoff, d)
    s, r = g.edge_index
    got_on = EM.emulate_ham(out.assembly, coef_on, None, d["Hon0"], g.z, None, None)
    got_off = EM.emulate_ham(out.assembly, coef_off, inv, d["Hoff0"], g.z, s, r)
    assert rel_err(got_on, Hon) < 1e-6 and rel_err(got_off, Hoff) < 1e-6
    # interleaved per-crystal rows
    ref = oout.concat_by_crystal(d, Hon, Hoff)
    H = torch.empty_like(ref)
    H[on_row] = got_on
    H[off_row] = got_off
    assert rel_err(H, ref) < 1e-6


def test_tensor_core_packing_matches_oracle(small):
    """The hi|lo operand images consumed by the tcgen05 kernel decode to the same operator as the oracle."""
    pre, out, opre, oout, g, d, rep, res = small
    torch.manual_seed(6)
    E, N, D = g.edge_index.shape[1], g.num_nodes, pre.irreps_node_features.dim
    x, e = torch.randn(N, D).double(), torch.randn(E, D).double()
    dd = {"edge_index": g.edge_index, "node_features": x, "edge_features": e, "edge_attrs": d["edge_attrs"],
          "edge_embedding": d["edge_embedding"]}
    s, r = g.edge_index
    with torch.no_grad():
        ref_pair = opre.pair_interactions[1](dict(dd))
        ref_msg = opre.convolutions[0].conv_tp(x[s], x[r], e, d["edge_attrs"], d["edge_embedding"])
    pb = pre.pair_interactions[1]
    xs = EM.emulate_linear(pb.up_op, _dbl(pb.linear_up_src.weight), x)
    xt = EM.emulate_linear(pb.up_op, _dbl(pb.linear_up_tar.weight), x)
    st = pb.conv_tp.op.pack_tc(pb.conv_tp.weights(pb.skip_linear.weight))
    got = EM.emulate_msgpack_tc(pb.conv_tp.op, st["tc_wbuf"].double(), [xs, xt, e], [s, r, None], d["edge_attrs"], d["edge_embedding"])
    assert rel_err(got, ref_pair) < 2e-6
    cb = pre.convolutions[0].conv_tp
    st = cb.op.pack_tc(cb.weights())
    got = EM.emulate_msgpack_tc(cb.op, st["tc_wbuf"].double(), [x, x, e], [s, r, None], d["edge_attrs"], d["edge_embedding"])
    assert rel_err(got, ref_msg) < 2e-6
    pe = pre.pair_embedding
    onehot = torch.nn.functional.one_hot(g.z, opre.num_types).double()
    h = EM.emulate_linear(pe.up_op, _dbl(pe.linear_up_src.weight), onehot[s]) + EM.emulate_linear(pe.up_op, _dbl(pe.linear_up_dst.weight), onehot[r])
    st = pe.conv_tp.op.pack_tc(pe.conv_tp.weights())
    got = EM.emulate_msgpack_tc(pe.conv_tp.op, st["tc_wbuf"].double(), [h], [None], d["edge_attrs"], d["edge_embedding"])
    dd2 = dict(dd); dd2["node_features"] = onehot
    with torch.no_grad():
        ref_emb = opre.pair_embedding(dd2)
    assert rel_err(got, ref_emb) < 2e-6


def test_rotated_frame_program_matches_oracle(small):
    """The edge-aligned ('rot') pipeline -- per-edge Wigner matrices built the way wigner_kernel builds them, rotated +
    packed operand images, the (path, m1) step tables and the rotation back -- reproduces the oracle's messages."""
    pre, out, opre, oout, g, d, rep, res = small
    torch.manual_seed(7)
    E, N, D = g.edge_index.shape[1], g.num_nodes, pre.irreps_node_features.dim
    x, e = torch.randn(N, D).double(), torch.randn(E, D).double()
    dd = {"edge_index": g.edge_index, "node_features": x, "edge_features": e, "edge_attrs": d["edge_attrs"],
          "edge_embedding": d["edge_embedding"]}
    s, r = g.edge_index
    vec = d["edge_vectors"].double().numpy()
    with torch.no_grad():
        ref_pair = opre.pair_interactions[1](dict(dd))
        ref_msg = opre.convolutions[0].conv_tp(x[s], x[r], e, d["edge_attrs"], d["edge_embedding"])
    # the Wigner matrices take Y(edge) to the polar axis: D^l Y_l = sqrt(2l+1) e_{m=0}, and are orthogonal
    cb = pre.convolutions[0].conv_tp
    Dw = EM.emulate_wigner(vec, cb.op)
    from hamgnn_b200 import so3
    for l in range(cb.op.rot_lmax + 1):
        dl = 2 * l + 1
        Dl = Dw[:, cb.op.rot_doff[l]:cb.op.rot_doff[l] + dl * dl].reshape(E, dl, dl)
        y0 = np.einsum("emi,ei->em", Dl, so3.real_sh(l, vec))
        want = np.zeros(dl); want[l] = np.sqrt(dl)
        assert np.abs(y0 - want).max() < 1e-12
        assert np.abs(np.einsum("emi,eni->emn", Dl, Dl) - np.eye(dl)).max() < 1e-12
    pb = pre.pair_interactions[1]
    xs = EM.emulate_linear(pb.up_op, _dbl(pb.linear_up_src.weight), x)
    xt = EM.emulate_linear(pb.up_op, _dbl(pb.linear_up_tar.weight), x)
    st = pb.conv_tp.op.pack_tc(pb.conv_tp.weights(pb.skip_linear.weight))
    got = EM.emulate_msgpack_rot(pb.conv_tp.op, st["tc_wbuf"].double(), [xs, xt, e], [s, r, None], vec, d["edge_embedding"])
    assert rel_err(got, ref_pair) < 2e-6
    st = cb.op.pack_tc(cb.weights())
    got = EM.emulate_msgpack_rot(cb.op, st["tc_wbuf"].double(), [x, x, e], [s, r, None], vec, d["edge_embedding"])
    assert rel_err(got, ref_msg) < 2e-6
    pe = pre.pair_embedding
    onehot = torch.nn.functional.one_hot(g.z, opre.num_types).double()
    h = EM.emulate_linear(pe.up_op, _dbl(pe.linear_up_src.weight), onehot[s]) + EM.emulate_linear(pe.up_op, _dbl(pe.linear_up_dst.weight), onehot[r])
    st = pe.conv_tp.op.pack_tc(pe.conv_tp.weights())
    got = EM.emulate_msgpack_rot(pe.conv_tp.op, st["tc_wbuf"].double(), [h], [None], vec, d["edge_embedding"])
    dd2 = dict(dd); dd2["node_features"] = onehot
    with torch.no_grad():
        ref_emb = opre.pair_embedding(dd2)
    assert rel_err(got, ref_emb) < 2e-6
    # ---- msgpack_rotf_kernel reads L' as the un-split fp32 row-major image at hgb_rot_step_t.pad2: same matrix as the (hi | lo) image
    for op_, st_ in ((pb.conv_tp.op, pb.conv_tp.op.pack_tc(pb.conv_tp.weights(pb.skip_linear.weight))), (cb.op, cb.op.pack_tc(cb.weights()))):
        wb = st_["tc_wbuf"].double()
        for t in range(len(op_.irreps_out)):
            mp = op_.tc_types_c[t].mpad
            for si in range(op_.rot_step_begin[t], op_.rot_step_begin[t + 1]):
                s_ = op_.rot_steps_c[si]
                assert s_.pad2 > 0 and s_.pad2 % 4 == 0
                plain = wb[s_.pad2:s_.pad2 + mp * mp].view(mp, mp)
                assert torch.allclose(plain, EM._decode_image(wb, s_.lf_off, mp, mp), rtol=0, atol=1e-6 * float(plain.abs().max()) + 1e-30)
    # ---- the fp16 x 2 split program (rot16 tables, packed fp16 images with power-of-two scales): same bar
    st16 = pe.conv_tp.op.pack_rot16(pe.conv_tp.weights())
    got16 = EM.emulate_msgpack_rot16(pe.conv_tp.op, st16, [h], [None], vec, d["edge_embedding"])
    assert rel_err(got16, ref_emb) < 2e-6
    st16 = pb.conv_tp.op.pack_rot16(pb.conv_tp.weights(pb.skip_linear.weight))
    got16 = EM.emulate_msgpack_rot16(pb.conv_tp.op, st16, [xs, xt, e], [s, r, None], vec, d["edge_embedding"])
    assert rel_err(got16, ref_pair) < 2e-6
    st16 = cb.op.pack_rot16(cb.weights())
    got16 = EM.emulate_msgpack_rot16(cb.op, st16, [x, x, e], [s, r, None], vec, d["edge_embedding"])
    assert rel_err(got16, ref_msg) < 2e-6
    # ---- the A-stationary regrouping of the same steps (rot2 tables) + the segmented receiver reduction
    got2 = EM.emulate_msgpack_rot2(pe.conv_tp.op, st["tc_wbuf"].double(), [h], [None], vec, d["edge_embedding"])
    assert rel_err(got2, ref_emb) < 2e-6
    st = pb.conv_tp.op.pack_tc(pb.conv_tp.weights(pb.skip_linear.weight))
    got2 = EM.emulate_msgpack_rot2(pb.conv_tp.op, st["tc_wbuf"].double(), [xs, xt, e], [s, r, None], vec, d["edge_embedding"])
    assert rel_err(got2, ref_pair) < 2e-6
    st = cb.op.pack_tc(cb.weights())
    from hamgnn_b200.plan import receiver_segments
    seg_ptr, seg_order = receiver_segments(r, N)
    assert torch.equal(r[seg_order], torch.sort(r, stable=True).values) and int(seg_ptr[-1]) == E
    got2 = EM.emulate_msgpack_rot2(cb.op, st["tc_wbuf"].double(), [x, x, e], [s, r, None], vec, d["edge_embedding"],
                                   seg_ptr=seg_ptr, seg_order=seg_order)
    assert rel_err(got2, torch.zeros(N, D, dtype=torch.float64).index_add_(0, r, ref_msg)) < 2e-6


def test_rotated_frame_tables_are_well_formed():
    """Structure the kernel relies on: steps grouped by output component with the group-end flag on the last one,
    operand images inside the packed tile, 16-byte aligned offsets, the identity L' image, TMEM budget, J orthogonal."""
    import ctypes as C
    from hamgnn_b200 import lib as L, so3
    from hamgnn_b200.plan import Branch, MessagePackOp
    D = Irreps("64x0e+64x0o+32x1o+16x1e+12x2o+25x2e+18x3o+9x3e+4x4o+9x4e+4x5o+4x5e+2x6e")
    op = MessagePackOp([Branch(D, 2, 0), Branch(D, 1, 2)], "0e+1o+2e+3o+4e+5o", D, 64, [64, 64], src_dims=[877, 877, 877], direct_src=2)
    assert C.sizeof(L.RotStepT) == 32 and C.sizeof(L.RotBlockT) == 32 and C.sizeof(L.RotPlan) == 240
    assert op.rot_supported() and op.rot_lmax == 6 and op.rot_dstride == 476 and op.rot_doff == [0, 4, 16, 44, 96, 180, 304]
    assert op.rot_tile_stride % 1024 == 0
    T = op.ROT_TILE
    for b in op.rot_blocks_c[:op.rot_n_blocks]:
        assert b.kpad % 8 == 0 and b.kpad >= b.nsrc * b.mul and b.xoff % 4 == 0
        assert b.xoff + (2 * b.l1 + 1) * 2 * b.kpad * T <= op.rot_tile_stride
    n_direct = 0
    for t in range(len(op.irreps_out)):
        ty = op.tc_types_c[t]
        d3 = 2 * ty.l + 1
        steps = [op.rot_steps_c[i] for i in range(op.rot_step_begin[t], op.rot_step_begin[t + 1])]
        assert steps, t
        last_m3, open_group = -1, False
        for st in steps:
            assert st.kind == 0 and (st.new_path & 1) and 0 <= st.m3 < d3 and st.kpad % 8 == 0
            assert st.a_off % 4 == 0 and st.w_off % 4 == 0 and st.lf_off % 4 == 0
            assert st.a_off + 2 * st.kpad * T <= op.rot_tile_stride
            assert (st.m3 == last_m3) if open_group else (st.m3 > last_m3)
            last_m3, open_group = st.m3, not (st.new_path & 4)
            if st.branch < 0:
                n_direct += 1
                assert st.lf_off == op.tc_ident_off[int(ty.mpad)] and st.scale == 1.0
            else:
                assert 0 <= st.g_off and st.g_off + ty.mul <= op.n_channels[st.branch] and st.scale != 0.0
        assert not open_group
        dbl = 0 if 16 < ty.mpad <= 32 else 1
        assert (4 + 2 * dbl) * ty.mpad + d3 * ty.mul <= (512 if ty.mpad > 32 else 256)      # 2 CTAs/SM for the l >= 1 slots
    assert n_direct == sum(m.ir.dim for m in D)                                       # one un-gated step per (slot, output component)
    assert op.rot_n_steps == 2883
    for l in range(op.rot_lmax + 1):
        J = so3.wigner_J(l)
        assert np.abs(J @ J.T - np.eye(2 * l + 1)).max() < 1e-12
    # the identity images decode to I
    w = {"tp": [torch.randn(n) for n in op.tp_numel], "fc": [[torch.randn(64, 64), torch.randn(64, 64), torch.randn(64, c)] for c in op.n_channels],
         "lin_mid": [torch.randn(b[1]) for b in op.lin_mid_blocks], "lin_out": [torch.randn(b[1]) for b in op.lin_out_blocks],
         "direct": torch.randn(op.direct_blocks[1])}
    st = op.pack_tc(w)
    for mp, off in op.tc_ident_off.items():
        assert torch.equal(EM._decode_image(st["tc_wbuf"], off, mp, mp), torch.eye(mp))
    # the un-split fp32 copy of every step's L' agrees with its (hi | lo) image
    wb = st["tc_wbuf"]
    for t in range(len(op.irreps_out)):
        mp = int(op.tc_types_c[t].mpad)
        for i in range(op.rot_step_begin[t], op.rot_step_begin[t + 1], 7):
            s_ = op.rot_steps_c[i]
            assert s_.pad2 % 4 == 0 and s_.pad2 + mp * mp <= op.tc_w_total
            plain = wb[s_.pad2:s_.pad2 + mp * mp].view(mp, mp)
            img = EM._decode_image(wb, s_.lf_off, mp, mp)
            assert float((plain - img).abs().max()) <= 3e-7 * float(plain.abs().max()) + 1e-30


def test_radial_gate_tiles_match_oracle(small):
    """The W3 tiles of the tensor-core gate pre-pass decode to the oracle's FullyConnectedNet (both branches)."""
    pre, out, opre, oout, g, d, rep, res = small
    cb, ocb = pre.convolutions[1].conv_tp, opre.convolutions[1].conv_tp
    st = cb.op.pack_tc(cb.weights())
    rbf = d["edge_embedding"]
    got = EM.emulate_radial_gate_tc(cb.op, st["tc_wbuf"].double(), rbf)
    with torch.no_grad():
        ref_n, ref_e = ocb.node_weight_generator(rbf), ocb.edge_weight_generator(rbf)
    assert rel_err(got[0, :, :ref_n.shape[1]], ref_n) < 2e-6 and rel_err(got[1, :, :ref_e.shape[1]], ref_e) < 2e-6


@pytest.mark.parametrize("cfg_name", ["small", "default"])
def test_tensor_core_launcher_accepts_every_message_plan(cfg_name):
    """Host-side validation of hgb_msgpack_tcg_forward (slot classes, staging-buffer limits of both message kernels)
    runs before any CUDA call, so it can be exercised without a GPU: with dummy device pointers the call must get
    past validation and fail only at the first CUDA runtime call."""
    import ctypes as C
    from hamgnn_b200 import lib as L, plan as PL, so3
    from hamgnn_b200.hamgnn_conv import HamGNNConvE3
    if torch.cuda.is_available():
        pytest.skip("with a GPU the dummy pointers would be dereferenced; the GPU parity tests cover this path")
    pre = HamGNNConvE3(dict(SMALL_CFG) if cfg_name == "small" else {})
    ops = [v for m in pre.modules() for v in vars(m).values() if isinstance(v, PL.MessagePackOp)]
    assert ops
    lib = L.load()
    dummy = 0x10000
    for op in ops:
        if not op.tc_supported():
            continue
        plan = L.MsgpackPlan()
        plan.n_types, plan.n_paths = len(op.irreps_out), op.n_paths
        plan.n_branches, plan.n_sources = len(op.branches), len(op.src_dims)
        plan.sh_dim, plan.rbf_dim, plan.h1, plan.h2 = op.irreps_sh.dim, op.rbf_dim, op.h1, op.h2
        plan.out_dim = op.irreps_out.dim
        for q, d in enumerate(op.src_dims):
            plan.src_dim[q] = d
        plan.act_const = so3.normalize2mom_const("silu")
        plan.types = plan.paths = plan.cg_ij = plan.cg_val = plan.cg_kstart = plan.wbuf = dummy
        plan.types_host = C.cast(op.tc_types_c, C.c_void_p).value
        plan.paths_host = C.cast(op.tc_paths_c, C.c_void_p).value
        ns, nb = len(op.src_dims), len(op.branches)
        srcs = (C.c_void_p * 4)(*([dummy] * ns + [None] * (4 - ns)))
        rws = (C.c_void_p * 4)(*([None] * 4))
        gstride = (max(op.n_channels) + 3) // 4 * 4
        w3o = (C.c_int32 * 2)(*(list(op.tc_w3_off) + [0] * (2 - nb)))
        nch = (C.c_int32 * 2)(*(list(op.n_channels) + [0] * (2 - nb)))
        rc = lib.hgb_msgpack_tcg_forward(C.byref(plan), srcs, rws, dummy, dummy, w3o, nch, gstride, dummy, 1000, dummy, None, None)
        msg = lib.hgb_last_error().decode()
        assert rc != 0 and "failed" in msg and "hgb_msgpack_tcg_forward" not in msg, msg
