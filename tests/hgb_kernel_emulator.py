"""TEST INFRASTRUCTURE -- CPU emulation of the *arithmetic* of the CUDA kernels, driven by the very same
packed tables / ctypes structs that are uploaded to the GPU (types, paths, sparse CG lists, wbuf, linear
blocks, gate descriptor, Hamiltonian CSR).  It lets the CPU test-suite check the host planners against the
oracle without a GPU; it is never used by the product path (hamgnn_b200/ has no CPU fallback).
"""
from __future__ import annotations

import numpy as np
import torch

from hamgnn_b200 import lib as L
from hamgnn_b200.plan import GateLayout, HamAssembly, LinearOp, MessagePackOp


def _silu(x):
    return x / (1.0 + torch.exp(-x))


def emulate_msgpack(op: MessagePackOp, wbuf: torch.Tensor, sources, rows, sh, rbf, out_rows=None, n_out=None):
    """Mirrors msgpack_kernel: T -> A -> (A W) * g -> L' per path, summed per output slot."""
    from hamgnn_b200 import so3
    E = sh.shape[0]
    dt = wbuf.dtype
    D = op.irreps_out.dim
    msg = torch.zeros(E, D, dtype=dt)
    act = so3.normalize2mom_const("silu")
    h2 = []
    for b in range(len(op.branches)):
        w1 = wbuf[op.fc1_off[b]:op.fc1_off[b] + op.rbf_dim * op.h1].view(op.rbf_dim, op.h1)
        w2 = wbuf[op.fc2_off[b]:op.fc2_off[b] + op.h1 * op.h2].view(op.h1, op.h2)
        h = _silu(rbf @ w1) * act
        h2.append(_silu(h @ w2) * act)
    gathered = [s if r is None else s[r] for s, r in zip(sources, rows)]
    for t in range(len(op.irreps_out)):
        ty = op.types_c[t]
        d3 = 2 * ty.l + 1
        acc = torch.zeros(E, d3, ty.mpad, dtype=dt)
        for p in range(ty.path_begin, ty.path_end):
            pa = op.paths_c[p]
            d1 = 2 * pa.l1 + 1
            K = pa.nsrc * pa.mul_in
            x = torch.cat([gathered[pa.src0 + s][:, pa.in_off:pa.in_off + pa.mul_in * d1] for s in range(pa.nsrc)], dim=1)
            x = x.reshape(E, K, d1)
            if pa.kind == 0:
                T = torch.zeros(E, d1, d3, dtype=dt)
                ks = op.cg_ks[pa.cg_kstart:pa.cg_kstart + d3 + 1]
                for k in range(d3):
                    for n in range(ks[k], ks[k + 1]):
                        ij = int(op.cg_ij[pa.cg_off + n])
                        T[:, ij & 255, k] += float(op.cg_val[pa.cg_off + n]) * sh[:, pa.sh_off + (ij >> 8)]
                A = torch.einsum("zui,zik->zku", x, T)
                W = wbuf[pa.w_off:pa.w_off + K * ty.mpad].view(K, ty.mpad)
                W3 = wbuf[pa.w3_off:pa.w3_off + op.h2 * ty.mpad].view(op.h2, ty.mpad)
                Lf = wbuf[pa.lf_off:pa.lf_off + ty.mpad * ty.mpad].view(ty.mpad, ty.mpad)
                g = h2[pa.branch] @ W3
                B = (A @ W) * g[:, None, :]
                acc += B @ Lf
            else:
                Lf = wbuf[pa.lf_off:pa.lf_off + K * ty.mpad].view(K, ty.mpad)
                acc += x.transpose(1, 2) @ Lf
        msg[:, ty.out_off:ty.out_off + ty.mul * d3] = acc[:, :, :ty.mul].transpose(1, 2).reshape(E, ty.mul * d3)
    if out_rows is None:
        return msg
    out = torch.zeros(n_out, D, dtype=dt)
    return out.index_add_(0, out_rows, msg)


def emulate_linear(op: LinearOp, weight: torch.Tensor, x: torch.Tensor):
    y = torch.zeros(x.shape[0], op.irreps_out.dim, dtype=x.dtype)
    w = weight * torch.from_numpy(op._scale_np).to(x.dtype)
    for b in op.blocks:
        xi = x[:, b.in_off:b.in_off + b.mul_in * b.dim].reshape(-1, b.mul_in, b.dim)
        W = w[b.w_off:b.w_off + b.mul_in * b.mul_out].view(b.mul_in, b.mul_out)
        y[:, b.out_off:b.out_off + b.mul_out * b.dim] += torch.einsum("zui,uw->zwi", xi, W).reshape(x.shape[0], -1)
    return y


def emulate_gate(g: GateLayout, h: torch.Tensor):
    d = g.desc
    out = torch.zeros(h.shape[0], d.out_dim, dtype=h.dtype)
    ssp = lambda v: torch.nn.functional.softplus(v) - np.log(2.0)
    for s in range(d.n_scalar_slots):
        v = h[:, d.sc_in_off[s]:d.sc_in_off[s] + d.sc_n[s]]
        out[:, d.sc_out_off[s]:d.sc_out_off[s] + d.sc_n[s]] = ssp(v) * d.c_ssp if d.sc_act[s] == 0 else torch.tanh(v) * d.c_tanh
    for s in range(d.n_gated):
        mul, dim = d.gd_mul[s], d.gd_dim[s]
        gate = ssp(h[:, d.gd_gate_off[s]:d.gd_gate_off[s] + mul]) * d.c_ssp
        v = h[:, d.gd_in_off[s]:d.gd_in_off[s] + mul * dim].reshape(-1, mul, dim)
        out[:, d.gd_out_off[s]:d.gd_out_off[s] + mul * dim] = (v * gate[:, :, None]).reshape(h.shape[0], -1)
    return out


def emulate_resblock(rb, x, extra=None, post=None, post_w=None):
    h = emulate_linear(rb.op1, rb.linear1.weight.detach().to(x.dtype), x)
    a = emulate_gate(rb.gate, h)
    y = x + emulate_linear(rb.op2, rb.linear2.weight.detach().to(x.dtype), a)
    if extra is not None:
        y = y + extra
    if post is not None:
        y = emulate_linear(post, post_w.detach().to(x.dtype), y)
    return y


def emulate_ham(asm: HamAssembly, coef, partner, h0, z, na, nb, symmetrize=True):
    nn2 = asm.nao ** 2
    raw = torch.zeros(coef.shape[0], nn2, dtype=coef.dtype)
    val = torch.from_numpy(asm.val).to(coef.dtype)
    for q in range(nn2):
        n0, n1 = asm.row_ptr[q], asm.row_ptr[q + 1]
        if n1 > n0:
            raw[:, q] = (coef[:, torch.from_numpy(asm.col[n0:n1]).long()] * val[n0:n1]).sum(-1)
    m = raw.view(-1, asm.nao, asm.nao)
    if symmetrize:
        other = m if partner is None else m[partner]
        m = 0.5 * (m + other.transpose(1, 2))
    if h0 is not None:
        m = m + h0.view(-1, asm.nao, asm.nao)
    mask = torch.from_numpy(asm.mask).to(coef.dtype)
    za = z if na is None else z[na]
    zb = z if nb is None else z[nb]
    return (m * mask[za][:, :, None] * mask[zb][:, None, :]).reshape(-1, nn2)


def _decode_image(wbuf, off, N, kc):
    """hi|lo operand image (UMMA interleaved K-major, N rows, kc K-columns) -> dense [kc, N] fp32 (hi + lo)."""
    k = torch.arange(kc)[:, None]
    n = torch.arange(N)[None, :]
    idx = (k // 4) * (N * 4) + n * 4 + (k % 4)
    hi = wbuf[off + idx]
    lo = wbuf[off + N * kc + idx]
    # hi must be representable in tf32 (13 low mantissa bits clear)
    assert int((hi.float().view(torch.int32) & 8191).abs().max()) == 0
    return hi + lo


def emulate_msgpack_tc(op: MessagePackOp, wbuf: torch.Tensor, sources, rows, sh, rbf, out_rows=None, n_out=None):
    """Mirrors msgpack_tc_kernel's arithmetic from the tensor-core packing (tc_types_c / tc_paths_c / tc wbuf)."""
    from hamgnn_b200 import so3
    E = sh.shape[0]
    dt = wbuf.dtype
    D = op.irreps_out.dim
    msg = torch.zeros(E, D, dtype=dt)
    act = so3.normalize2mom_const("silu")
    h2 = []
    for b in range(len(op.branches)):
        w1 = wbuf[op.tc_fc1_off[b]:op.tc_fc1_off[b] + op.rbf_dim * op.h1].view(op.rbf_dim, op.h1)
        w2 = wbuf[op.tc_fc2_off[b]:op.tc_fc2_off[b] + op.h1 * op.h2].view(op.h1, op.h2)
        h2.append(_silu(_silu(rbf @ w1) * act @ w2) * act)
    gathered = [s if r is None else s[r] for s, r in zip(sources, rows)]
    for t in range(len(op.irreps_out)):
        ty = op.tc_types_c[t]
        assert ty.mpad % 16 == 0
        d3 = 2 * ty.l + 1
        acc = torch.zeros(E, d3, ty.mpad, dtype=dt)
        for p in range(ty.path_begin, ty.path_end):
            pa = op.tc_paths_c[p]
            d1 = 2 * pa.l1 + 1
            K = pa.nsrc * pa.mul_in
            Kpad = (K + 7) // 8 * 8
            x = torch.cat([gathered[pa.src0 + s][:, pa.in_off:pa.in_off + pa.mul_in * d1] for s in range(pa.nsrc)], dim=1)
            x = torch.nn.functional.pad(x.reshape(E, K, d1), (0, 0, 0, Kpad - K))
            woff = pa.w_off if pa.kind == 0 else pa.lf_off
            W = torch.cat([_decode_image(wbuf, woff + 2 * ty.mpad * 32 * c, ty.mpad, min(32, Kpad - u0))
                           for c, u0 in enumerate(range(0, Kpad, 32))], dim=0)
            if pa.kind == 0:
                T = torch.zeros(E, d1, d3, dtype=dt)
                ks = op.cg_ks[pa.cg_kstart:pa.cg_kstart + d3 + 1]
                for k in range(d3):
                    for n in range(ks[k], ks[k + 1]):
                        ij = int(op.cg_ij[pa.cg_off + n])
                        T[:, ij & 255, k] += float(op.cg_val[pa.cg_off + n]) * sh[:, pa.sh_off + (ij >> 8)]
                A = torch.einsum("zui,zik->zku", x, T)
                W3 = _decode_image(wbuf, pa.w3_off, ty.mpad, op.h2)
                Lf = _decode_image(wbuf, pa.lf_off, ty.mpad, ty.mpad)      # [k = w, n = w']
                g = h2[pa.branch] @ W3
                acc += ((A @ W) * g[:, None, :]) @ Lf
            else:
                acc += x.transpose(1, 2) @ W
        msg[:, ty.out_off:ty.out_off + ty.mul * d3] = acc[:, :, :ty.mul].transpose(1, 2).reshape(E, ty.mul * d3)
    if out_rows is None:
        return msg
    return torch.zeros(n_out, D, dtype=dt).index_add_(0, out_rows, msg)


# ------------------------------------------------------------------------------------------------ SOC (a16)
def emulate_sorted_head(head, weight, x):
    """SortedHeadOp.forward: one emulate_linear per column chunk on the gathered weights."""
    y = torch.zeros(x.shape[0], head.out_dim, dtype=x.dtype)
    flat = weight.detach().reshape(-1).to(x.dtype)
    for op, col0, gather in head.chunks:
        y[:, col0:col0 + op.irreps_out.dim] = emulate_linear(op, flat[torch.from_numpy(gather)], x)
    return y


def emulate_csr_rows(row_ptr, col, val, n_out, x):
    y = torch.zeros(x.shape[0], n_out, dtype=x.dtype)
    v = torch.from_numpy(val).to(x.dtype)
    rows = np.repeat(np.arange(n_out), np.diff(row_ptr))
    y.index_add_(1, torch.from_numpy(rows).long(), x[:, torch.from_numpy(col).long()] * v)
    return y


def emulate_finalize_su2(nao, mask_tab, raw, partner, h0_re, h0_im, z, na, nb, symmetrize=True):
    """ham_finalize_su2_kernel: raw = [rows][2][M][M]; returns (real rows, imaginary rows)."""
    M = 2 * nao
    re, im = raw[:, :M * M].reshape(-1, M, M), raw[:, M * M:].reshape(-1, M, M)
    if symmetrize:
        o_re, o_im = (re, im) if partner is None else (re[partner], im[partner])
        re, im = 0.5 * (re + o_re.transpose(1, 2)), 0.5 * (im - o_im.transpose(1, 2))
    mask = torch.from_numpy(mask_tab).to(raw.dtype)
    za = z if na is None else z[na]
    zb = z if nb is None else z[nb]
    m2 = (mask[za].repeat(1, 2)[:, :, None] * mask[zb].repeat(1, 2)[:, None, :])
    re, im = (re * m2).reshape(-1, M * M), (im * m2).reshape(-1, M * M)
    if h0_re is not None:
        re = re + h0_re
    if h0_im is not None:
        im = im + h0_im
    return re, im


def emulate_ksi_shell_average(nao, shells, ksi):
    m = ksi.reshape(-1, nao, nao).clone()
    for a, b in shells:
        m[:, a:b] = m[:, a:b].mean(dim=1, keepdim=True).expand(-1, b - a, -1)
    for a, b in shells:
        m[:, :, a:b] = m[:, :, a:b].mean(dim=2, keepdim=True).expand(-1, -1, b - a)
    return m.reshape(ksi.shape[0], -1)


def emulate_finalize_so3(nao, hns, ksi, lmat, partner, h0_re, h0_im, symmetrize=True, h0_offdiag_only=False):
    M = 2 * nao
    A = []
    for c in range(3):
        a = (ksi * lmat[:, :, c]).reshape(-1, nao, nao)
        if symmetrize:
            o = a if partner is None else a[partner]
            a = 0.5 * (a - o.transpose(1, 2))
        A.append(a)
    h = hns.reshape(-1, nao, nao)
    re = torch.zeros(h.shape[0], M, M, dtype=hns.dtype)
    im = torch.zeros_like(re)
    re[:, :nao, :nao], re[:, nao:, nao:], re[:, :nao, nao:], re[:, nao:, :nao] = h, h, A[1], A[1]
    im[:, :nao, :nao], im[:, nao:, nao:], im[:, :nao, nao:], im[:, nao:, :nao] = A[2], -A[2], A[0], -A[0]
    if h0_re is not None:
        h0 = h0_re.reshape(-1, M, M).clone()
        if h0_offdiag_only:
            h0[:, :nao, :nao] = 0
            h0[:, nao:, nao:] = 0
        re = re + h0
    if h0_im is not None:
        im = im + h0_im.reshape(-1, M, M)
    return re.reshape(-1, M * M), im.reshape(-1, M * M)


def emulate_radial_gate_tc(op: MessagePackOp, wbuf: torch.Tensor, rbf: torch.Tensor):
    """radial_gate_tc_kernel from the tensor-core packing: layers 1-2 from the plain copies, the last layer from the
    (hi | lo) tiles of GATE_TILE_COLS gate columns.  Returns [n_branches, E, max n_channels]."""
    from hamgnn_b200 import so3
    act = so3.normalize2mom_const("silu")
    out = torch.zeros(len(op.branches), rbf.shape[0], max(op.n_channels), dtype=wbuf.dtype)
    for b in range(len(op.branches)):
        w1 = wbuf[op.tc_fc1_off[b]:op.tc_fc1_off[b] + op.rbf_dim * op.h1].view(op.rbf_dim, op.h1)
        w2 = wbuf[op.tc_fc2_off[b]:op.tc_fc2_off[b] + op.h1 * op.h2].view(op.h1, op.h2)
        h2 = _silu(_silu(rbf @ w1) * act @ w2) * act
        assert op.tc_w3img_off[b] % 4 == 0
        TN = op.GATE_TILE_COLS
        for t, n0 in enumerate(range(0, op.n_channels[b], TN)):
            W = _decode_image(wbuf, op.tc_w3img_off[b] + t * 2 * op.h2 * TN, TN, op.h2)      # [h2, TN]
            n1 = min(n0 + TN, op.n_channels[b])
            out[b, :, n0:n1] = (h2 @ W)[:, :n1 - n0]
            assert float(W[:, n1 - n0:].abs().max()) == 0 if n1 - n0 < TN else True            # zero padding
    return out


# ------------------------------------------------------------------------------------------------ rotated frame
def emulate_wigner(vec: np.ndarray, op: MessagePackOp) -> np.ndarray:
    """wigner_kernel: per-edge D^l(R), R = R_y(-theta) R_z(-phi) taking the unit edge vector to the polar axis z,
    built as J Z(-theta) J^T Z(-phi) from the vector components (no angles), fp64 -> [E, dstride]."""
    v = np.asarray(vec, dtype=np.float64)
    E = v.shape[0]
    n = np.sqrt((v * v).sum(1))
    x, y, z = v[:, 0] / n, v[:, 1] / n, v[:, 2] / n
    rho = np.sqrt(x * x + y * y)
    ok = rho > 1e-30
    cp, sp = np.where(ok, x / np.where(ok, rho, 1.0), 1.0), np.where(ok, y / np.where(ok, rho, 1.0), 0.0)
    ct, st = z, rho
    lmax = op.rot_lmax
    # cos / sin of m * (-phi) and m * (-theta) by the angle-addition recurrence
    ca, sa, cb, sb = [np.ones(E)], [np.zeros(E)], [np.ones(E)], [np.zeros(E)]
    for m in range(1, lmax + 1):
        ca.append(ca[-1] * cp - sa[-1] * (-sp)); sa.append(sa[-1] * cp + ca[-2] * (-sp))
        cb.append(cb[-1] * ct - sb[-1] * (-st)); sb.append(sb[-1] * ct + cb[-2] * (-st))
    out = np.zeros((E, op.rot_dstride))
    for l in range(lmax + 1):
        d = 2 * l + 1
        J = op.rot_wigner_j[op.rot_doff[l]:op.rot_doff[l] + d * d].reshape(d, d)
        Jt = J.T
        N = np.zeros((E, d, d))
        N[:, :, l] = Jt[:, l]
        for m in range(1, l + 1):   # N = J^T Z(-phi): column mixing
            N[:, :, l + m] = Jt[None, :, l + m] * ca[m][:, None] + Jt[None, :, l - m] * sa[m][:, None]
            N[:, :, l - m] = -Jt[None, :, l + m] * sa[m][:, None] + Jt[None, :, l - m] * ca[m][:, None]
        M2 = np.zeros((E, d, d))
        M2[:, l, :] = N[:, l, :]
        for m in range(1, l + 1):   # M2 = Z(-theta) N: row mixing
            M2[:, l + m, :] = cb[m][:, None] * N[:, l + m, :] - sb[m][:, None] * N[:, l - m, :]
            M2[:, l - m, :] = sb[m][:, None] * N[:, l + m, :] + cb[m][:, None] * N[:, l - m, :]
        Dl = np.einsum("ri,eij->erj", J, M2)
        out[:, op.rot_doff[l]:op.rot_doff[l] + d * d] = Dl.reshape(E, d * d)
    return out


def emulate_rotate_pack(op: MessagePackOp, sources, rows, Dw: torch.Tensor) -> torch.Tensor:
    """rotate_pack_kernel: XP[tile][block xoff + m1 * 2 kpad T + chunk c * 2 KC T + (hi: (u%KC)/4 slab, z, u%4 | lo)].
    Returns [n_tiles, tile_stride] holding the un-split rotated values in the hi image and zeros in the lo image
    (the emulation keeps full precision; the device writes hi = tf32(x'), lo = x' - hi)."""
    T, KC = op.ROT_TILE, op.ROT_KC
    E = Dw.shape[0]
    nt = (E + T - 1) // T
    XP = torch.zeros(nt, op.rot_tile_stride, dtype=Dw.dtype)
    gathered = [s if r is None else s[r] for s, r in zip(sources, rows)]
    zz = torch.arange(E)
    tile, zl = zz // T, zz % T
    for bi in range(op.rot_n_blocks):
        b = op.rot_blocks_c[bi]
        d1 = 2 * b.l1 + 1
        K = b.nsrc * b.mul
        x = torch.cat([gathered[b.src0 + s][:, b.in_off:b.in_off + b.mul * d1] for s in range(b.nsrc)], dim=1).reshape(E, K, d1)
        Dl = Dw[:, op.rot_doff[b.l1]:op.rot_doff[b.l1] + d1 * d1].reshape(E, d1, d1)
        xr = torch.einsum("zmi,zui->zum", Dl, x)      # x'[u][m] = sum_i D[m][i] x[u][i]
        for m in range(d1):
            base = b.xoff + m * 2 * b.kpad * T
            for u in range(K):
                c, ul = u // KC, u % KC
                off = base + c * 2 * KC * T + (ul // 4) * (T * 4) + zl * 4 + (ul % 4)
                XP[tile, off] = xr[:, u, m]
    return XP


def _decode_a(XP, a_off, kpad, E, T, KC):
    """[E, kpad] operand (hi + lo) of one step from the packed rotated input."""
    zz = torch.arange(E)
    tile, zl = zz // T, zz % T
    cols = []
    for u in range(kpad):
        c, ul = u // KC, u % KC
        kc = min(KC, kpad - c * KC)
        off = a_off + c * 2 * KC * T + (ul // 4) * (T * 4) + zl * 4 + (ul % 4)
        cols.append(XP[tile, off] + XP[tile, off + kc * T])
    return torch.stack(cols, dim=1)


def emulate_msgpack_rot(op: MessagePackOp, wbuf: torch.Tensor, sources, rows, vec, rbf, out_rows=None, n_out=None):
    """Mirrors the 'rot' pipeline: wigner_kernel -> rotate_pack_kernel -> radial gate -> msgpack_rot_kernel
    (steps from rot_steps_c, W / L' images from the tensor-core packing, rotation back in the epilogue)."""
    from hamgnn_b200 import so3
    T, KC = op.ROT_TILE, op.ROT_KC
    E = rbf.shape[0]
    dt = wbuf.dtype
    D = op.irreps_out.dim
    Dw = torch.from_numpy(emulate_wigner(np.asarray(vec), op)).to(dt)
    XP = emulate_rotate_pack(op, sources, rows, Dw)
    act = so3.normalize2mom_const("silu")
    g = []
    for b in range(len(op.branches)):
        w1 = wbuf[op.tc_fc1_off[b]:op.tc_fc1_off[b] + op.rbf_dim * op.h1].view(op.rbf_dim, op.h1)
        w2 = wbuf[op.tc_fc2_off[b]:op.tc_fc2_off[b] + op.h1 * op.h2].view(op.h1, op.h2)
        w3 = wbuf[op.tc_w3_off[b]:op.tc_w3_off[b] + op.h2 * op.n_channels[b]].view(op.h2, op.n_channels[b])
        g.append(_silu(_silu(rbf @ w1) * act @ w2) * act @ w3)
    msg = torch.zeros(E, D, dtype=dt)
    for t in range(len(op.irreps_out)):
        ty = op.tc_types_c[t]
        d3, mp = 2 * ty.l + 1, ty.mpad
        Cacc = torch.zeros(E, d3, mp, dtype=dt)
        Lf = None
        for si in range(op.rot_step_begin[t], op.rot_step_begin[t + 1]):
            st = op.rot_steps_c[si]
            A = _decode_a(XP, st.a_off, st.kpad, E, T, KC)
            W = torch.cat([_decode_image(wbuf, st.w_off + 2 * mp * KC * c, mp, min(KC, st.kpad - u0))
                           for c, u0 in enumerate(range(0, st.kpad, KC))], dim=0)
            if st.kind == 0:
                if st.new_path & 1:
                    Lf = _decode_image(wbuf, st.lf_off, mp, mp)
                gv = torch.zeros(E, mp, dtype=dt)
                gv[:, :ty.mul] = g[st.branch][:, st.g_off:st.g_off + ty.mul] if st.branch >= 0 else 1.0
                Cacc[:, st.m3, :] += ((A @ W) * (gv * st.scale)) @ Lf
            else:
                Cacc[:, st.m3, :] += A @ W
        D3 = Dw[:, op.rot_doff[ty.l]:op.rot_doff[ty.l] + d3 * d3].reshape(E, d3, d3)
        out = torch.einsum("zmk,zmw->zwk", D3, Cacc[:, :, :ty.mul])     # C = D^T C'
        msg[:, ty.out_off:ty.out_off + ty.mul * d3] = out.reshape(E, ty.mul * d3)
    if out_rows is None:
        return msg
    return torch.zeros(n_out, D, dtype=dt).index_add_(0, out_rows, msg)


# ------------------------------------------------------------------------------------------------ rotated frame, A-stationary
def emulate_unrotate(op: MessagePackOp, CP: torch.Tensor, Dw: torch.Tensor, seg_ptr=None, seg_order=None):
    """unrotate_kernel: message[z][t][w][k] = sum_m3 D^{l3}_z[m3][k] C'[z][ccol[t][m3] + w]; with segments the rows of a
    segment are summed in list order (the deterministic receiver reduction)."""
    E = CP.shape[0]
    msg = torch.zeros(E, op.irreps_out.dim, dtype=CP.dtype)
    for t in range(len(op.irreps_out)):
        ty = op.tc_types_c[t]
        d3, M = 2 * ty.l + 1, ty.mul
        Cp = torch.zeros(E, d3, M, dtype=CP.dtype)
        for m in range(d3):
            c = int(op.rot2_ccol[t, m])
            if c >= 0:
                Cp[:, m, :] = CP[:, c:c + M]
        D3 = Dw[:, op.rot_doff[ty.l]:op.rot_doff[ty.l] + d3 * d3].reshape(E, d3, d3)
        msg[:, ty.out_off:ty.out_off + M * d3] = torch.einsum("zmk,zmw->zwk", D3, Cp).reshape(E, M * d3)
    if seg_ptr is None:
        return msg
    out = torch.zeros(len(seg_ptr) - 1, op.irreps_out.dim, dtype=CP.dtype)
    for i in range(len(seg_ptr) - 1):
        for j in range(int(seg_ptr[i]), int(seg_ptr[i + 1])):
            out[i] += msg[int(seg_order[j])]
    return out


def emulate_msgpack_rot2(op: MessagePackOp, wbuf: torch.Tensor, sources, rows, vec, rbf, seg_ptr=None, seg_order=None):
    """Mirrors the 'rot2' pipeline: wigner -> rotate_pack -> radial gate -> msgpack_rot2_kernel (passes / pieces, the two
    gate streams with tensor / SIMT / dummy entries, tensor destination groups) -> unrotate_kernel."""
    from hamgnn_b200 import so3
    T, KC, KC2 = op.ROT_TILE, op.ROT_KC, op.R2_KC
    E = rbf.shape[0]
    dt = wbuf.dtype
    Dw = torch.from_numpy(emulate_wigner(np.asarray(vec), op)).to(dt)
    XP = emulate_rotate_pack(op, sources, rows, Dw)
    act = so3.normalize2mom_const("silu")
    g = []
    for b in range(len(op.branches)):
        w1 = wbuf[op.tc_fc1_off[b]:op.tc_fc1_off[b] + op.rbf_dim * op.h1].view(op.rbf_dim, op.h1)
        w2 = wbuf[op.tc_fc2_off[b]:op.tc_fc2_off[b] + op.h1 * op.h2].view(op.h1, op.h2)
        w3 = wbuf[op.tc_w3_off[b]:op.tc_w3_off[b] + op.h2 * op.n_channels[b]].view(op.h2, op.n_channels[b])
        g.append(_silu(_silu(rbf @ w1) * act @ w2) * act @ w3)
    n_pass, n_piece, n_batch, n_dst = op.rot2_n
    CP = torch.zeros(E, op.rot2_rowstride, dtype=dt)
    seen = torch.zeros(op.rot2_rowstride, dtype=torch.bool)
    for pi in range(n_pass):
        ps = op.rot2_passes_c[pi]
        assert ps.ncols <= op.R2_ACC
        acc = torch.zeros(E, ps.ncols, dtype=dt)
        NH = op.R2_NH
        cur = [ps.stream_begin[h] for h in range(NH)]
        end = [ps.stream_end[h] for h in range(NH)]
        s_run = [None] * NH
        for qi in range(ps.piece_begin, ps.piece_end):
            pc = op.rot2_pieces_c[qi]
            assert pc.ncols % 16 == 0 and pc.ncols <= op.R2_NB and pc.kpad % 8 == 0
            assert pc.w_off % 4 == 0 and pc.l_off % 4 == 0 and pc.l_floats % 4 == 0 and pc.l_floats <= op.R2_LMAX_FLOATS
            A = _decode_a(XP, pc.a_off, pc.kpad, E, T, KC)
            W = torch.cat([_decode_image(wbuf, pc.w_off + 2 * pc.ncols * KC2 * c, pc.ncols, min(KC2, pc.kpad - u0))
                           for c, u0 in enumerate(range(0, pc.kpad, KC2))], dim=0)
            B = A @ W
            Bg = torch.zeros_like(B)
            covered = torch.zeros(pc.ncols // 8, dtype=torch.bool)
            gst = op.rot2_gstride

            def gate4(off):
                if off == 0xFFFFFFFF:
                    return torch.ones(E, 4, dtype=dt)
                assert off % T == 0
                br, col = divmod(off // T, gst)
                assert br < len(g) and col + 4 <= gst
                out = torch.zeros(E, 4, dtype=dt)           # columns beyond the branch's width: zero-initialised workspace
                nvv = max(0, min(4, g[br].shape[1] - col))
                out[:, :nvv] = g[br][:, col:col + nvv]
                return out

            for h in range(NH):
                first = True
                while True:
                    assert cur[h] < end[h]
                    bt = op.rot2_batches_c[cur[h]]
                    cur[h] += 1
                    meta = bt.meta & 0xFFFFFFFF
                    kind = meta & 3
                    assert bool((meta >> 2) & 1) == first
                    first = False
                    if kind != 2:
                        c8 = (meta >> 8) & 0xFF
                        assert not covered[c8]
                        covered[c8] = True
                        gv = torch.cat([gate4(bt.goff_a), gate4(bt.goff_b)], dim=1)
                        bg = B[:, 8 * c8:8 * c8 + 8] * gv
                        if kind == 0:
                            Bg[:, 8 * c8:8 * c8 + 8] = bg
                        else:
                            M, acc0 = (meta >> 16) & 0x1F, (meta >> 21) & 0xFF
                            m4 = (M + 3) // 4 * 4
                            assert M <= op.R2_SIMT_MAX and bt.l_off % 4 == 0 and bt.l_off + 8 * m4 <= pc.l_floats
                            Lr = wbuf[pc.l_off + bt.l_off:pc.l_off + bt.l_off + 8 * m4].view(8, m4)
                            if (meta >> 4) & 1:
                                s_run[h] = torch.zeros(E, m4, dtype=dt)
                            s_run[h] = s_run[h] + bg @ Lr
                            if (meta >> 5) & 1:
                                assert float(s_run[h][:, M:].abs().max()) == 0 if M < m4 else True
                                acc[:, acc0:acc0 + M] += s_run[h][:, :M]
                                s_run[h] = None
                    if (meta >> 3) & 1:
                        break
            # every non-padding 8-column batch of the piece is gated exactly once
            s_used = 0
            for di in range(pc.dst_begin, pc.dst_begin + pc.ndst):
                ds = op.rot2_dsts_c[di]
                assert ds.col0 % 8 == 0 and ds.kcols % 8 == 0 and ds.col0 + ds.kcols <= pc.ncols and ds.mp % 16 == 0
                assert ds.s_off == s_used and ds.s_off + ds.mp <= op.R2_SW and ds.acc_col0 + ds.mul <= ps.ncols
                assert ds.l_rel + 2 * ds.kcols * ds.mp <= pc.l_floats
                assert bool(covered[ds.col0 // 8:(ds.col0 + ds.kcols) // 8].all())
                s_used += ds.mp
                Lst = _decode_image(wbuf, pc.l_off + ds.l_rel, ds.mp, ds.kcols)          # [kcols, mp]
                S = Bg[:, ds.col0:ds.col0 + ds.kcols] @ Lst
                acc[:, ds.acc_col0:ds.acc_col0 + ds.mul] += S[:, :ds.mul]
                assert float(S[:, ds.mul:].abs().max()) == 0 if ds.mul < ds.mp else True
        assert cur == end and s_run == [None] * NH
        assert not seen[ps.out_col0:ps.out_col0 + ps.ncols].any()
        seen[ps.out_col0:ps.out_col0 + ps.ncols] = True
        CP[:, ps.out_col0:ps.out_col0 + ps.ncols] = acc
    assert seen.all()
    return emulate_unrotate(op, CP, Dw, seg_ptr, seg_order)


# ------------------------------------------------------------------------------------------------ rotated frame, fp16 x 2 split
def _pow2_scale(amax: torch.Tensor, shift: int):
    """pow2_scale of msgpack_rot16_kernel.cuh: s = 2^(141 - E - shift), E = biased fp32 exponent of amax (s = 1 for E = 0)."""
    E = (amax.float().view(torch.int32) >> 23) & 0xff
    sb = torch.where(E == 0, torch.full_like(E, 127), (127 + 141 - E - shift).clamp(1, 253))
    return torch.ldexp(torch.ones_like(amax, dtype=torch.float64), (sb - 127)), torch.ldexp(torch.ones_like(amax, dtype=torch.float64), (127 - sb))


def _split16(v: torch.Tensor):
    hi = v.float().to(torch.float16)
    lo = (v.float() - hi.float()).to(torch.float16)
    return hi.double(), lo.double()


def _mm3(ah, al, bh, bl):
    """lo.hi + hi.lo + hi.hi, what the three kind::f16 MMAs accumulate (the products are exact in fp32)."""
    return al @ bh + ah @ bl + ah @ bh


def _decode_image16(buf16: torch.Tensor, word0: int, N: int, kw: int):
    """(hi, lo) dense [2 kw channels, N] from a packed fp16 image of kw word-columns starting at 32-bit word `word0`."""
    k = torch.arange(2 * kw)[:, None]
    n = torch.arange(N)[None, :]
    kwd = k // 2
    h = 2 * (word0 + (kwd // 4) * (N * 4) + n * 4 + (kwd % 4)) + (k % 2)
    return buf16[h].double(), buf16[h + 2 * N * kw].double()


def emulate_rotate_pack16(op: MessagePackOp, sources, rows, Dw: torch.Tensor):
    """rotate_pack16_kernel: per (edge, block) a power-of-two scale from the row maximum of the UNROTATED block (bound
    sqrt(d1) < 4 on the rotation), then x' s split into fp16 hi / lo, two channels per word.  Returns the halfword buffer
    [n_tiles, 2 tile_stride] and sx [n_tiles, n_blocks, 128]."""
    T, KC = op.ROT_TILE, op.ROT_KC
    E = Dw.shape[0]
    nt = (E + T - 1) // T
    XP = torch.zeros(nt, 2 * op.rot16_tile_stride, dtype=torch.float16)
    SX = torch.ones(nt, op.rot16_n_blocks, T, dtype=torch.float64)
    gathered = [s if r is None else s[r] for s, r in zip(sources, rows)]
    zz = torch.arange(E)
    tile, zl = zz // T, zz % T
    for bi in range(op.rot16_n_blocks):
        b = op.rot16_blocks_c[bi]
        d1 = 2 * b.l1 + 1
        K = b.nsrc * b.mul
        kw = b.kpad // 2
        x = torch.cat([gathered[b.src0 + s][:, b.in_off:b.in_off + b.mul * d1] for s in range(b.nsrc)], dim=1).reshape(E, K, d1)
        s, inv = _pow2_scale(x.abs().amax(dim=(1, 2)), 0 if b.l1 == 0 else 2)
        SX[tile, bi, zl] = inv
        Dl = Dw[:, op.rot_doff[b.l1]:op.rot_doff[b.l1] + d1 * d1].reshape(E, d1, d1)
        xr = torch.einsum("zmi,zui->zum", Dl, x * s[:, None, None])
        assert float(xr.abs().max()) < 32768.0
        hi, lo = _split16(xr)
        for m in range(d1):
            base = b.xoff + m * 2 * kw * T
            for u in range(K):
                w = u // 2
                c, wl = w // KC, w % KC
                kc = min(KC, kw - c * KC)
                word = base + c * 2 * KC * T + (wl // 4) * (T * 4) + zl * 4 + (wl % 4)
                XP[tile, 2 * word + (u % 2)] = hi[:, u, m].to(torch.float16)
                XP[tile, 2 * (word + kc * T) + (u % 2)] = lo[:, u, m].to(torch.float16)
    return XP, SX


def _decode_a16(XP, a_off, kw, E, T, KC):
    zz = torch.arange(E)
    tile, zl = zz // T, zz % T
    hs, ls = [], []
    for u in range(2 * kw):
        w = u // 2
        c, wl = w // KC, w % KC
        kc = min(KC, kw - c * KC)
        word = a_off + c * 2 * KC * T + (wl // 4) * (T * 4) + zl * 4 + (wl % 4)
        hs.append(XP[tile, 2 * word + (u % 2)].double())
        ls.append(XP[tile, 2 * (word + kc * T) + (u % 2)].double())
    return torch.stack(hs, dim=1), torch.stack(ls, dim=1)


def emulate_msgpack_rot16(op: MessagePackOp, st: dict, sources, rows, vec, rbf, out_rows=None, n_out=None):
    """Mirrors wigner -> rotate_pack16 -> radial gate -> msgpack_rot16_kernel from the rot16 tables, the packed fp16 images
    (st['r16_wbuf'], st['r16_inv'] from MessagePackOp.pack_rot16) and the kernel's scale bookkeeping; fp16 quantisation
    and the dropped lo.lo products are emulated, sums run in fp64."""
    from hamgnn_b200 import so3
    T, KC = op.ROT_TILE, op.ROT_KC
    E = rbf.shape[0]
    dt = torch.float64
    wbuf = st["tc_wbuf"].double().cpu()
    buf16 = st["r16_wbuf"].cpu()
    inv_img = st["r16_inv"].double().cpu()
    D = op.irreps_out.dim
    Dw = torch.from_numpy(emulate_wigner(np.asarray(vec), op)).to(dt)
    XP, SX = emulate_rotate_pack16(op, sources, rows, Dw)
    zz = torch.arange(E)
    tile, zl = zz // T, zz % T
    act = so3.normalize2mom_const("silu")
    g = []
    for b in range(len(op.branches)):
        w1 = wbuf[op.tc_fc1_off[b]:op.tc_fc1_off[b] + op.rbf_dim * op.h1].view(op.rbf_dim, op.h1)
        w2 = wbuf[op.tc_fc2_off[b]:op.tc_fc2_off[b] + op.h1 * op.h2].view(op.h1, op.h2)
        w3 = wbuf[op.tc_w3_off[b]:op.tc_w3_off[b] + op.h2 * op.n_channels[b]].view(op.h2, op.n_channels[b])
        g.append(_silu(_silu(rbf @ w1) * act @ w2) * act @ w3)
    msg = torch.zeros(E, D, dtype=dt)
    for t in range(len(op.irreps_out)):
        ty = op.tc_types_c[t]
        d3, mp = 2 * ty.l + 1, ty.mpad
        Cacc = torch.zeros(E, d3, mp, dtype=dt)
        for si in range(op.rot16_step_begin[t], op.rot16_step_begin[t + 1]):
            s_ = op.rot16_steps_c[si]
            kw = s_.kpad
            wimg, limg = s_.pad2 & 0xffff, (s_.pad2 >> 16) & 0xffff
            ah, al = _decode_a16(XP, s_.a_off, kw, E, T, KC)
            wh, wl = [], []
            for c, w0 in enumerate(range(0, kw, KC)):
                h_, l_ = _decode_image16(buf16, s_.w_off + 2 * mp * KC * c, mp, min(KC, kw - w0))
                wh.append(h_); wl.append(l_)
            B = _mm3(ah, al, torch.cat(wh, 0), torch.cat(wl, 0))
            sc = s_.scale * inv_img[wimg] * SX[tile, s_.pad, zl]
            gv = torch.zeros(E, mp, dtype=dt)
            if s_.branch >= 0:
                gv[:, :ty.mul] = g[s_.branch][:, s_.g_off:s_.g_off + ty.mul] * sc[:, None]
            else:
                gv[:, :ty.mul] = sc[:, None]
            P = B * gv
            if mp == 16:   # msgpack_rotf_kernel<F16>: L' on the fp32 FMA pipes from the un-split image at lf_off of the tensor-core wbuf
                Cacc[:, s_.m3, :] += P @ wbuf[s_.lf_off:s_.lf_off + mp * mp].view(mp, mp)
                continue
            ps, pinv = _pow2_scale(P.abs().amax(dim=1), 0)
            ph, pl = _split16(P * ps[:, None])
            lh, ll = _decode_image16(buf16, s_.lf_off, mp, mp // 2)
            S = pl @ lh + ph @ ll + ph @ lh      # GEMM2: (gl.bh) + (bq.bl) + (bq.bh)
            Cacc[:, s_.m3, :] += S * (pinv * inv_img[limg])[:, None]
        D3 = Dw[:, op.rot_doff[ty.l]:op.rot_doff[ty.l] + d3 * d3].reshape(E, d3, d3)
        out = torch.einsum("zmk,zmw->zwk", D3, Cacc[:, :, :ty.mul])
        msg[:, ty.out_off:ty.out_off + ty.mul * d3] = out.reshape(E, ty.mul * d3)
    if out_rows is None:
        return msg
    return torch.zeros(n_out, D, dtype=dt).index_add_(0, out_rows, msg)


# ------------------------------------------------------------------------------------------------ band-energy head
def emulate_band_kspace(hon, hoff, son, soff, nao, seg_ptr, seg_edge, src, dst, shift, kvec, orb_index, n_orb):
    """band_onsite_kernel + band_offsite_kernel (csrc/band.cu) from the very arguments of hgb_band_kspace: one (i, j) segment at a
    time, the images of the pair added in list order; returns (hk, sk) complex [n_k, n_orb, n_orb]."""
    nk = kvec.shape[0]
    hk = np.zeros((nk, n_orb, n_orb), dtype=np.complex128)
    sk = np.zeros_like(hk)
    oi = np.asarray(orb_index).reshape(-1, nao)
    hon, hoff, son, soff = (np.asarray(t, np.float64) for t in (hon, hoff, son, soff))
    for a in range(oi.shape[0]):
        v = oi[a] >= 0
        r = oi[a][v]
        hk[:, r[:, None], r[None, :]] = hon[a].reshape(nao, nao)[np.ix_(v, v)]
        sk[:, r[:, None], r[None, :]] = son[a].reshape(nao, nao)[np.ix_(v, v)]
    for s in range(len(seg_ptr) - 1):
        e0, e1 = int(seg_ptr[s]), int(seg_ptr[s + 1])
        first = int(seg_edge[e0])
        i, j = int(src[first]), int(dst[first])
        vi, vj = oi[i] >= 0, oi[j] >= 0
        acc_h = np.zeros((nk, int(vi.sum()), int(vj.sum())), dtype=np.complex128)
        acc_s = np.zeros_like(acc_h)
        for q in range(e0, e1):
            e = int(seg_edge[q])
            assert int(src[e]) == i and int(dst[e]) == j, "a segment holds the images of ONE atom pair"
            ph = np.exp(2j * np.pi * (np.asarray(kvec, np.float64) @ np.asarray(shift[e], np.float64)))
            acc_h += ph[:, None, None] * hoff[e].reshape(nao, nao)[np.ix_(vi, vj)]
            acc_s += ph[:, None, None] * soff[e].reshape(nao, nao)[np.ix_(vi, vj)]
        hk[:, oi[i][vi][:, None], oi[j][vj][None, :]] += acc_h
        sk[:, oi[i][vi][:, None], oi[j][vj][None, :]] += acc_s
    return hk, sk
