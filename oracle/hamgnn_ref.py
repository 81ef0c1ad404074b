"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the HamGNN hot path.

PARITY UNPINNED (see oracle/e3lite.py header): the reference cannot be imported in the build
container (e3nn / torch_scatter / PyG / easydict / pymatgen are absent) and ships no tests or golden
vectors; this file restates the reference arithmetic on top of oracle/e3lite.py, module by module,
with the same parameter names so that a reference state_dict maps 1:1:

  FuseSrcDst                 hamgnn/nn/attention_utils.py:85-120   (AttentionHeadsToVector)
  LinearScaleWithWeights     hamgnn/nn/tensor_products.py:25-47
  tp_instructions            hamgnn/nn/message_passing.py:136-171 (== tensor_products.py:116-149)
  MessagePackBlock           hamgnn/nn/message_passing.py:26-231   (non-lite branch :216-231)
  EmbeddingTP                hamgnn/nn/tensor_products.py:51-189
  ResidualBlock              hamgnn/nn/interaction_blocks.py:264-358 + utils/irreps_utils.py:33-65
  ConvBlockE3                hamgnn/nn/convolution.py:23-160
  PairInteractionBlock       hamgnn/nn/interaction_blocks.py:30-164
  PairInteractionEmbedding   hamgnn/nn/embeddings.py:215-337
  edge geometry / SH / RBF   toolbox/nequip/nn/embedding/_edge.py:59-67, nn/embeddings.py:73-100,
                             utils/basis_functions.py:177-208, utils/cutoff_functions.py:35-61
  HamGNNConvE3               hamgnn/models/hamgnn_conv.py:88-284
  HamLayer / HamGNNOut       hamgnn/models/hamgnn_output.py:38-58, 258-272, 345-526, 851-891,
                             1056-1096, 1187-1285, 2288-2365, 2784-2872, 2916-2990, 3771-3799,
                             3966-4021  (non-SOC, non-magnetic branch)
  SU2Decomposition           hamgnn/nn/tensor_decomposition.py:39-86 (irreps_from_l1l2), 439-627
                             (E3TensorDecomposition, spinful=True: __init__ + get_H)
  SOC branches of HamGNNOut  hamgnn/models/hamgnn_output.py:150-153, 188-211, 281-293 (ctor),
                             3026-3144 (so3: xi.L construction), 3146-3178 (su2), 2367-2431
                             (symmetrize_orbital_coefficients), 1231-1285 (is_soc / anti-hermitian
                             symmetrisation), 3603-3625 (+H0, real;imag stacking), 3889-3931 (result dict)
  overlap head               hamgnn/models/hamgnn_output.py:2996-3019, 4006-4014

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import it.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
from torch import nn

from .e3lite import (FullyConnectedNet, Gate, Irrep, Irreps, Linear, TensorProduct,
                     spherical_harmonics, wigner_3j)


class AttrDict(dict):
    """Tiny EasyDict / PyG-Data stand-in: attribute + item access."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def to_dict(self):
        return dict(self)


def scatter_sum(src, index, dim_size):
    out = src.new_zeros((dim_size,) + src.shape[1:])
    return out.index_add_(0, index, src)


# ---------------------------------------------------------------------------- building blocks
def fuse_src_dst(irreps: Irreps, xs, xd):
    """AttentionHeadsToVector on stack([src, dst], dim=-2): per irrep chunk [src chunk | dst chunk]."""
    out = []
    for sl in irreps.slices():
        out += [xs[:, sl], xd[:, sl]]
    return torch.cat(out, dim=1)


def scale_irreps(irreps: Irreps, factor) -> Irreps:
    return Irreps([(max(1, int(mul * factor)), ir) for mul, ir in irreps])


def tp_instructions(irreps1: Irreps, irreps2: Irreps, target: Irreps, mode="uvw", trainable=True):
    out_list, instr = [], []
    for i, (mul_in, ir_in) in enumerate(irreps1):
        for j, (_, ir_e) in enumerate(irreps2):
            for _, (mul_out, ir_out) in enumerate(target):
                if ir_out in ir_in * ir_e:
                    k = len(out_list)
                    out_list.append((mul_out if mode == "uvw" else mul_in, ir_out))
                    instr.append((i, j, k, mode, trainable))
    irreps_mid, perm, _ = Irreps(out_list).sort()
    instr = sorted([(a, b, perm[c], m, t) for a, b, c, m, t in instr], key=lambda x: x[2])
    return irreps_mid, instr


class LinearScaleWithWeights(nn.Module):
    def __init__(self, irreps_in, irreps_out):
        super().__init__()
        irreps_in = Irreps(irreps_in)
        instr = [(i, 0, i, "uvu", True) for i in range(len(irreps_in))]
        self.tp = TensorProduct(irreps_in, Irreps("1x0e"), irreps_in, instr,
                                internal_weights=False, shared_weights=False)
        self.weight_numel = self.tp.weight_numel
        self.linear_out = Linear(irreps_in, irreps_out)

    def forward(self, x, weight):
        y = torch.ones_like(x[:, 0:1])
        return self.linear_out(self.tp(x, y, weight))


class MessagePackBlock(nn.Module):
    def __init__(self, irreps_node_feats, irreps_edge_feats, irreps_local_env_edge, irreps_out,
                 irreps_edge_scalars, radial_MLP=(64, 64)):
        super().__init__()
        self.irreps_node_feats = Irreps(irreps_node_feats)
        self.irreps_edge_feats = Irreps(irreps_edge_feats)
        self.irreps_sh = Irreps(irreps_local_env_edge)
        self.irreps_out = Irreps(irreps_out)
        self.combined = scale_irreps(self.irreps_node_feats, 2)
        self.mid_node, self.node_instr = tp_instructions(self.combined, self.irreps_sh, self.irreps_out)
        self.mid_edge, self.edge_instr = tp_instructions(self.irreps_edge_feats, self.irreps_sh, self.irreps_out)
        self.node_tensor_product = TensorProduct(self.combined, self.irreps_sh, self.mid_node, self.node_instr)
        self.edge_tensor_product = TensorProduct(self.irreps_edge_feats, self.irreps_sh, self.mid_edge, self.edge_instr)
        self.node_linear_scaler = LinearScaleWithWeights(self.mid_node.simplify(), self.irreps_out)
        self.edge_linear_scaler = LinearScaleWithWeights(self.mid_edge.simplify(), self.irreps_out)
        nin = Irreps(irreps_edge_scalars).num_irreps
        self.node_weight_generator = FullyConnectedNet([nin] + list(radial_MLP) + [self.node_linear_scaler.weight_numel])
        self.edge_weight_generator = FullyConnectedNet([nin] + list(radial_MLP) + [self.edge_linear_scaler.weight_numel])
        self.node_linear_out = Linear(self.irreps_out, self.irreps_out)
        self.edge_linear_out = Linear(self.irreps_out, self.irreps_out)

    def forward(self, xs, xd, e, sh, scalars):
        node_inter = fuse_src_dst(self.irreps_node_feats, xs, xd)
        w_node = self.node_weight_generator(scalars)
        node_up = self.node_tensor_product(node_inter, sh)
        node_dn = self.node_linear_scaler(node_up, w_node)
        w_edge = self.edge_weight_generator(scalars)
        edge_up = self.edge_tensor_product(e, sh)
        edge_dn = self.edge_linear_scaler(edge_up, w_edge)
        return self.node_linear_out(node_dn) + self.edge_linear_out(edge_dn)


class EmbeddingTP(nn.Module):
    """TensorProductWithMemoryOptimizationWithWeight (uvw mode)."""

    def __init__(self, irreps_input_1, irreps_input_2, irreps_out, irreps_scalar, radial_MLP):
        super().__init__()
        i1, i2, io = Irreps(irreps_input_1), Irreps(irreps_input_2), Irreps(irreps_out)
        self.irreps_mid, self.instructions = tp_instructions(i1, i2, io)
        self.tensor_product = TensorProduct(i1, i2, self.irreps_mid, self.instructions)
        self.linear_scaler = LinearScaleWithWeights(self.irreps_mid.simplify(), io)
        nin = Irreps(irreps_scalar).num_irreps
        self.weight_generator = FullyConnectedNet([nin] + list(radial_MLP) + [self.linear_scaler.weight_numel])

    def forward(self, x, y, scalars):
        return self.linear_scaler(self.tensor_product(x, y), self.weight_generator(scalars))


def irreps2gate(irreps: Irreps):
    scal = Irreps([(m, ir) for m, ir in irreps if ir.l == 0]).simplify()
    gated = Irreps([(m, ir) for m, ir in irreps if ir.l != 0]).simplify()
    gates = Irreps([(m, "0e") for m, _ in gated]).simplify() if gated.dim > 0 else Irreps([])
    act_s = [{1: "ssp", -1: "tanh"}[ir.p] for _, ir in scal]
    act_g = [{1: "ssp", -1: "abs"}[ir.p] for _, ir in gates]
    return scal, gates, gated, act_s, act_g


class ResidualBlock(nn.Module):
    def __init__(self, irreps_in, feature_irreps_hidden, resnet=True):
        super().__init__()
        self.irreps_in = Irreps(irreps_in)
        scal, gates, gated, a_s, a_g = irreps2gate(Irreps(feature_irreps_hidden))
        self.equivariant_nonlin = Gate(scal, a_s, gates, a_g, gated)
        self.linear1 = Linear(self.irreps_in, self.equivariant_nonlin.irreps_in)
        self.linear2 = Linear(self.equivariant_nonlin.irreps_out, self.irreps_in)
        self.resnet = resnet

    def forward(self, x):
        y = self.linear2(self.equivariant_nonlin(self.linear1(x)))
        return x + y if self.resnet else y


class ConvBlockE3(nn.Module):
    def __init__(self, irreps_in, irreps_out, irreps_edge_attrs, irreps_edge_embed, radial_MLP):
        super().__init__()
        self.residual = ResidualBlock(irreps_in, irreps_out)
        self.conv_tp = MessagePackBlock(irreps_in, irreps_in, irreps_edge_attrs, irreps_out, irreps_edge_embed, radial_MLP)
        self.skip_linear = Linear(irreps_in, irreps_out)

    def forward(self, data):
        sender, receiver = data["edge_index"]
        x = data["node_features"]
        skip = self.skip_linear(x)
        m = self.conv_tp(x[sender], x[receiver], data["edge_features"], data["edge_attrs"], data["edge_embedding"])
        agg = scatter_sum(m, receiver, x.shape[0])
        out = self.residual(agg) + skip
        data["node_features"] = out
        return out


class PairInteractionBlock(nn.Module):
    def __init__(self, irreps_node_feats, irreps_edge_attrs, irreps_edge_embed, irreps_edge_feats,
                 use_skip_connections, legacy_edge_update, radial_MLP):
        super().__init__()
        self.use_skip_connections, self.legacy_edge_update = use_skip_connections, legacy_edge_update
        self.linear_up_src = Linear(irreps_node_feats, irreps_node_feats)
        self.linear_up_tar = Linear(irreps_node_feats, irreps_node_feats)
        self.conv_tp = MessagePackBlock(irreps_node_feats, irreps_edge_feats, irreps_edge_attrs, irreps_edge_feats,
                                        irreps_edge_embed, radial_MLP)
        if use_skip_connections:
            self.skip_linear = Linear(irreps_edge_feats, irreps_edge_feats)

    def forward(self, data):
        src, dst = data["edge_index"]
        x, e = data["node_features"], data["edge_features"]
        mix = self.conv_tp(self.linear_up_src(x)[src], self.linear_up_tar(x)[dst], e,
                           data["edge_attrs"], data["edge_embedding"])
        if self.use_skip_connections:
            e = mix + self.skip_linear(e)
        elif self.legacy_edge_update:
            pass
        else:
            e = mix
        data["edge_features"] = e
        return e


class PairInteractionEmbeddingBlock(nn.Module):
    def __init__(self, irreps_node_feats, irreps_edge_attrs, irreps_edge_embed, irreps_edge_feats, radial_MLP):
        super().__init__()
        self.linear_up_src = Linear(irreps_node_feats, irreps_node_feats)
        self.linear_up_dst = Linear(irreps_node_feats, irreps_node_feats)
        self.conv_tp = EmbeddingTP(irreps_node_feats, irreps_edge_attrs, irreps_edge_feats, irreps_edge_embed, radial_MLP)

    def forward(self, data):
        src, dst = data["edge_index"]
        x = data["node_features"]
        h = self.linear_up_src(x[src]) + self.linear_up_dst(x[dst])
        data["edge_features"] = self.conv_tp(h, data["edge_attrs"], data["edge_embedding"])
        return data["edge_features"]


class _Wrap(nn.Module):
    def __init__(self, lin):
        super().__init__()
        self.linear = lin


# ---------------------------------------------------------------------------- HamGNN_pre
DEFAULT_PRE = dict(
    cutoff=26.0, irreps_edge_sh="0e + 1o + 2e + 3o + 4e + 5o", edge_sh_normalization="component",
    edge_sh_normalize=True,
    irreps_node_features="64x0e+64x0o+32x1o+16x1e+12x2o+25x2e+18x3o+9x3e+4x4o+9x4e+4x5o+4x5e+2x6e",
    num_layers=3, num_radial=64, num_types=96, rbf_func="bessel", radial_MLP=[64, 64],
    legacy_edge_update=False)


class HamGNNConvE3(nn.Module):
    """hamgnn/models/hamgnn_conv.py:88-284 (default branch: bessel rbf, no corr-prod, no charge doping,
    no internal graph, no gradient checkpointing)."""

    def __init__(self, cfg: Dict):
        super().__init__()
        c = dict(DEFAULT_PRE)
        c.update(cfg)
        self.cfg = c
        self.num_types = c["num_types"]
        self.irreps_edge_sh = Irreps(c["irreps_edge_sh"])
        self.irreps_node_features = Irreps(c["irreps_node_features"])
        self.cutoff = float(c["cutoff"])
        self.num_radial = c["num_radial"]
        self.num_layers = c["num_layers"]
        if c["rbf_func"].lower() != "bessel":
            raise ValueError(f"Unsupported radial basis function: {c['rbf_func']}")
        ir_attr = Irreps([(self.num_types, (0, 1))])
        ir_emb = Irreps([(self.num_radial, (0, 1))])
        D = self.irreps_node_features
        rm = c["radial_MLP"]
        self.pair_embedding = PairInteractionEmbeddingBlock(ir_attr, self.irreps_edge_sh, ir_emb, D, rm)
        self.chemical_embedding = _Wrap(Linear(ir_attr, D))
        self.convolutions = nn.ModuleList()
        self.pair_interactions = nn.ModuleList()
        legacy = c["legacy_edge_update"]
        for i in range(self.num_layers):
            self.convolutions.append(ConvBlockE3(D, D, self.irreps_edge_sh, ir_emb, rm))
            self.pair_interactions.append(PairInteractionBlock(
                D, self.irreps_edge_sh, ir_emb, D,
                use_skip_connections=((i > 0) if legacy else True), legacy_edge_update=legacy, radial_MLP=rm))

    # -- a1 / a2
    def edge_geometry(self, data):
        j, i = data["edge_index"]
        vec = (data["pos"][i] + data["nbr_shift"]) - data["pos"][j]
        ls = [ir.l for _, ir in self.irreps_edge_sh]
        unit = torch.nn.functional.normalize(vec, dim=-1)
        data["edge_attrs"] = spherical_harmonics(ls, unit[:, [1, 2, 0]], self.cfg["edge_sh_normalize"],
                                                 self.cfg["edge_sh_normalization"])
        r = vec.norm(dim=-1)
        data["edge_vectors"] = vec / r[:, None]
        data["edge_lengths"] = r
        freqs = torch.arange(1, self.num_radial + 1, dtype=vec.dtype) * math.pi / self.cutoff
        rbf = torch.sin(r[:, None] * freqs[None, :]) / r[:, None]
        cut = 0.5 * (torch.cos(r * math.pi / self.cutoff) + 1.0) * (r < self.cutoff).to(vec.dtype)
        data["edge_embedding"] = rbf * cut[:, None]

    def forward(self, data):
        onehot = torch.nn.functional.one_hot(data["z"], num_classes=self.num_types).to(data["pos"].dtype)
        data["node_attrs"] = onehot
        data["node_features"] = onehot
        self.edge_geometry(data)
        self.pair_embedding(data)
        data["node_features"] = self.chemical_embedding.linear(data["node_features"])
        for i in range(self.num_layers):
            self.convolutions[i](data)
            self.pair_interactions[i](data)
        return AttrDict(node_attr=data["node_features"], edge_attr=data["edge_features"])


# ---------------------------------------------------------------------------- HamGNN_out
def openmx_basis(nao_max):
    """hamgnn/models/hamgnn_output.py:367-526 -- (index_change, row irreps, basis_def)."""
    s1, s2, s3 = [0], [1], [2]
    p1, p2 = [3, 4, 5], [6, 7, 8]
    d1, d2 = [9, 10, 11, 12, 13], [14, 15, 16, 17, 18]
    f1 = [19, 20, 21, 22, 23, 24, 25]
    if nao_max == 14:
        idx = [0, 1, 2, 5, 3, 4, 8, 6, 7, 11, 13, 9, 12, 10]
        row = "1x0e+1x0e+1x0e+1x1o+1x1o+1x2e"
        a = s1 + s2 + p1
        b = s1 + s2 + s3 + p1 + p2
        c = s1 + s2 + p1 + p2
        d = s1 + s2 + p1 + p2 + d1
        full = list(range(14))
        bd = {1: a, 2: a, 3: b, 4: c, 5: d, 6: d, 7: d, 8: d, 9: d, 10: d, 11: full, 12: full, 13: d, 14: d, 15: d,
              16: d, 17: d, 18: d, 19: full, 20: full, 35: full, 23: full, 25: full}
    elif nao_max == 13:
        idx = [0, 1, 4, 2, 3, 7, 5, 6, 10, 12, 8, 11, 9]
        row = "1x0e+1x0e+1x1o+1x1o+1x2e"
        full = list(range(13))
        bd = {1: [0, 1, 2, 3, 4], 5: full, 6: full, 7: full, 8: full}
    elif nao_max == 19:
        idx = [0, 1, 2, 5, 3, 4, 8, 6, 7, 11, 13, 9, 12, 10, 16, 18, 14, 17, 15]
        row = "1x0e+1x0e+1x0e+1x1o+1x1o+1x2e+1x2e"
        a = s1 + s2 + p1
        b = s1 + s2 + s3 + p1 + p2
        c = s1 + s2 + p1 + p2
        d = s1 + s2 + p1 + p2 + d1
        f14 = list(range(14))
        f19 = list(range(19))
        bd = {1: a, 2: a, 3: b, 4: c, 5: d, 6: d, 7: d, 8: d, 9: d, 10: d, 11: f14, 12: f14, 13: d, 14: d, 15: d,
              16: d, 17: d, 18: d, 19: f14, 20: f14, 25: f14, 42: f19, 83: f19, 34: f19, 24: f14, 53: f19, 28: f14,
              35: f19, 26: f14, 77: f19, 52: f19, 23: f14, 51: f19}
    elif nao_max == 26:
        idx = [0, 1, 2, 5, 3, 4, 8, 6, 7, 11, 13, 9, 12, 10, 16, 18, 14, 17, 15, 22, 23, 21, 24, 20, 25, 19]
        row = "1x0e+1x0e+1x0e+1x1o+1x1o+1x2e+1x2e+1x3o"
        s2p1 = s1 + s2 + p1
        s3p2 = s1 + s2 + s3 + p1 + p2
        s2p2 = s1 + s2 + p1 + p2
        s2p2d1 = s2p2 + d1
        s3p2d1 = s3p2 + d1
        s3p2d2 = s3p2 + d1 + d2
        s3p2d2f1 = s3p2d2 + f1
        bd = {1: s2p1, 2: s2p1, 3: s3p2, 4: s2p2}
        for Z in (5, 6, 7, 8, 9, 10, 13, 14, 15, 16, 17, 18):
            bd[Z] = s2p2d1
        for Z in (11, 12, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30):
            bd[Z] = s3p2d1
        for Z in list(range(31, 52)) + [54, 55, 56]:
            bd[Z] = s3p2d2
        for Z in (52, 53, 57, 58, 59, 60, 61, 62, 66, 67, 71, 72, 73, 74, 75, 76, 77, 78, 79, 80, 81, 82, 83):
            bd[Z] = s3p2d2f1
    else:
        raise NotImplementedError(f"NAO max '{nao_max}' not supported for 'openmx'.")
    return idx, Irreps(row), bd


class SU2Decomposition:
    """E3TensorDecomposition(None, out_js_list, spinful=True) of hamgnn/nn/tensor_decomposition.py:439-627.

    Per (l1, l2) block the network output holds the irreps  (+)_L L  followed by, for every L, (L x 1) =
    (+)_{L'=|L-1|}^{L+1} L'  (irreps_from_l1l2, :39-86), all with parity (-1)^(l1+l2); the whole list is then
    repeated once more (real half | imaginary half, :543-544).  get_H (:575-627) recouples (L x 1) with
    w3j(L, 1, L'), couples (L, n) back to the orbital pair with w3j(l1, l2, L) and maps the spin index
    n = (scalar; y, z, x) to the four spin blocks (uu, ud, du, dd) with `oyzx2spin` (:557-564)."""

    def __init__(self, out_js_list, nao_max):
        self.out_js_list, self.nao_max = list(out_js_list), nao_max
        base = []
        self.in_slices, self.in_slices_sp, self.wms, self.wms_sp = [0], [], [], []
        dim = 0
        for l1, l2 in self.out_js_list:
            p = (-1) ** (l1 + l2)
            Ls = list(range(abs(l1 - l2), l1 + l2 + 1))
            base += [(1, (L, p)) for L in Ls]
            sl = [0, sum(2 * L + 1 for L in Ls)]
            wm_sp = [None]
            for L in Ls:
                L1s = list(range(abs(L - 1), L + 2))
                base += [(1, (Lp, p)) for Lp in L1s]
                sl.append(sl[-1] + sum(2 * Lp + 1 for Lp in L1s))
                wm_sp.append(torch.cat([wigner_3j(L, 1, Lp, dtype=torch.float64) for Lp in L1s], dim=-1))
            dim += sl[-1]
            self.in_slices.append(dim)
            self.in_slices_sp.append(sl)
            self.wms.append(torch.cat([wigner_3j(l1, l2, L, dtype=torch.float64) for L in Ls], dim=-1))
            self.wms_sp.append(wm_sp)
        self.base_irreps = Irreps(base)
        self.required_irreps_out = Irreps(base + base)          # :543-544
        s2 = math.sqrt(2.0)
        self.oyzx2spin = torch.tensor([[1, 0, 1, 0], [0, -1j, 0, 1], [0, 1j, 0, 1], [1, 0, -1, 0]],
                                      dtype=torch.complex128) / s2

    def get_H(self, net_out):
        cdt = torch.complex128 if net_out.dtype == torch.float64 else torch.complex64
        half = net_out.shape[-1] // 2
        c = torch.complex(net_out[:, :half], net_out[:, half:])
        n = c.shape[0]
        block = torch.zeros(n, 4, self.nao_max, self.nao_max, dtype=cdt)
        nrow = int(math.isqrt(len(self.out_js_list)))
        si = sj = 0
        for i, (l1, l2) in enumerate(self.out_js_list):
            blk = c[:, self.in_slices[i]:self.in_slices[i + 1]]
            d1, d2 = 2 * l1 + 1, 2 * l2 + 1
            sl = self.in_slices_sp[i]
            parts = []
            for j in range(1, len(self.wms_sp[i])):
                parts.append(torch.einsum("jkl,il->ijk", self.wms_sp[i][j].to(cdt), blk[:, sl[j]:sl[j + 1]]))
            Hb = torch.cat([blk[:, sl[0]:sl[1]].unsqueeze(-1), torch.cat(parts, dim=-2)], dim=-1)   # [n, d1*d2, 4]
            Hb = torch.einsum("imn,klm,jn->ijkl", Hb, self.wms[i].to(cdt), self.oyzx2spin.to(cdt))
            block[:, :, si:si + d1, sj:sj + d2] += Hb
            if (i + 1) % nrow == 0:
                si, sj = si + d1, 0
            else:
                sj += d2
        return block


class HamLayer(nn.Module):
    def __init__(self, irreps_in, feature_irreps_hidden, irreps_out):
        super().__init__()
        self.residual_block = ResidualBlock(irreps_in, feature_irreps_hidden, resnet=True)
        self.linear_transform = Linear(irreps_in, irreps_out)

    def forward(self, x):
        return self.linear_transform(self.residual_block(x))


class HamGNNPlusPlusOut(nn.Module):
    """hamgnn/models/hamgnn_output.py (openmx basis): the non-SOC branch, the two SOC branches (soc_basis 'su2' and
    'so3', non-collinear, not spin-constrained) and the overlap head (ham_only=False)."""

    def __init__(self, irreps_in_node, irreps_in_edge, nao_max=19, ham_type="openmx", ham_only=True,
                 symmetrize=True, add_H0=False, zero_point_shift=False, calculate_sparsity=True,
                 soc_switch=False, soc_basis="so3", add_H_nonsoc=False):
        super().__init__()
        assert ham_type.lower() == "openmx"
        self.nao_max, self.symmetrize, self.add_H0 = nao_max, symmetrize, add_H0
        self.ham_only, self.zero_point_shift, self.calculate_sparsity = ham_only, zero_point_shift, calculate_sparsity
        self.soc_switch, self.soc_basis, self.add_H_nonsoc = soc_switch, soc_basis.lower(), add_H_nonsoc
        idx, self.row, self.basis_def = openmx_basis(nao_max)
        self.col = self.row
        self.index_change = torch.tensor(idx, dtype=torch.long)
        irs = []
        for _, li in self.row:
            for _, lj in self.col:
                for L in range(abs(li.l - lj.l), li.l + lj.l + 1):
                    irs.append((1, (L, (-1) ** (li.l + lj.l))))
        self.hamiltonian_irreps = Irreps(irs)
        self.ham_dims = [ir.dim for _, ir in self.hamiltonian_irreps]
        self.onsite_hamiltonian_network = HamLayer(irreps_in_node, irreps_in_node, self.hamiltonian_irreps)
        self.offsite_hamiltonian_network = HamLayer(irreps_in_edge, irreps_in_edge, self.hamiltonian_irreps)
        if soc_switch:
            if self.soc_basis == "su2":                                    # :190-198, 281-293
                js = [(li.l, lj.l) for _, li in self.row for _, lj in self.col]
                self.hamiltonian_decomposition = SU2Decomposition(js, nao_max)
                self.hamiltonian_irreps_su2 = self.hamiltonian_decomposition.required_irreps_out
                head = 2 * self.hamiltonian_irreps_su2                     # list repetition, as `2*o3.Irreps`
                self.onsite_hamiltonian_network = HamLayer(irreps_in_node, irreps_in_node, head)
                self.offsite_hamiltonian_network = HamLayer(irreps_in_edge, irreps_in_edge, head)
            elif self.soc_basis == "so3":                                  # :200-209
                ksi = Irreps([(nao_max ** 2, (0, 1))])
                self.onsite_ksi_network = HamLayer(irreps_in_node, irreps_in_node, ksi)
                self.offsite_ksi_network = HamLayer(irreps_in_edge, irreps_in_edge, ksi)
            else:
                raise NotImplementedError(f"SOC basis '{soc_basis}' not supported!")
        if not ham_only:                                                   # :245-254
            self.onsite_overlap_network = HamLayer(irreps_in_node, irreps_in_node, self.hamiltonian_irreps)
            self.offsite_overlap_network = HamLayer(irreps_in_edge, irreps_in_edge, self.hamiltonian_irreps)

    def merge_tensor_components(self, comps):
        B = comps[0].shape[0]
        H = comps[0].new_zeros(B, self.nao_max, self.nao_max)
        k, r0 = 0, 0
        for _, li in self.row:
            c0 = 0
            for _, lj in self.col:
                for L in range(abs(li.l - lj.l), li.l + lj.l + 1):
                    cg = math.sqrt(2 * L + 1) * wigner_3j(li.l, lj.l, L, dtype=comps[0].dtype).unsqueeze(0)
                    H[:, r0:r0 + li.dim, c0:c0 + lj.dim] += (cg * comps[k][:, None, None, :]).sum(-1)
                    k += 1
                c0 += lj.dim
            r0 += li.dim
        return H.reshape(B, -1)

    def reorder_matrix(self, M):
        m = M.reshape(-1, self.nao_max, self.nao_max)
        m = m[:, self.index_change[:, None], self.index_change[None, :]]
        return m.reshape(-1, self.nao_max ** 2)

    def _sym(self, M, inv=None, hermitian=True):
        """symmetrize_hamiltonian :1231-1285 (real blocks): 0.5 (H +- H[inv]^T)."""
        if not self.symmetrize:
            return M
        m = M.reshape(-1, self.nao_max, self.nao_max)
        other = m if inv is None else m[inv]
        sgn = 1.0 if hermitian else -1.0
        return (0.5 * (m + sgn * other.permute(0, 2, 1))).reshape(-1, self.nao_max ** 2)

    def symmetrize_orbital_coefficients(self, M):
        """:2367-2431 -- average over the m components of every p/d/f shell, rows first, then columns."""
        m = M.reshape(-1, self.nao_max, self.nao_max).clone()
        blocks = [(3, 6), (6, 9), (9, 14)] if self.nao_max >= 14 else []
        if self.nao_max >= 19:
            blocks.append((14, 19))
        if self.nao_max == 26:
            blocks.append((19, 26))
        for a, b in blocks:
            m[:, a:b] = m[:, a:b].mean(dim=1, keepdim=True).expand(-1, b - a, -1)
        for a, b in blocks:
            m[:, :, a:b] = m[:, :, a:b].mean(dim=2, keepdim=True).expand(-1, -1, b - a)
        return m.reshape(M.shape[0], -1)

    def _mask(self, Hon, Hoff, data):
        tab = torch.zeros(99, self.nao_max, dtype=Hon.real.dtype)
        for Z, orb in self.basis_def.items():
            tab[Z, orb] = 1
        z = data["z"]
        src, dst = data["edge_index"]
        mo = tab[z]
        on = (mo[:, :, None] * mo[:, None, :]).reshape(Hon.shape)
        off = (tab[z[src]][:, :, None] * tab[z[dst]][:, None, :]).reshape(Hoff.shape)
        return Hon * on, Hoff * off

    def concat_by_crystal(self, data, on, off):
        counts = data["node_counts"].tolist()
        src = data["edge_index"][0]
        epc = scatter_sum(torch.ones_like(src), data["batch"][src], len(counts)).tolist()
        ons, offs = torch.split(on, counts, 0), torch.split(off, epc, 0)
        out = []
        for a, b in zip(ons, offs):
            out += [a, b]
        return torch.cat(out, 0)

    def sparsity_ratio(self, data):
        z = data["z"]
        n_orb = torch.full((256,), self.nao_max, dtype=torch.long)
        defined = torch.zeros(256, dtype=torch.bool)
        for Z, orb in self.basis_def.items():
            n_orb[Z] = len(orb)
            defined[Z] = True
        nn2 = self.nao_max ** 2
        total = eff = 0
        if "Hon" in data:
            total += z.numel() * nn2
            eff += int((n_orb[z] ** 2).sum())
        if "Hoff" in data:
            src, dst = data["edge_index"]
            total += src.numel() * nn2
            both = defined[z[src]] & defined[z[dst]]
            eff += int(torch.where(both, n_orb[z[src]] * n_orb[z[dst]], torch.full_like(src, nn2)).sum())
        return torch.tensor(total / eff if eff > 0 else float("inf"), dtype=torch.float32)

    def forward(self, data, rep):
        missing = [int(Z) for Z in data["z"].unique().tolist() if int(Z) not in self.basis_def]
        if missing:
            raise ValueError("The following elements are missing from basis_def: " + ", ".join(f"Z={m}" for m in missing))
        if "hamiltonian" not in data and "Hon" in data:
            data["hamiltonian"] = self.concat_by_crystal(data, data["Hon"], data["Hoff"])
        if "overlap" not in data and "Son" in data:
            data["overlap"] = self.concat_by_crystal(data, data["Son"], data["Soff"])
        src = data["edge_index"][0]
        B = len(data["node_counts"])
        epc = scatter_sum(torch.ones_like(src), data["batch"][src], B)
        off0 = torch.cumsum(epc, 0) - epc
        inv = data["inv_edge_idx"] + off0[data["batch"][src]]

        nao = self.nao_max
        overlap = None
        if not self.ham_only:                                              # :2996-3019
            s_on = torch.split(self.onsite_overlap_network(rep["node_attr"]), self.ham_dims, dim=-1)
            Son = self._sym(self.reorder_matrix(self.merge_tensor_components(s_on)))
            s_off = torch.split(self.offsite_overlap_network(rep["edge_attr"]), self.ham_dims, dim=-1)
            Soff = self._sym(self.reorder_matrix(self.merge_tensor_components(s_off)), inv)
            Son, Soff = self._mask(Son, Soff, data)
            overlap = self.concat_by_crystal(data, Son, Soff)

        if self.soc_switch:
            if self.soc_basis == "su2":                                    # :3146-3178
                dec = self.hamiltonian_decomposition

                def spin_matrix(net_out):
                    Hc = dec.get_H(net_out)                                # [n, 4, nao, nao] complex
                    Hc = Hc.reshape(-1, nao, nao)[:, self.index_change[:, None], self.index_change[None, :]]
                    Hc = Hc.reshape(-1, 2, 2, nao, nao).transpose(2, 3)    # [n, s1, a, s2, b]
                    return Hc.reshape(-1, 2 * nao, 2 * nao)

                Hon = spin_matrix(self.onsite_hamiltonian_network(rep["node_attr"]))
                Hoff = spin_matrix(self.offsite_hamiltonian_network(rep["edge_attr"]))
                if self.symmetrize:
                    Hon = 0.5 * (Hon + Hon.conj().transpose(1, 2))
                    Hoff = 0.5 * (Hoff + Hoff[inv].conj().transpose(1, 2))
                Hon, Hoff = Hon.reshape(-1, 2, nao, 2, nao).clone(), Hoff.reshape(-1, 2, nao, 2, nao).clone()
                for a in range(2):
                    for b in range(2):
                        Hon[:, a, :, b, :], Hoff[:, a, :, b, :] = self._mask(Hon[:, a, :, b, :], Hoff[:, a, :, b, :], data)
                Hon, Hoff = Hon.reshape(-1, (2 * nao) ** 2), Hoff.reshape(-1, (2 * nao) ** 2)
                on_re, on_im, off_re, off_im = Hon.real, Hon.imag, Hoff.real, Hoff.imag
            else:                                                          # so3, :3026-3144
                if self.add_H_nonsoc:
                    Hon_ns, Hoff_ns = data["Hon_nonsoc"], data["Hoff_nonsoc"]
                    for key in ("Hon0", "Hoff0"):
                        h0 = data[key].reshape(-1, 2 * nao, 2 * nao).clone()
                        h0[:, :nao, :nao] = 0
                        h0[:, nao:, nao:] = 0
                        data[key] = h0.reshape(-1, (2 * nao) ** 2)
                else:
                    c_on = torch.split(self.onsite_hamiltonian_network(rep["node_attr"]), self.ham_dims, dim=-1)
                    Hon_ns = self._sym(self.reorder_matrix(self.merge_tensor_components(c_on)))
                    c_off = torch.split(self.offsite_hamiltonian_network(rep["edge_attr"]), self.ham_dims, dim=-1)
                    Hoff_ns = self._sym(self.reorder_matrix(self.merge_tensor_components(c_off)), inv)
                    Hon_ns, Hoff_ns = self._mask(Hon_ns, Hoff_ns, data)
                ksi_on = self.symmetrize_orbital_coefficients(self.onsite_ksi_network(rep["node_attr"]))
                ksi_off = self.symmetrize_orbital_coefficients(self.offsite_ksi_network(rep["edge_attr"]))

                def build(Hns, ksi, Lmat, inv_):
                    A = [self._sym(ksi * Lmat[:, :, c], inv_, hermitian=False).reshape(-1, nao, nao) for c in range(3)]
                    re = Hns.new_zeros(Hns.shape[0], 2 * nao, 2 * nao)
                    im = torch.zeros_like(re)
                    re[:, :nao, :nao] = Hns.reshape(-1, nao, nao)
                    re[:, nao:, nao:] = Hns.reshape(-1, nao, nao)
                    re[:, :nao, nao:] = A[1]
                    re[:, nao:, :nao] = A[1]
                    im[:, :nao, :nao] = A[2]
                    im[:, nao:, nao:] = -A[2]
                    im[:, :nao, nao:] = A[0]
                    im[:, nao:, :nao] = -A[0]
                    return re.reshape(-1, (2 * nao) ** 2), im.reshape(-1, (2 * nao) ** 2)

                on_re, on_im = build(Hon_ns, ksi_on, data["Lon"], None)
                off_re, off_im = build(Hoff_ns, ksi_off, data["Loff"], inv)
            if self.add_H0:                                                # :3603-3608
                on_re, off_re = on_re + data["Hon0"], off_re + data["Hoff0"]
                on_im, off_im = on_im + data["iHon0"], off_im + data["iHoff0"]
            H_re = self.concat_by_crystal(data, on_re, off_re)
            H_im = self.concat_by_crystal(data, on_im, off_im)
            if "Hon" in data and "iHon" in data:                           # :3618-3625
                data["hamiltonian_real"] = self.concat_by_crystal(data, data["Hon"], data["Hoff"])
                data["hamiltonian_imag"] = self.concat_by_crystal(data, data["iHon"], data["iHoff"])
                data["hamiltonian"] = torch.cat((data["hamiltonian_real"], data["hamiltonian_imag"]), dim=0)
            if self.zero_point_shift:                                      # :3893-3919
                S = data["overlap"].reshape(-1, nao, nao)
                Hr = H_re.reshape(-1, 2, nao, 2, nao).clone()
                Tr = data["hamiltonian_real"].reshape(-1, 2, nao, 2, nao)
                sel = S > 1e-6
                diff = (Hr[:, 0, :, 0, :] + Hr[:, 1, :, 1, :]) - (Tr[:, 0, :, 0, :] + Tr[:, 1, :, 1, :])
                shift = diff[sel].sum() / (2.0 * S[sel].sum())
                Hr[:, 0, :, 0, :] = Hr[:, 0, :, 0, :] - shift * S
                Hr[:, 1, :, 1, :] = Hr[:, 1, :, 1, :] - shift * S
                H_re = Hr.reshape(-1, (2 * nao) ** 2)
            res = {"hamiltonian": torch.cat((H_re, H_im), dim=0), "hamiltonian_real": H_re, "hamiltonian_imag": H_im,
                   "band_energy": None, "wavefunction": None}
        else:
            c_on = torch.split(self.onsite_hamiltonian_network(rep["node_attr"]), self.ham_dims, dim=-1)
            Hon = self._sym(self.reorder_matrix(self.merge_tensor_components(c_on)))
            if self.add_H0:
                Hon = Hon + data["Hon0"]
            c_off = torch.split(self.offsite_hamiltonian_network(rep["edge_attr"]), self.ham_dims, dim=-1)
            Hoff = self._sym(self.reorder_matrix(self.merge_tensor_components(c_off)), inv)
            if self.add_H0:
                Hoff = Hoff + data["Hoff0"]
            Hon, Hoff = self._mask(Hon, Hoff, data)
            H = self.concat_by_crystal(data, Hon, Hoff)
            if self.zero_point_shift:
                S = data["overlap"]
                sel = S > 1e-6
                shift = (H - data["hamiltonian"])[sel].sum() / S[sel].sum()
                H = H - shift * S
            res = {"hamiltonian": H, "band_energy": None, "wavefunction": None, "band_gap": None, "H_sym": None}
        if overlap is not None:
            res["overlap"] = overlap
        if self.calculate_sparsity:
            res["sparsity_ratio"] = self.sparsity_ratio(data)
        return res
