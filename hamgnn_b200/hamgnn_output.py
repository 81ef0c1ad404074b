"""`HamGNNPlusPlusOut` ("HamGNN_out") on the B200 kernels -- host-side mirror of
/root/reference/hamgnn/models/hamgnn_output.py (ctor :96-256, forward :2916-4021) for the OpenMX basis tables
(:345-526): the non-SOC branch (:3771-3799), the two spin-orbit branches (soc_basis 'su2' :3146-3178 and 'so3'
:3026-3144, non-collinear, +H0 / real;imag stacking :3603-3625, result dict :3889-3931) and the overlap head
(ham_only=False, :2996-3019, 4006-4014).

Kernels: hgb_resblock_forward (HamLayer = ResidualBlock + o3.Linear, :38-58), hgb_ham_assemble
(merge_tensor_components :851-891 + reorder_matrix :1056-1096 as one CSR product), hgb_ham_finalize
(symmetrize :1231-1285, +H0 :3782-3795, orbital masks :2288-2365, per-crystal interleave :1187-1229);
SOC: hgb_linear_forward_ld (the used half of the 4x-wide su2 head in irrep-sorted layout), hgb_csr_rows
(E3TensorDecomposition.get_H, hamgnn/nn/tensor_decomposition.py:575-627), hgb_ham_finalize_su2,
hgb_ksi_shell_average (symmetrize_orbital_coefficients :2367-2431), hgb_ham_finalize_so3.
Spin-constrained / collinear heads, band energies and the SIESTA/ABACUS basis tables raise NotImplementedError
(SURVEY.md section 8f).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch
from torch import nn

from . import lib as L
from .hamgnn_conv import ResidualBlock, _W, _invalidate_hook
from .irreps import Irreps
from .plan import HamAssembly, LinearOp, SocSU2Assembly, SortedHeadOp


def openmx_basis(nao_max: int):
    """(index_change, row irreps, basis_def) of hamgnn_output.py:367-526."""
    s1, s2, s3 = [0], [1], [2]
    p1, p2 = [3, 4, 5], [6, 7, 8]
    d1, d2 = [9, 10, 11, 12, 13], [14, 15, 16, 17, 18]
    f1 = list(range(19, 26))
    sp = s1 + s2 + p1
    if nao_max in (14, 19):
        base = {1: sp, 2: sp, 3: s1 + s2 + s3 + p1 + p2, 4: s1 + s2 + p1 + p2}
        spd = s1 + s2 + p1 + p2 + d1
        for Z in (5, 6, 7, 8, 9, 10, 13, 14, 15, 16, 17, 18):
            base[Z] = spd
        full14 = list(range(14))
        for Z in (11, 12, 19, 20):
            base[Z] = full14
        if nao_max == 14:
            for Z in (35, 23, 25):
                base[Z] = full14
            return [0, 1, 2, 5, 3, 4, 8, 6, 7, 11, 13, 9, 12, 10], Irreps("1x0e+1x0e+1x0e+1x1o+1x1o+1x2e"), base
        full19 = list(range(19))
        for Z in (25, 24, 28, 26, 23):
            base[Z] = full14
        for Z in (42, 83, 34, 53, 35, 77, 52, 51):
            base[Z] = full19
        return ([0, 1, 2, 5, 3, 4, 8, 6, 7, 11, 13, 9, 12, 10, 16, 18, 14, 17, 15],
                Irreps("1x0e+1x0e+1x0e+1x1o+1x1o+1x2e+1x2e"), base)
    if nao_max == 13:
        full = list(range(13))
        return ([0, 1, 4, 2, 3, 7, 5, 6, 10, 12, 8, 11, 9], Irreps("1x0e+1x0e+1x1o+1x1o+1x2e"),
                {1: [0, 1, 2, 3, 4], 5: full, 6: full, 7: full, 8: full})
    if nao_max == 26:
        a = s1 + s2 + p1
        b = s1 + s2 + s3 + p1 + p2
        c = s1 + s2 + p1 + p2
        d = c + d1
        e = b + d1
        f = b + d1 + d2
        g = f + f1
        base = {1: a, 2: a, 3: b, 4: c}
        for Z in (5, 6, 7, 8, 9, 10, 13, 14, 15, 16, 17, 18):
            base[Z] = d
        for Z in (11, 12) + tuple(range(19, 31)):
            base[Z] = e
        for Z in tuple(range(31, 52)) + (54, 55, 56):
            base[Z] = f
        for Z in (52, 53, 57, 58, 59, 60, 61, 62, 66, 67, 71) + tuple(range(72, 84)):
            base[Z] = g
        return ([0, 1, 2, 5, 3, 4, 8, 6, 7, 11, 13, 9, 12, 10, 16, 18, 14, 17, 15, 22, 23, 21, 24, 20, 25, 19],
                Irreps("1x0e+1x0e+1x0e+1x1o+1x1o+1x2e+1x2e+1x3o"), base)
    raise NotImplementedError(f"NAO max '{nao_max}' not supported for 'openmx'.")


class HamLayer(nn.Module):
    """hamgnn_output.py:38-58."""

    def __init__(self, irreps_in, irreps_out):
        super().__init__()
        self.residual_block = ResidualBlock(irreps_in, irreps_in)
        self.op = LinearOp(irreps_in, irreps_out)
        self.linear_transform = _W(self.op.weight_numel)

    def forward_cuda(self, x):
        return self.residual_block.forward_cuda(x, post=self.op, post_w=self.linear_transform.weight)


class SocHamLayer(nn.Module):
    """HamLayer whose trailing o3.Linear is the su2 head `2 * hamiltonian_irreps_su2` (hamgnn_output.py:190-198);
    only the half of its outputs that get_H reads is evaluated (SortedHeadOp)."""

    def __init__(self, irreps_in, assembly: SocSU2Assembly):
        super().__init__()
        self.residual_block = ResidualBlock(irreps_in, irreps_in)
        self.head = SortedHeadOp(irreps_in, assembly.head_irreps, assembly.used)
        self.linear_transform = _W(self.head.weight_numel)

    def forward_cuda(self, x):
        return self.head.forward(self.linear_transform.weight, self.residual_block.forward_cuda(x))


class HamGNNPlusPlusOut(nn.Module):
    def __init__(self, irreps_in_node=None, irreps_in_edge=None, nao_max: int = 14, return_forces: bool = False,
                 create_graph: bool = False, ham_type: str = "openmx", ham_only: bool = False, symmetrize: bool = True,
                 include_triplet: bool = False, calculate_band_energy: bool = False, num_k: int = 8, k_path=None,
                 band_num_control=None, soc_switch: bool = True, nonlinearity_type: str = "gate",
                 export_reciprocal_values: bool = False, add_H0: bool = False, soc_basis: str = "so3",
                 spin_constrained: bool = False, use_learned_weight: bool = True, minMagneticMoment: float = 0.5,
                 collinear_spin: bool = False, zero_point_shift: bool = False, add_H_nonsoc: bool = False,
                 get_nonzero_mask_tensor: bool = False, calculate_sparsity: bool = True):
        super().__init__()
        self.derivative = return_forces
        self.create_graph = create_graph
        self.nao_max = nao_max
        self.ham_type = ham_type.lower()
        self.ham_only, self.symmetrize, self.add_H0 = ham_only, symmetrize, add_H0
        self.soc_switch, self.spin_constrained, self.collinear_spin = soc_switch, spin_constrained, collinear_spin
        self.zero_point_shift, self.calculate_sparsity = zero_point_shift, calculate_sparsity
        self.calculate_band_energy = calculate_band_energy
        self.num_k = int(num_k)
        # reference _configure_band_num_control (hamgnn_output.py:812-830): dict keys -> int, ignored when reciprocal values are exported
        self.band_num_control = ({int(k): v for k, v in band_num_control.items()} if isinstance(band_num_control, dict)
                                 else band_num_control if isinstance(band_num_control, int) and not isinstance(band_num_control, bool)
                                 else None)
        self.get_nonzero_mask_tensor = get_nonzero_mask_tensor
        if self.ham_type != "openmx":
            if self.ham_type in ("siesta", "abacus", "pasp"):
                raise NotImplementedError(f"ham_type '{ham_type}' basis tables are not part of this round's hot path")
            raise NotImplementedError(f"Hamiltonian type '{self.ham_type}' is not supported.")
        self.soc_basis, self.add_H_nonsoc = soc_basis.lower(), add_H_nonsoc
        for flag, name in ((spin_constrained, "spin_constrained"), (collinear_spin, "collinear_spin"),
                           (calculate_band_energy and soc_switch, "calculate_band_energy with soc_switch"),
                           (calculate_band_energy and not ham_only, "calculate_band_energy with ham_only=False"),
                           (calculate_band_energy and isinstance(k_path, (list, tuple, str)), "k_path (pass data.k_vecs instead)"),
                           (return_forces, "return_forces"),
                           (nonlinearity_type != "gate", "nonlinearity_type!='gate'"),
                           (export_reciprocal_values, "export_reciprocal_values")):
            if flag:
                raise NotImplementedError(f"HamGNN_out option {name} is outside the B200 hot path of this round "
                                          "(SURVEY.md section 8f)")
        idx, self.row, self.basis_def = openmx_basis(nao_max)
        self.col = self.row
        self.index_change = torch.tensor(idx, dtype=torch.long)
        self.assembly = HamAssembly(self.row, self.col, idx, self.basis_def)
        self.hamiltonian_irreps = self.assembly.hamiltonian_irreps
        self.onsite_hamiltonian_network = HamLayer(Irreps(irreps_in_node), self.hamiltonian_irreps)
        self.offsite_hamiltonian_network = HamLayer(Irreps(irreps_in_edge), self.hamiltonian_irreps)
        if soc_switch:
            if self.soc_basis == "su2":
                self.soc_assembly = SocSU2Assembly(self.row, self.col, idx, self.basis_def)
                self.hamiltonian_irreps_su2 = self.soc_assembly.hamiltonian_irreps_su2
                self.onsite_hamiltonian_network = SocHamLayer(Irreps(irreps_in_node), self.soc_assembly)
                self.offsite_hamiltonian_network = SocHamLayer(Irreps(irreps_in_edge), self.soc_assembly)
                if self.onsite_hamiltonian_network.head.pos.tolist() != self.offsite_hamiltonian_network.head.pos.tolist():
                    raise NotImplementedError("su2 head: node and edge features must carry the same set of irreps")
                self.soc_assembly.build_csr(self.onsite_hamiltonian_network.head.pos)
            elif self.soc_basis == "so3":
                ksi = Irreps([(nao_max ** 2, (0, 1))])
                self.onsite_ksi_network = HamLayer(Irreps(irreps_in_node), ksi)
                self.offsite_ksi_network = HamLayer(Irreps(irreps_in_edge), ksi)
                shells = [(3, 6), (6, 9), (9, 14)] if nao_max >= 14 else []     # symmetrize_orbital_coefficients :2400-2410
                if nao_max >= 19:
                    shells.append((14, 19))
                if nao_max == 26:
                    shells.append((19, 26))
                self._shell_lo = (C.c_int32 * 8)(*[a for a, _ in shells])
                self._shell_hi = (C.c_int32 * 8)(*[b for _, b in shells])
                self._n_shells = len(shells)
            else:
                raise NotImplementedError(f"SOC basis '{soc_basis}' not supported!")
        if not ham_only:
            self.onsite_overlap_network = HamLayer(Irreps(irreps_in_node), self.hamiltonian_irreps)
            self.offsite_overlap_network = HamLayer(Irreps(irreps_in_edge), self.hamiltonian_irreps)
        self._tables: Dict[str, tuple] = {}
        self.register_load_state_dict_post_hook(_invalidate_hook)

    # ---------------------------------------------------------------------------------------------
    def _lookup(self, device):
        key = str(device)
        if key not in self._tables:
            n_orb = torch.full((256,), self.nao_max, dtype=torch.long)
            defined = torch.zeros(256, dtype=torch.bool)
            for Z, orbs in self.basis_def.items():
                n_orb[Z] = len(orbs)
                defined[Z] = True
            self._tables[key] = (n_orb.to(device), defined.to(device))
        return self._tables[key]

    def validate_elements_in_basis_def(self, data):
        z = data["z"]
        zkey = (z.data_ptr(), z._version, z.numel())
        if getattr(self, "_z_checked", None) == zkey:       # one host sync per distinct z tensor, not per forward
            return True
        zs = z.unique().cpu().tolist()
        missing = [z for z in zs if z not in self.basis_def]
        if missing:
            raise ValueError("The following elements are missing from basis_def: " + ", ".join(f"Z={m}" for m in missing))
        self._z_checked = zkey
        return True

    def calculate_sparsity_ratio(self, data):
        """hamgnn_output.py:2784-2872."""
        if "sparsity_ratio" in data:
            c = data["sparsity_ratio"]
            return c.to(device=data["z"].device, dtype=torch.float32) if torch.is_tensor(c) else \
                torch.tensor(float(c), device=data["z"].device, dtype=torch.float32)
        z = data["z"]
        n_orb, defined = self._lookup(z.device)
        nn2 = self.nao_max ** 2
        total = z.new_zeros((), dtype=torch.long)
        eff = z.new_zeros((), dtype=torch.long)
        if "Hon" in data:
            total = total + z.numel() * nn2
            eff = eff + (n_orb[z] ** 2).sum()
        if "Hoff" in data and "edge_index" in data:
            src, dst = data["edge_index"][0], data["edge_index"][1]
            total = total + src.numel() * nn2
            both = defined[z[src]] & defined[z[dst]]
            eff = eff + torch.where(both, n_orb[z[src]] * n_orb[z[dst]], torch.full_like(src, nn2)).sum()
        t, e = total.double(), eff.double()
        return torch.where(e > 0, t / e, torch.full_like(t, float("inf"))).float()

    # ---- get_nonzero_mask_tensor (hamgnn_output.py:2588-2782): element-wise table look-ups, no arithmetic
    def create_orbital_validity_mask(self, atomic_numbers):
        """[99, nao_max] 0/1 table in the dtype of `atomic_numbers` (hamgnn_output.py:2588-2613), cached per device."""
        key = ("orbmask", str(atomic_numbers.device))
        if key not in self._tables:
            m = torch.zeros(99, self.nao_max, dtype=torch.long)
            for Z, orbs in self.basis_def.items():
                m[Z, list(orbs)] = 1
            self._tables[key] = m.to(atomic_numbers.device)
        return self._tables[key].type_as(atomic_numbers)

    def _pair_masks(self, data):
        src, dst = data["edge_index"][0], data["edge_index"][1]
        z = data["z"]
        m = self.create_orbital_validity_mask(z)
        on = m[z][:, :, None] * m[z][:, None, :]
        off = m[z[src]][:, :, None] * m[z[dst]][:, None, :]
        return on, off

    def build_interaction_masks(self, data):
        """[N + E, nao^2] bool, on-site rows first (plain cat, NOT the per-crystal interleave: hamgnn_output.py:2615-2665)."""
        on, off = self._pair_masks(data)
        nn2 = self.nao_max ** 2
        return torch.cat((on.bool().reshape(-1, nn2), off.bool().reshape(-1, nn2)), dim=0)

    def build_column_wise_interaction_masks(self, data):
        """[N + E, 2, nao^2] bool (hamgnn_output.py:2667-2719)."""
        m = self.build_interaction_masks(data)
        return torch.stack([m, m], dim=1)

    def build_spin_orbit_interaction_masks(self, data):
        """([N + E, (2 nao)^2], [2 (N + E), (2 nao)^2]) bool, per-crystal interleaved rows (hamgnn_output.py:2721-2782)."""
        on, off = self._pair_masks(data)
        M = 2 * self.nao_max

        def spin(x):   # blockwise_2x2_concat(x, x, x, x)
            return x.repeat(1, 2, 2).reshape(-1, M * M).bool()

        real_imag = self.concatenate_hamiltonians_by_crystal(data, spin(on), spin(off))
        return real_imag, torch.cat((real_imag, real_imag), dim=0)

    def _row_maps(self, data):
        """Row of every on-site / off-site block in the per-crystal interleaved output
        (concatenate_hamiltonians_by_crystal :1187-1229) and the batch-global inverse edge (:2985-2990)."""
        src = data["edge_index"][0]
        batch = data["batch"]
        counts = data["node_counts"]
        B = counts.shape[0]
        eb = batch[src]
        epc = torch.zeros(B, dtype=torch.long, device=src.device).index_add_(0, eb, torch.ones_like(src))
        e_off = torch.cumsum(epc, 0) - epc
        n_off = torch.cumsum(counts, 0) - counts
        N, E = batch.shape[0], src.shape[0]
        on_row = torch.arange(N, device=src.device) + e_off[batch]
        off_row = torch.arange(E, device=src.device) + (n_off + counts)[eb]
        inv = data["inv_edge_idx"] + e_off[eb]
        return on_row, off_row, inv

    def concatenate_hamiltonians_by_crystal(self, data, on, off):
        on_row, off_row, _ = self._row_maps(data)
        out = on.new_empty((on.shape[0] + off.shape[0],) + tuple(on.shape[1:]))
        out[on_row] = on
        out[off_row] = off
        return out

    def forward(self, data, graph_representation: dict = None):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise RuntimeError("hamgnn_b200.HamGNNPlusPlusOut is inference-only in this round: call it under torch.no_grad()")
        self.validate_elements_in_basis_def(data)
        if "hamiltonian" not in data and "Hon" in data:
            data["hamiltonian"] = self.concatenate_hamiltonians_by_crystal(data, data["Hon"], data["Hoff"])
        if "overlap" not in data and "Son" in data:
            data["overlap"] = self.concatenate_hamiltonians_by_crystal(data, data["Son"], data["Soff"])
        node_attr, edge_attr = graph_representation["node_attr"], graph_representation["edge_attr"]
        L.require_cuda(node_attr, edge_attr)
        dev = node_attr.device
        src, dst = L.i64c(data["edge_index"][0]), L.i64c(data["edge_index"][1])
        z = L.i64c(data["z"])
        on_row, off_row, inv = self._row_maps(data)
        N, E, nao = node_attr.shape[0], edge_attr.shape[0], self.nao_max
        nn2 = nao * nao
        plan = self.assembly.plan(dev)
        lib, st = L.load(), L.stream_ptr(dev)

        def spinless(on_net, off_net, h0_on, h0_off, out, rows_on, rows_off):
            """CG merge + reorder + symmetrise (+H0) + masks of one (on-site, off-site) HamLayer pair."""
            coef_on = on_net.forward_cuda(node_attr)
            raw_on = torch.empty(N, nn2, device=dev, dtype=torch.float32)
            L.check(lib.hgb_ham_assemble(C.byref(plan), coef_on.data_ptr(), N, raw_on.data_ptr(), st), "hgb_ham_assemble")
            L.check(lib.hgb_ham_finalize(C.byref(plan), raw_on.data_ptr(), None, L.ptr(h0_on), z.data_ptr(), None, None,
                                         L.ptr(rows_on), N, int(self.symmetrize), out[0].data_ptr(), st), "hgb_ham_finalize")
            coef_off = off_net.forward_cuda(edge_attr)
            raw_off = torch.empty(E, nn2, device=dev, dtype=torch.float32)
            L.check(lib.hgb_ham_assemble(C.byref(plan), coef_off.data_ptr(), E, raw_off.data_ptr(), st), "hgb_ham_assemble")
            L.check(lib.hgb_ham_finalize(C.byref(plan), raw_off.data_ptr(), inv.data_ptr(), L.ptr(h0_off), z.data_ptr(),
                                         src.data_ptr(), dst.data_ptr(), L.ptr(rows_off), E, int(self.symmetrize),
                                         out[1].data_ptr(), st), "hgb_ham_finalize")

        overlap = None
        if not self.ham_only:
            overlap = torch.empty(N + E, nn2, device=dev, dtype=torch.float32)
            spinless(self.onsite_overlap_network, self.offsite_overlap_network, None, None, (overlap, overlap), on_row, off_row)

        if self.soc_switch:
            result = self._forward_soc(data, node_attr, edge_attr, spinless, on_row, off_row, inv, src, dst, z)
        else:
            H = torch.empty(N + E, nn2, device=dev, dtype=torch.float32)
            spinless(self.onsite_hamiltonian_network, self.offsite_hamiltonian_network,
                     L.f32c(data["Hon0"]) if self.add_H0 else None, L.f32c(data["Hoff0"]) if self.add_H0 else None,
                     (H, H), on_row, off_row)
            if self.zero_point_shift:
                S = data["overlap"]
                sel = S > 1e-6
                shift = ((H - data["hamiltonian"]) * sel).sum() / (S * sel).sum()
                H = H - shift * S
            result = {"hamiltonian": H, "band_energy": None, "wavefunction": None, "band_gap": None, "H_sym": None}
            if self.calculate_band_energy:
                self._band_energies(data, H, on_row, off_row, result)
            if self.get_nonzero_mask_tensor:
                result["mask"] = self.build_interaction_masks(data)
        if overlap is not None:
            result["overlap"] = overlap
        if self.calculate_sparsity:
            result["sparsity_ratio"] = self.calculate_sparsity_ratio(data)
        return result

    # ---------------------------------------------------------------------------------------------
    def _band_energies(self, data, H, on_row, off_row, result):
        """Band-energy head (hamgnn_output.py:3802-3880): k points from data.k_vecs [n_crystals, num_k, 3] if present, otherwise
        num_k random points in (-1, 1)^3 mapped with inv(cell)^T like the reference's default branch; predicted bands into the
        result, reference bands (from data.Hon / data.Hoff, when present) into data.band_energy / wavefunction / band_gap / H_sym."""
        from .band import BandEnergyHead, OPENMX_NUM_VALENCE
        if getattr(self, "_band_head", None) is None:
            self._band_head = BandEnergyHead(self.nao_max, self.basis_def, OPENMX_NUM_VALENCE, self.num_k, self.band_num_control)
        dev = H.device
        if "k_vecs" not in data:
            cells = data["cell"].detach().float().cpu().numpy().reshape(-1, 3, 3)
            ks = [(2.0 * np.random.rand(self.num_k, 3) - 1.0).dot(np.linalg.inv(c).T) for c in cells]
            data["k_vecs"] = torch.tensor(np.stack(ks), dtype=torch.float32, device=dev)
        hon, hoff = H[on_row], H[off_row]
        be, wf, gap, hsym = self._band_head(hon, hoff, data)
        result.update({"band_energy": be, "wavefunction": wf, "band_gap": gap, "H_sym": hsym})
        if "Hon" in data and "Hoff" in data:
            with torch.no_grad():
                data["band_energy"], data["wavefunction"], data["band_gap"], data["H_sym"] = self._band_head(
                    L.f32c(data["Hon"]), L.f32c(data["Hoff"]), data)

    # ---------------------------------------------------------------------------------------------
    def _forward_soc(self, data, node_attr, edge_attr, spinless, on_row, off_row, inv, src, dst, z):
        """Spin-orbit branches of forward (hamgnn_output.py:3022-3178, 3603-3625, 3889-3931), non-collinear."""
        dev = node_attr.device
        N, E, nao = node_attr.shape[0], edge_attr.shape[0], self.nao_max
        M = 2 * nao
        MM = M * M
        lib, st = L.load(), L.stream_ptr(dev)
        H = torch.empty(2 * (N + E), MM, device=dev, dtype=torch.float32)   # rows [0, N+E): real, [N+E, 2(N+E)): imaginary
        H_re, H_im = H[:N + E], H[N + E:]
        h0 = {k: (L.f32c(data[k]) if self.add_H0 else None) for k in ("Hon0", "Hoff0", "iHon0", "iHoff0")}
        if self.soc_basis == "su2":
            asm = self.soc_assembly
            tb = asm.tables(dev)
            for net, x, n, partner, na, nb, rows, kre, kim in (
                    (self.onsite_hamiltonian_network, node_attr, N, None, None, None, on_row, "Hon0", "iHon0"),
                    (self.offsite_hamiltonian_network, edge_attr, E, inv, src, dst, off_row, "Hoff0", "iHoff0")):
                coef = net.forward_cuda(x)
                raw = torch.empty(n, 2 * MM, device=dev, dtype=torch.float32)
                L.check(lib.hgb_csr_rows(tb["row_ptr"].data_ptr(), tb["col"].data_ptr(), tb["val"].data_ptr(), asm.n_out,
                                         net.head.out_dim, coef.data_ptr(), n, raw.data_ptr(), st), "hgb_csr_rows")
                L.check(lib.hgb_ham_finalize_su2(nao, tb["mask"].data_ptr(), raw.data_ptr(), L.ptr(partner), L.ptr(h0[kre]),
                                                 L.ptr(h0[kim]), z.data_ptr(), L.ptr(na), L.ptr(nb), rows.data_ptr(), n,
                                                 int(self.symmetrize), H_re.data_ptr(), H_im.data_ptr(), st),
                        "hgb_ham_finalize_su2")
        else:
            nn2 = nao * nao
            if self.add_H_nonsoc:
                hns_on, hns_off = L.f32c(data["Hon_nonsoc"]), L.f32c(data["Hoff_nonsoc"])
            else:
                hns_on = torch.empty(N, nn2, device=dev, dtype=torch.float32)
                hns_off = torch.empty(E, nn2, device=dev, dtype=torch.float32)
                spinless(self.onsite_hamiltonian_network, self.offsite_hamiltonian_network, None, None,
                         (hns_on, hns_off), None, None)
            for net, x, n, hns, lkey, partner, rows, kre, kim in (
                    (self.onsite_ksi_network, node_attr, N, hns_on, "Lon", None, on_row, "Hon0", "iHon0"),
                    (self.offsite_ksi_network, edge_attr, E, hns_off, "Loff", inv, off_row, "Hoff0", "iHoff0")):
                ksi = net.forward_cuda(x)
                L.check(lib.hgb_ksi_shell_average(nao, self._shell_lo, self._shell_hi, self._n_shells, ksi.data_ptr(), n, st),
                        "hgb_ksi_shell_average")
                lmat = L.f32c(data[lkey])
                if tuple(lmat.shape) != (n, nn2, 3):
                    raise ValueError(f"data.{lkey} must have shape [{n}, {nn2}, 3], got {tuple(lmat.shape)}")
                L.check(lib.hgb_ham_finalize_so3(nao, hns.data_ptr(), ksi.data_ptr(), lmat.data_ptr(), L.ptr(partner),
                                                 L.ptr(h0[kre]), L.ptr(h0[kim]), rows.data_ptr(), n, int(self.symmetrize),
                                                 int(self.add_H_nonsoc), H_re.data_ptr(), H_im.data_ptr(), st),
                        "hgb_ham_finalize_so3")
        if "Hon" in data and "iHon" in data:
            data["hamiltonian_real"] = self.concatenate_hamiltonians_by_crystal(data, data["Hon"], data["Hoff"])
            data["hamiltonian_imag"] = self.concatenate_hamiltonians_by_crystal(data, data["iHon"], data["iHoff"])
            data["hamiltonian"] = torch.cat((data["hamiltonian_real"], data["hamiltonian_imag"]), dim=0)
        if self.zero_point_shift:
            S = data["overlap"].reshape(-1, nao, nao)
            Hr = H_re.view(-1, 2, nao, 2, nao)
            Tr = data["hamiltonian_real"].reshape(-1, 2, nao, 2, nao)
            sel = S > 1e-6
            diff = (Hr[:, 0, :, 0, :] + Hr[:, 1, :, 1, :]) - (Tr[:, 0, :, 0, :] + Tr[:, 1, :, 1, :])
            shift = (diff * sel).sum() / (2.0 * (S * sel).sum())
            Hr[:, 0, :, 0, :] -= shift * S
            Hr[:, 1, :, 1, :] -= shift * S
        result = {"hamiltonian": H, "hamiltonian_real": H_re, "hamiltonian_imag": H_im, "band_energy": None,
                  "wavefunction": None}
        if self.get_nonzero_mask_tensor:
            result["mask_real_imag"] = self.build_spin_orbit_interaction_masks(data)[0]
        return result
