"""ctypes binding of libhamgnn_b200.so -- the C ABI declared in include/hamgnn_b200.h.

There is no CPU fallback: if the library is missing (or cannot be loaded) every op raises.
`HGB_LIB` overrides the library path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HGB_LIB", os.path.join(_HERE, "libhamgnn_b200.so"))

EXPORTS = ["hgb_abi_version", "hgb_last_error", "hgb_launch_count", "hgb_edge_embed", "hgb_msgpack_forward",
           "hgb_linear_forward", "hgb_resblock_forward", "hgb_ham_assemble", "hgb_ham_finalize", "hgb_tc_gemm_selftest", "hgb_msgpack_tc_forward", "hgb_msgpack_tcg_forward",
           "hgb_linear_forward_ld", "hgb_csr_rows", "hgb_ham_finalize_su2", "hgb_ksi_shell_average", "hgb_ham_finalize_so3",
           "hgb_msgpack_tcg_forward_v2", "hgb_radial_gate", "hgb_wigner", "hgb_msgpack_rot_forward", "hgb_msgpack_rot2_forward", "hgb_msgpack_rot16_forward", "hgb_timing_enable", "hgb_timing_collect", "hgb_mma_probe", "hgb_tma_probe", "hgb_segment_sum", "hgb_neighbor_list", "hgb_edge_lookup", "hgb_band_kspace"]

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class TypeT(C.Structure):
    _fields_ = [("mul", i32), ("mpad", i32), ("l", i32), ("out_off", i32), ("path_begin", i32), ("path_end", i32),
                ("pad0", i32), ("pad1", i32)]


class PathT(C.Structure):
    _fields_ = [("kind", i32), ("branch", i32), ("src0", i32), ("nsrc", i32), ("in_off", i32), ("mul_in", i32),
                ("l1", i32), ("l2", i32), ("l3", i32), ("sh_off", i32), ("cg_off", i32), ("cg_kstart", i32),
                ("w_off", i32), ("w3_off", i32), ("lf_off", i32), ("pad0", i32)]


class MsgpackPlan(C.Structure):
    _fields_ = [("n_types", i32), ("n_paths", i32), ("n_branches", i32), ("n_sources", i32),
                ("sh_dim", i32), ("rbf_dim", i32), ("h1", i32), ("h2", i32), ("out_dim", i32),
                ("src_dim", i32 * 4), ("fc1_off", i32 * 2), ("fc2_off", i32 * 2), ("act_const", f32),
                ("types", vp), ("paths", vp), ("types_host", vp), ("paths_host", vp),
                ("cg_ij", vp), ("cg_val", vp), ("cg_kstart", vp), ("wbuf", vp)]


class RotBlockT(C.Structure):
    """One (source group, in-slot) block of the rotated + packed input of the 'rot' message kernel."""
    _fields_ = [("src0", i32), ("nsrc", i32), ("in_off", i32), ("mul", i32), ("l1", i32), ("kpad", i32), ("xoff", i32),
                ("pad", i32)]


class RotStepT(C.Structure):
    """One (path, m1) step of the 'rot' message kernel: B = X'_{m1} W_p, gated by scale * g, C_{m3} += (B.g) L'_p."""
    _fields_ = [("a_off", i32), ("w_off", i32), ("lf_off", i32), ("g_off", i32), ("scale", f32), ("kpad", C.c_int16),
                ("kind", C.c_int8), ("branch", C.c_int8), ("m3", C.c_int8), ("new_path", C.c_int8), ("pad", C.c_int16),
                ("pad2", i32)]


class RotPlan(C.Structure):
    _fields_ = [("n_blocks", i32), ("tile_stride", i32), ("lmax", i32), ("dstride", i32), ("doff", i32 * 12),
                ("step_begin", i32 * 33), ("pad", i32), ("blocks", vp), ("steps", vp), ("blocks_host", vp), ("steps_host", vp),
                ("wigner_j", vp)]


ROT2_GATE_GROUPS = int(os.environ.get("HGB_ROT2_GATE_GROUPS", "2"))   # must match the library build (include/hamgnn_b200.h)


class Rot2PassT(C.Structure):
    """One pass of the A-stationary rotated-frame kernel: output component m3 x a subset of output slots; one gate stream per
    gate-warp group (4 groups)."""
    _fields_ = [("piece_begin", i32), ("piece_end", i32), ("ncols", i32), ("out_col0", i32), ("stream_begin", i32 * ROT2_GATE_GROUPS),
                ("stream_end", i32 * ROT2_GATE_GROUPS)]


class Rot2PieceT(C.Structure):
    """One input image x the concatenated weights of every path of the pass that consumes it."""
    _fields_ = [("a_off", i32), ("w_off", i32), ("l_off", i32), ("l_floats", i32), ("gpf_begin", i32), ("dst_begin", i32),
                ("kpad", C.c_int16), ("ncols", C.c_int16), ("ndst", C.c_int16), ("gpf_n", C.c_int16)]


class Rot2GpfT(C.Structure):
    """One contiguous run of gate columns of a piece (L2 prefetch list)."""
    _fields_ = [("off", C.c_uint32), ("bytes", C.c_uint32)]


class Rot2BatchT(C.Structure):
    """Gate-stream entry = 8 consecutive B columns of a piece, in the stream of one gate-warp half.
    meta = kind (0 tensor, 1 FMA pipes, 2 dummy) | first-of-piece << 2 | last-of-piece << 3 | group-first << 4 | group-last << 5 |
    (B column / 8) << 8 | mul << 16 | acc_col0 << 21; goff_a / goff_b: float offset of the gate block (4 columns x 128 edges) of
    columns 0-3 / 4-7 inside the tile's gate block [branch][gstride][128], 0xFFFFFFFF = un-gated; l_off: FMA-pipe L' rows."""
    _fields_ = [("meta", i32), ("goff_a", C.c_uint32), ("goff_b", C.c_uint32), ("l_off", i32)]


class Rot2DstT(C.Structure):
    """One destination slot of a piece: S[s_off ...] = (B.g)[:, col0 : col0 + kcols] L'stack; C'[acc_col0 + w] += S[s_off + w]."""
    _fields_ = [("col0", C.c_int16), ("kcols", C.c_int16), ("mp", C.c_int16), ("s_off", C.c_int16), ("acc_col0", C.c_int16),
                ("mul", C.c_int16), ("l_rel", i32)]


class Rot2Plan(C.Structure):
    _fields_ = [("n_passes", i32), ("n_pieces", i32), ("n_batches", i32), ("n_dsts", i32), ("rowstride", i32), ("n_slots", i32),
                ("slot_l", i32 * 32), ("slot_mul", i32 * 32), ("slot_out_off", i32 * 32), ("ccol", (i32 * 13) * 32),
                ("passes", vp), ("pieces", vp), ("batches", vp), ("dsts", vp), ("gpf", vp), ("n_gpf", i64),
                ("passes_host", vp), ("pieces_host", vp), ("batches_host", vp), ("dsts_host", vp), ("gpf_host", vp)]


class LinBlockT(C.Structure):
    _fields_ = [("in_off", i32), ("out_off", i32), ("mul_in", i32), ("mul_out", i32), ("dim", i32), ("w_off", i32)]


class LinearPlan(C.Structure):
    _fields_ = [("n_blocks", i32), ("in_dim", i32), ("out_dim", i32), ("pad", i32), ("blocks", vp), ("w", vp)]


class GateDesc(C.Structure):
    _fields_ = [("n_scalar_slots", i32), ("sc_in_off", i32 * 4), ("sc_out_off", i32 * 4), ("sc_n", i32 * 4),
                ("sc_act", i32 * 4), ("n_gated", i32), ("gd_in_off", i32 * 16), ("gd_out_off", i32 * 16),
                ("gd_mul", i32 * 16), ("gd_dim", i32 * 16), ("gd_gate_off", i32 * 16), ("c_ssp", f32), ("c_tanh", f32),
                ("in_dim", i32), ("out_dim", i32)]


class HamPlan(C.Structure):
    _fields_ = [("nao", i32), ("n_coef", i32), ("nnz", i32), ("pad", i32), ("row_ptr", vp), ("col", vp), ("val", vp),
                ("orb_mask", vp)]


class HgbError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it is absent: the CUDA path is the only path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HgbError(f"{LIB_PATH} not found: build it with `python -m hamgnn_b200.build` "
                       "(there is no CPU fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise HgbError(f"{LIB_PATH} does not export {name}")
    lib.hgb_abi_version.restype = C.c_int
    lib.hgb_last_error.restype = C.c_char_p
    lib.hgb_launch_count.restype = i64
    lib.hgb_edge_embed.argtypes = [vp, vp, vp, i64, C.POINTER(i32), i32, f32, C.POINTER(f32), i32, vp, vp, vp, vp, vp]
    lib.hgb_msgpack_forward.argtypes = [C.POINTER(MsgpackPlan), C.POINTER(vp), C.POINTER(vp), vp, vp, i64, vp, vp, vp]
    lib.hgb_linear_forward.argtypes = [C.POINTER(LinearPlan), vp, vp, i64, vp, i32, vp]
    lib.hgb_resblock_forward.argtypes = [C.POINTER(LinearPlan), C.POINTER(GateDesc), C.POINTER(LinearPlan),
                                         C.POINTER(LinearPlan), vp, vp, i64, vp, vp]
    lib.hgb_ham_assemble.argtypes = [C.POINTER(HamPlan), vp, i64, vp, vp]
    lib.hgb_ham_finalize.argtypes = [C.POINTER(HamPlan), vp, vp, vp, vp, vp, vp, vp, i64, i32, vp, vp]
    lib.hgb_msgpack_tc_forward.argtypes = [C.POINTER(MsgpackPlan), C.POINTER(vp), C.POINTER(vp), vp, vp, vp, i64, vp, vp, vp]
    lib.hgb_msgpack_tcg_forward.argtypes = [C.POINTER(MsgpackPlan), C.POINTER(vp), C.POINTER(vp), vp, vp, C.POINTER(i32),
                                            C.POINTER(i32), i32, vp, i64, vp, vp, vp]
    lib.hgb_tc_gemm_selftest.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    lib.hgb_msgpack_tcg_forward_v2.argtypes = [C.POINTER(MsgpackPlan), C.POINTER(vp), C.POINTER(vp), vp, vp, C.POINTER(i32),
                                               C.POINTER(i32), C.POINTER(i32), i32, vp, i64, vp, vp, vp]
    lib.hgb_radial_gate.argtypes = [C.POINTER(MsgpackPlan), vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), i32, vp, i64, vp]
    lib.hgb_wigner.argtypes = [C.POINTER(RotPlan), vp, i64, vp, vp]
    lib.hgb_msgpack_rot_forward.argtypes = [C.POINTER(MsgpackPlan), C.POINTER(RotPlan), C.POINTER(vp), C.POINTER(vp), vp, vp,
                                            C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), i32, vp, vp, i64, i64, vp, vp, vp]
    lib.hgb_msgpack_rot2_forward.argtypes = [C.POINTER(MsgpackPlan), C.POINTER(RotPlan), C.POINTER(Rot2Plan), C.POINTER(vp),
                                             C.POINTER(vp), vp, vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), i32, vp, vp, vp,
                                             i64, i64, vp, vp, vp, i64, vp]
    lib.hgb_msgpack_rot16_forward.argtypes = [C.POINTER(MsgpackPlan), C.POINTER(RotPlan), C.POINTER(vp), C.POINTER(vp), vp, vp,
                                              C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), i32, vp, vp, vp, vp, i64, vp, i32,
                                              i64, i64, vp, vp, i32, vp]
    lib.hgb_band_kspace.argtypes = [vp, vp, vp, vp, i64, i32, vp, i64, vp, vp, vp, vp, vp, i32, vp, i32, vp, vp, vp]
    lib.hgb_timing_enable.argtypes = [i32]
    lib.hgb_timing_collect.argtypes = [C.POINTER(f32), C.POINTER(i64)]
    lib.hgb_mma_probe.argtypes = [i32, i32, i32, i32, i32, vp, vp]
    lib.hgb_tma_probe.argtypes = [vp, i32, i32, i32, vp, vp]
    lib.hgb_segment_sum.argtypes = [vp, i32, vp, vp, i64, vp, vp]
    lib.hgb_neighbor_list.argtypes = [vp, vp, C.POINTER(C.c_double), C.POINTER(i32), i64, vp, vp, i64, vp, vp, vp, vp]
    lib.hgb_edge_lookup.argtypes = [vp, vp, i64, i32, vp, vp, vp, i64, vp, vp]
    lib.hgb_linear_forward_ld.argtypes = [C.POINTER(LinearPlan), vp, vp, i64, vp, i64, i32, vp]
    lib.hgb_csr_rows.argtypes = [vp, vp, vp, i32, i32, vp, i64, vp, vp]
    lib.hgb_ham_finalize_su2.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp, vp, vp]
    lib.hgb_ksi_shell_average.argtypes = [i32, C.POINTER(i32), C.POINTER(i32), i32, vp, i64, vp]
    lib.hgb_ham_finalize_so3.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, vp, vp, vp]
    for name in EXPORTS[3:]:
        getattr(lib, name).restype = C.c_int
    if lib.hgb_abi_version() != 1:
        raise HgbError(f"ABI version mismatch: library {lib.hgb_abi_version()} != binding 1")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        raise HgbError(f"{what}: {load().hgb_last_error().decode()}")


KERNEL_IDS = ["radial_gate", "rotate_pack", "msgpack_rot2", "unrotate", "wigner", "edge_embed", "linear", "resblock",
              "ham_assemble", "ham_finalize", "other", "msgpack_rot", "segment_sum"]


def timing_enable(on: bool) -> None:
    load().hgb_timing_enable(1 if on else 0)


def timing_collect() -> dict:
    """{kernel name: (total ms, launches)} of the launches recorded since hgb_timing_enable(1) (synchronises)."""
    ms = (f32 * 16)()
    cnt = (i64 * 16)()
    load().hgb_timing_collect(ms, cnt)
    return {n: (float(ms[i]), int(cnt[i])) for i, n in enumerate(KERNEL_IDS) if cnt[i]}


def launch_count() -> int:
    return int(load().hgb_launch_count())


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors: torch.Tensor):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise HgbError("hamgnn_b200 ops need CUDA tensors: the B200 kernels are the only implementation "
                           "(no CPU fallback)")


def f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise HgbError(f"expected float32 tensor, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def i64c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.int64:
        raise HgbError(f"expected int64 tensor, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()
