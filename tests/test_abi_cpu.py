"""The C ABI without a GPU: libhamgnn_b200.so loads, exports every function include/hamgnn_b200.h declares, the ctypes
mirrors in hamgnn_b200/lib.py have the C compiler's struct sizes, and argument validation fails loudly through
hgb_last_error before any CUDA call is made."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

from hamgnn_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hamgnn_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hgb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.LIB_PATH):
        from hamgnn_b200 import build
        build.build()
    return C.CDLL(L.LIB_PATH)


def test_header_functions_are_exported_and_bound(lib):
    fns = _declared_functions()
    assert len(fns) >= 20 and "hgb_msgpack_rot_forward" in fns and "hgb_wigner" in fns
    for name in fns:
        assert hasattr(lib, name), f"{L.LIB_PATH} does not export {name} (declared in include/hamgnn_b200.h)"
    assert sorted(L.EXPORTS) == fns, (sorted(set(fns) - set(L.EXPORTS)), sorted(set(L.EXPORTS) - set(fns)))


def test_ctypes_struct_sizes_match_the_c_compiler():
    structs = {"hgb_type_t": L.TypeT, "hgb_path_t": L.PathT, "hgb_msgpack_plan": L.MsgpackPlan, "hgb_linblock_t": L.LinBlockT,
               "hgb_linear_plan": L.LinearPlan, "hgb_gate_desc": L.GateDesc, "hgb_ham_plan": L.HamPlan,
               "hgb_rot_block_t": L.RotBlockT, "hgb_rot_step_t": L.RotStepT, "hgb_rot_plan": L.RotPlan,
               "hgb_rot2_pass_t": L.Rot2PassT, "hgb_rot2_piece_t": L.Rot2PieceT, "hgb_rot2_batch_t": L.Rot2BatchT,
               "hgb_rot2_dst_t": L.Rot2DstT, "hgb_rot2_gpf_t": L.Rot2GpfT, "hgb_rot2_plan": L.Rot2Plan}
    prog = '#include <stdio.h>\n#include "hamgnn_b200.h"\nint main(void){\n' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));\n' for n in structs) + "return 0;}\n"
    with tempfile.TemporaryDirectory() as d:
        cfile, exe = os.path.join(d, "s.c"), os.path.join(d, "s")
        open(cfile, "w").write(prog)
        subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        name, size = line.split()
        assert C.sizeof(structs[name]) == int(size), (name, C.sizeof(structs[name]), int(size))


def test_argument_errors_are_loud_without_a_gpu(lib):
    lib.hgb_abi_version.restype = C.c_int
    lib.hgb_last_error.restype = C.c_char_p
    assert lib.hgb_abi_version() == 1
    # NULL plan -> validation error before any CUDA call
    lib.hgb_wigner.restype = C.c_int
    rc = lib.hgb_wigner(None, None, C.c_int64(4), None, None)
    assert rc != 0 and b"hgb_wigner" in lib.hgb_last_error()
    lib.hgb_msgpack_rot_forward.restype = C.c_int
    rc = lib.hgb_msgpack_rot_forward(*([None] * 11), C.c_int64(0), C.c_int64(128), C.c_int64(4), None, None, None)
    assert rc != 0 and b"hgb_msgpack_rot_forward" in lib.hgb_last_error()
    lib.hgb_edge_embed.restype = C.c_int
    rc = lib.hgb_edge_embed(None, None, None, C.c_int64(-1), None, 0, C.c_float(1.0), None, 0, None, None, None, None, None)
    assert rc != 0 and b"hgb_edge_embed" in lib.hgb_last_error()


def test_product_refuses_to_run_without_the_library(monkeypatch):
    """No CPU fallback: a missing library raises (it is never replaced by the oracle)."""
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", os.path.join(ROOT, "does_not_exist.so"))
    with pytest.raises(L.HgbError):
        L.load()
    for mod in ("hamgnn_b200.plan", "hamgnn_b200.hamgnn_conv", "hamgnn_b200.hamgnn_output", "hamgnn_b200.lib", "hamgnn_b200.dist"):
        src = open(os.path.join(ROOT, *mod.split(".")) + ".py").read()
        assert "oracle" not in re.sub(r"#.*", "", src).replace('"""', ""), f"{mod} must not reference the oracle"


def test_streamed_forward_refuses_cpu():
    """hamgnn_b200.pipeline.streamed_forward is a CUDA-only caller like everything else on the product path."""
    import torch
    from hamgnn_b200 import lib as L
    from hamgnn_b200.pipeline import streamed_forward
    lin = torch.nn.Linear(2, 2)
    with pytest.raises(L.HgbError):
        next(streamed_forward(lin, lin, [], device="cpu"))
