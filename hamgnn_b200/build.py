"""In-tree build of libhamgnn_b200.so (sm_100a only) with plain nvcc.

    python -m hamgnn_b200.build            # rebuild if sources are newer than the library
The library is the product's only compute path: hamgnn_b200.lib refuses to run without it.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhamgnn_b200.so")
SOURCES = ["capi.cu", "edge_embed.cu", "msgpack.cu", "rowops.cu", "ham.cu", "tc_gemm_test.cu", "msgpack_tc.cu", "msgpack_tcg.cu", "mma_probe.cu", "neighbor.cu", "band.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "hamgnn_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    for macro in ("HGB_ROT2_GATE_GROUPS", "HGB_ROT2_PRODUCERS"):     # experiment switches of msgpack_rot2_kernel (defaults in the sources)
        if os.environ.get(macro):
            flags.append(f"-D{macro}={int(os.environ[macro])}")
    out = os.environ.get("HGB_LIB_OUT", LIB)      # experiment builds go next to the product library
    cmd = [nvcc] + flags + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libhamgnn_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        f.write(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
