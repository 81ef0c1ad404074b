"""cp.async.bulk cost per copy (one SM, L2-resident source) -- csrc/mma_probe.cu: hgb_tma_probe."""
import ctypes as C, sys, os
import torch
lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hamgnn_b200", "libhamgnn_b200.so"))
lib.hgb_tma_probe.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
src = torch.randn(8 * 8 * 32768 // 4 + 1024, device="cuda")
out = torch.zeros(2, dtype=torch.int64, device="cuda")
print("kind      bytes depth warps | cycles per copy per warp (pipelined) | issue cycles per copy | aggregate GB/s per SM at 1.965 GHz")
for prefetch in (0, 1):
    for nbytes in (512, 4096, 16384):
        for depth in (1, 4):
            for warps in (1, 2, 4, 8):
                if nbytes * depth * warps > 200 * 1024:
                    continue
                for rep in range(2):
                    assert lib.hgb_tma_probe(src.data_ptr(), nbytes, 256, depth | (warps << 8) | (prefetch << 16), out.data_ptr(), None) == 0
                    torch.cuda.synchronize()
                o = out.cpu().tolist()
                print(f"{'prefetch' if prefetch else 'copy    '} {nbytes:6d} {depth:3d} {warps:3d} | {o[0]:6d} | {o[1]:6d} | {warps * nbytes / max(1, o[0]) * 1.965:7.1f}")
