"""tcgen05.mma.kind::tf32 issue / completion cost per instruction (M 128, K 8) and the TMEM-load round trip behind a queue
of MMAs -- see csrc/mma_probe.cu."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hamgnn_b200", "libhamgnn_b200.so"))
lib.hgb_mma_probe.argtypes = [C.c_int32] * 5 + [C.c_void_p, C.c_void_p]
out = torch.zeros(4, dtype=torch.int64, device="cuda")
print("mode  N count ndest | issue cyc/MMA | total cyc/MMA | tcgen05.ld + wait::ld round trip (cycles) with `count` MMAs queued")
for ts in (0, 1):
    for n in (16, 64, 96, 256):
        for ndest in (1, 2):
            if ndest * n > 384:
                continue
            for count in (2, 8, 16, 48, 192):
                for rep in range(2):
                    rc = lib.hgb_mma_probe(n, count, ndest, ts, 1, out.data_ptr(), None)
                    assert rc == 0
                    torch.cuda.synchronize()
                o = out.cpu().tolist()
                print(f"{'TS' if ts else 'SS'} {n:4d} {count:5d} {ndest:5d} | {o[0]/count:8.1f} | {o[1]/count:8.1f} | {o[2]:6d}")
