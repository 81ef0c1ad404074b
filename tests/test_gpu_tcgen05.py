"""tcgen05 3xTF32 GEMM building block vs fp64 matmul (GPU only)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N,mode", [(8, 16, 0), (32, 64, 0), (64, 25, 0), (128, 64, 0), (24, 12, 0), (56, 18, 0),
                                      (128, 4, 0), (64, 64, 1), (32, 25, 1), (8, 16, 1)])
def test_tc_gemm_selftest_matches_fp64(K, N, mode):
    from hamgnn_b200 import lib as L
    lib = L.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(K * 1000 + N)
    tiles = 5
    A = torch.randn(tiles * 128, K, generator=g).to(dev)
    B = torch.randn(K, N, generator=g).to(dev)
    C = torch.full((tiles * 128, N), float("nan"), device=dev)
    L.check(lib.hgb_tc_gemm_selftest(A.data_ptr(), B.data_ptr(), C.data_ptr(), tiles, K, N, mode, L.stream_ptr(dev)), "selftest")
    torch.cuda.synchronize()
    ref = (A.double() @ B.double())
    err = float((C.double() - ref).abs().max() / ref.abs().max())
    fp32 = float(((A @ B).double() - ref).abs().max() / ref.abs().max())
    print(f"K={K} N={N} mode={mode}: 3xTF32 rel err {err:.2e} (fp32 SIMT matmul {fp32:.2e})")
    assert err < 2e-6
