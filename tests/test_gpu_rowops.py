"""GPU parity of the row-wise kernels through the C ABI (hgb_resblock_forward -> resblock2_kernel, 16-row tiles with the
weights staged in shared memory; hgb_linear_forward): ResidualBlock (Linear -> e3nn Gate -> Linear + residual, +skip) and
HamLayer (ResidualBlock -> Linear) on the default irreps, ragged row counts (1, 15, 16, 17, 1000 rows), against the
fp64 emulation of the same plans (tests/hgb_kernel_emulator.py, itself checked against the oracle modules on the CPU).
Tolerance 2e-6 relative (fp32 FMA arithmetic, no tensor-core split)."""
import os

import pytest
import torch

import hgb_kernel_emulator as EM
from hamgnn_b200.hamgnn_conv import ResidualBlock
from hamgnn_b200.hamgnn_output import HamLayer
from hamgnn_b200.irreps import Irreps
from hgb_testlib import rel_err

pytestmark = pytest.mark.gpu
D = "64x0e+64x0o+32x1o+16x1e+12x2o+25x2e+18x3o+9x3e+4x4o+9x4e+4x5o+4x5e+2x6e"
HAM = "+".join(["1x0e"] * 9 + ["1x1o"] * 12 + ["1x2e"] * 10 + ["1x1e"] * 4 + ["1x3o"] * 4 + ["1x4e"] * 3)


@pytest.mark.parametrize("n", [1, 15, 16, 17, 1000])
def test_residual_block_and_hamlayer(n):
    dev = torch.device("cuda:0")
    torch.manual_seed(n)
    rb = ResidualBlock(Irreps(D), Irreps(D))
    hl = HamLayer(Irreps(D), Irreps(HAM))
    x, extra = torch.randn(n, Irreps(D).dim), torch.randn(n, Irreps(D).dim)
    ref_rb = EM.emulate_resblock(rb, x.double(), extra=extra.double())
    ref_hl = EM.emulate_resblock(hl.residual_block, x.double(), post=hl.op, post_w=hl.linear_transform.weight.detach().double())
    rb.to(dev); hl.to(dev)
    got = {}
    for v1 in ("0", "1"):
        os.environ["HGB_RESBLOCK_V1"] = v1
        try:
            y_rb = rb.forward_cuda(x.to(dev), extra=extra.to(dev))
            y_hl = hl.forward_cuda(x.to(dev))
            torch.cuda.synchronize()
        finally:
            os.environ.pop("HGB_RESBLOCK_V1", None)
        got[v1] = (y_rb.cpu(), y_hl.cpu())
        e1, e2 = rel_err(got[v1][0], ref_rb), rel_err(got[v1][1], ref_hl)
        print(f"n={n} {'resblock_kernel' if v1 == '1' else 'resblock2_kernel'}: ResidualBlock {e1:.2e} HamLayer {e2:.2e}")
        assert e1 < 2e-6 and e2 < 2e-6
    assert tuple(got["0"][1].shape) == (n, Irreps(HAM).dim)
