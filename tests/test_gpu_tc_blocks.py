"""tcgen05 kind::tf32 building blocks of the tower kernels (csrc/tcdebug.cu): every operand orientation the towers
use, both accumulator heights, against a float64 product.  fp32-class bar: the three-pass (hi, lo) product must be
as good as a plain fp32 matmul.  Only K-major operands are used by the towers: measured on B200, kind::tf32 with an
MN-major SWIZZLE_128B descriptor (modes 1 and 2 of the self-test) returns zeros, so operands that reduce over the
batch rows are written as transposed K-major images instead."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(mode, M, N, K, passes, seed=0):
    from ceo_firm_matching import _native as Nn
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(seed)
    shape_a = (K, M) if mode == 1 else (M, K)
    shape_b = (N, K) if mode == 0 else (K, N)
    A = torch.randn(*shape_a, generator=g)
    B = torch.randn(*shape_b, generator=g)
    cols = 32
    while cols < N:
        cols *= 2
    out = torch.full((128, cols), float("nan"), device=dev)
    with torch.cuda.device(dev):
        Ad_, Bd_ = A.to(dev), B.to(dev)      # keep both alive: the allocator would hand B the block of a freed A
        Nn.check(Nn.lib().cfm_tc_selftest(Nn.ptr(Ad_), Nn.ptr(Bd_), Nn.ptr(out), mode, M, N, K, passes,
                                          Nn.stream_ptr()))
        torch.cuda.synchronize()
    Ad, Bd = A.double(), B.double()
    ref = Ad @ Bd.T if mode == 0 else (Ad.T @ Bd if mode == 1 else Ad @ Bd)
    return out.cpu().double(), ref


def _rows_of(out, M):
    """TMEM lanes holding the M accumulator rows: M = 128 -> lane i; M = 64 -> lane i (rows 0-63 in lanes 0-63)
    or the 4 x 16 split (row i -> lane 32*(i//16) + i%16), whichever the hardware uses (reported on failure)."""
    if M == 128:
        return [out]
    split = torch.cat([out[32 * q:32 * q + 16] for q in range(4)], 0)
    return [out[:64], split]


@pytest.mark.parametrize("mode,M,N,K", [
    (0, 128, 64, 96), (0, 128, 32, 96), (0, 128, 64, 64), (0, 128, 64, 8), (0, 128, 64, 32), (0, 128, 16, 160),
    (0, 64, 64, 64), (0, 64, 32, 64), (0, 64, 8, 64), (0, 64, 208, 64), (0, 64, 32, 128),
])
def test_three_pass_is_fp32_class(mode, M, N, K):
    out, ref = _run(mode, M, N, K, 3)
    scale = float(ref.abs().max())
    errs = [float((cand[:, :N] - ref).abs().max()) / scale for cand in _rows_of(out, M)]
    best = int(np.argmin(errs))
    print(f"mode {mode} M{M} N{N} K{K}: rel err candidates {errs} -> layout {best}")
    assert M == 128 or best == LAYOUT_M64, f"M=64 accumulator rows are laid out as candidate {best}"
    assert errs[best] < 2e-6, f"three-pass product off by {errs[best]:.3e} of scale"


LAYOUT_M64 = 1      # index into _rows_of(): what tower_tc.cu assumes (kept in sync with tc_m64_lane())


@pytest.mark.parametrize("mode,M,N,K", [(0, 128, 64, 96), (0, 64, 208, 64)])
def test_single_pass_is_tf32_class(mode, M, N, K):
    out, ref = _run(mode, M, N, K, 1)
    scale = float(ref.abs().max())
    err = min(float((cand[:, :N] - ref).abs().max()) / scale for cand in _rows_of(out, M))
    print(f"single pass mode {mode}: {err:.3e}")
    assert 1e-6 < err < 5e-3
