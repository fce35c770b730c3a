"""Parity of the tcgen05 shifted-row GEMM engine (detectorfreesfm_b200/csrc/gemm_engine.cuh) through the C ABI
test hook dfsfm_debug_gemm, against a float64 evaluation of the same sum on the host."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def split(x):
    hi = x.half()
    lo = (x - hi.float()).half()
    return torch.stack([hi, lo], 0).contiguous()


def ref_gemm(a, w, shifts, cpad, M, N):
    """a [rows, C] fp32, w [Nrows, taps*cpad] fp32 -> [M, N] float64 with zero fill outside [0, rows)."""
    rows, C = a.shape
    a64, w64 = a.double(), w.double()
    out = torch.zeros(M, N, dtype=torch.float64)
    for t, sh in enumerate(shifts):
        idx = torch.arange(M) + sh
        ok = (idx >= 0) & (idx < rows)
        at = torch.zeros(M, cpad, dtype=torch.float64)
        cc = min(C, cpad)
        at[ok, :cc] = a64[idx[ok], :cc]
        out += at @ w64[:N, t * cpad:(t + 1) * cpad].t()
    return out


CASES = [
    # rows, C, Nrows, taps/shifts, cpad, bn, split, M, N
    (300, 64, 64, [0], 64, 64, 1, 300, 64),
    (300, 64, 64, [0], 64, 64, 0, 300, 64),
    (1000, 128, 128, [0], 128, 128, 1, 1000, 128),
    (777, 256, 256, [0], 256, 256, 1, 777, 256),
    (900, 128, 128, [-32, -31, -30, -1, 0, 1, 30, 31, 32], 128, 128, 1, 900, 128),
    (650, 208, 208, [-12, -11, -10, -1, 0, 1, 10, 11, 12], 208, 208, 1, 650, 208),
    (500, 256, 768, [0], 256, 256, 1, 500, 768),
    (400, 512, 256, [0], 512, 256, 1, 400, 256),
    (640, 64, 128, [sy * 21 + sx for sy in range(-2, 3) for sx in range(-2, 3)], 64, 128, 1, 640, 128),
    # many tiles per cluster (persistent loop, accumulator ring) and an odd number of 128-row tiles
    (128 * 331 + 17, 128, 128, [-3, 0, 5], 128, 128, 1, 128 * 331 + 17, 128),
    (128 * 75, 256, 512, [0], 256, 256, 1, 128 * 75, 512),
    (128 * 301, 64, 64, [0, 1], 64, 64, 0, 128 * 301, 64),
]


@pytest.fixture(params=[1, 2], ids=["engine1", "engine2-cta-pairs"])
def engine(request, lib):
    prev = lib.dfsfm_get_engine()
    lib.dfsfm_set_engine(request.param)
    yield request.param
    lib.dfsfm_set_engine(prev)


@pytest.mark.parametrize("rows,C,Nrows,shifts,cpad,bn,sp,M,N", CASES)
def test_debug_gemm(lib, engine, rows, C, Nrows, shifts, cpad, bn, sp, M, N):
    from detectorfreesfm_b200 import _lib
    g = torch.Generator().manual_seed(rows * 7 + C)
    a = torch.randn(rows, C, generator=g)
    w = torch.randn(Nrows, len(shifts) * cpad, generator=g) / (len(shifts) * cpad) ** 0.5
    if not sp:  # single-pass mode is exact only for fp16-representable operands
        a, w = a.half().float(), w.half().float()
    a_d, w_d = split(a).cuda(), split(w).cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    sh = torch.tensor(shifts, dtype=torch.int32)
    _lib.check(lib.dfsfm_debug_gemm(_lib.ptr(a_d), rows, C, _lib.ptr(w_d), Nrows, len(shifts), ctypes.c_void_p(sh.data_ptr()),
                                    cpad, bn, sp, _lib.ptr(out), M, N, None))
    torch.cuda.synchronize()
    ref = ref_gemm(a, w, shifts, cpad, M, N)
    err = (out.cpu().double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    # split-fp16 carries ~22 mantissa bits per operand; fp32 accumulation over K <= 1872
    assert err <= 2e-5 * max(scale, 1.0), f"max err {err} (scale {scale})"
