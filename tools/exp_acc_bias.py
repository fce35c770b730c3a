"""GPU experiment: signed accumulation bias of the tcgen05 engine (positive operands => truncation shows as a
negative mean relative error).  Usage: python tools/exp_acc_bias.py"""
import ctypes
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from detectorfreesfm_b200 import _lib
from tests.test_engine_gpu import split, ref_gemm

lib = _lib.load_library()
for taps, cpad, bn in ((1, 256, 256), (9, 128, 128), (9, 256, 256), (25, 64, 128)):
    for sp in (0, 1):
        g = torch.Generator().manual_seed(1)
        rows, M, N = 512, 512, bn
        a = torch.rand(rows, cpad, generator=g) + 0.5
        w = (torch.rand(N, taps * cpad, generator=g) + 0.5) / (taps * cpad)
        if not sp:
            a, w = a.half().float(), w.half().float()
        shifts = [0] * taps
        out = torch.zeros(M, N, device="cuda")
        sh = torch.tensor(shifts, dtype=torch.int32)
        a_d, w_d = split(a).cuda(), split(w).cuda()  # keep alive until the launch has run
        _lib.check(lib.dfsfm_debug_gemm(_lib.ptr(a_d), rows, cpad, _lib.ptr(w_d), N, taps,
                                        ctypes.c_void_p(sh.data_ptr()), cpad, bn, sp, _lib.ptr(out), M, N, None))
        torch.cuda.synchronize()
        ref = ref_gemm(a, w, shifts, cpad, M, N)
        rel = (out.cpu().double() - ref) / ref
        ref32 = (a.double() @ w[:, :cpad].double().t()) if taps == 1 else None
        print(f"K={taps*cpad:5d} bn={bn} split={sp}: mean rel err {rel.mean().item():+.3e}  max |rel| {rel.abs().max().item():.3e}")
