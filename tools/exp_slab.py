"""GPU experiment: do row-shifted UMMA descriptors inside a 128-byte-swizzled slab work, and with which base-offset convention?"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from detectorfreesfm_b200 import _lib
from tests.test_engine_gpu import split, ref_gemm
lib = _lib.load_library()
for bn, C, Wp in ((128, 128, 33), (64, 64, 37), (128, 128, 418)):
    shifts = [dy * Wp + dx for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    rows, M, N = 2000, 2000, bn
    g = torch.Generator().manual_seed(3)
    a = torch.randn(rows, C, generator=g); w = torch.randn(N, 9 * C, generator=g) / (9 * C) ** 0.5
    a_d, w_d = split(a).cuda(), split(w).cuda()
    ref = ref_gemm(a, w, shifts, C, M, N)
    sh = torch.tensor(shifts, dtype=torch.int32)
    for bo in (0, 1):
        out = torch.full((M, N), float("nan"), device="cuda")
        rc = lib.dfsfm_debug_gemm_slab(_lib.ptr(a_d), rows, C, _lib.ptr(w_d), N, 9, ctypes.c_void_p(sh.data_ptr()), C, bn, bo, _lib.ptr(out), M, N, None)
        torch.cuda.synchronize()
        err = (out.cpu().double() - ref).abs().max().item()
        print(f"bn={bn} C={C} Wp={Wp} base_offset_mode={bo}: rc={rc} max err {err:.3e} (scale {ref.abs().max().item():.2f})")
