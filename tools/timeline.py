"""In-kernel timeline of the engine-2 GEMM launches of one HP-1 pair (dfsfm_debug_timeline_*).  Usage: python tools/timeline.py [first] [count]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from detectorfreesfm_b200 import B200LoFTR, _lib
from tests import weights
from tests import util

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 200
lib = _lib.load_library()
raw = len(sys.argv) > 3 and sys.argv[3] == "raw"   # raw: medians (and max) of all 16 stamps of every launch
m = B200LoFTR(util.loftr_config()).cuda().eval()
m.load_state_dict(weights.loftr_state_dict(0, calibrated=True))
ims = [im.cuda() for im in util.synth_scene(2, 832, 832, 1000, noise=0.025)[0]]
for _ in range(2):
    m({"image0": ims[0], "image1": ims[1]})
torch.cuda.synchronize()
N = 256
assert lib.dfsfm_debug_timeline_arm(N) == 0
m({"image0": ims[0], "image1": ims[1]})
stamps = np.zeros((N, 148, 16), dtype=np.uint64)
info = np.zeros((N, 2), dtype=np.int32)
n = lib.dfsfm_debug_timeline_read(stamps.ctypes.data_as(ctypes.c_void_p), info.ctypes.data_as(ctypes.c_void_p), N)
lib.dfsfm_debug_timeline_arm(0)
print("launches captured", n)
st = stamps[:n].astype(np.int64)
print("idx ctas tiles |  total | setup depwait firstdata mma1 acc1wait epi1 | lasttile: mma_end acc_done epi_done | alldone exit   (us, medians over CTAs; relative to launch start)")
prev_end = None
for i in range(first, min(n, first + count)):
    c, tiles = info[i]
    s = st[i, :c]
    t0 = s[:, 0].min()
    lead = s[0::2]
    def med(ev, base=None, arr=s):
        v = arr[:, ev]
        ok = v > 0
        if not ok.any(): return float("nan")
        return float(np.median(v[ok] - t0)) / 1e3
    def mx(ev, arr=s):
        v = arr[:, ev]
        ok = v > 0
        if not ok.any(): return float("nan")
        return float((v[ok] - t0).max()) / 1e3
    gap = "" if prev_end is None else f" gap_prev_end->start {(t0 - prev_end) / 1e3:6.1f}"
    if raw:
        print(f"{i:3d} {c:4d} {tiles:5d} | " + " ".join(f"e{ev}:{med(ev):5.1f}/{mx(ev):5.1f}" for ev in range(1, 16)) + gap)
        prev_end = s[:, 11].max()
        continue
    print(f"{i:3d} {c:4d} {tiles:5d} | {mx(11):6.1f} | e1 {med(1):5.1f} e2 {med(2):5.1f} e3 {med(3, arr=lead):5.1f} e4 {med(4, arr=lead):5.1f} e6 {med(6):5.1f} e7 {med(7):5.1f} |"
          f" e5 {mx(5, arr=lead):5.1f} e8 {mx(8):5.1f} e9 {mx(9):5.1f} | e10 {mx(10):5.1f} e11 {mx(11):5.1f}{gap}")
    prev_end = s[:, 11].max()
