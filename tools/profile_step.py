"""One HP-1 pair (832x832) + one small HP-2 chunk, for ncu captures.  Usage: python tools/profile_step.py [pairs] [tracks]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from detectorfreesfm_b200 import B200LoFTR, B200MultiviewMatcher
from tests import weights
from tests import util
from tests.util import multiview_config, to_cuda

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tracks = int(sys.argv[2]) if len(sys.argv) > 2 else 0
m = B200LoFTR(util.loftr_config()).cuda().eval()
m.load_state_dict(weights.loftr_state_dict(0, calibrated=True))
ims = [im.cuda() for im in util.synth_scene(2, 832, 832, 1000, noise=0.025)[0]]
for _ in range(n_pairs):
    d = {"image0": ims[0], "image1": ims[1]}
    m(d)
torch.cuda.synchronize()
print("matches", len(d["mconf"]))
if tracks:
    rm = B200MultiviewMatcher(multiview_config(15, 7), test=True).cuda().eval()
    rm.load_state_dict(weights.multiview_state_dict(0))
    chunk = to_cuda(util.synth_chunk(M=tracks, n_img=10, max_views=9, hw=(600, 800), seed=11, scales=torch.ones(1, 10, 2)))
    for _ in range(2):
        rm(dict(chunk))
    torch.cuda.synchronize()
print("done")
