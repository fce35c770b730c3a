import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import loftr_oracle as lo
from tests import weights
from tests import util
sd = weights.loftr_state_dict(0)
im0, im1 = util.synth_image(832, 832, 1000), util.synth_image(832, 832, 1001)
for n in (16, 32, 64, 128):
    if n > (os.cpu_count() or 1): break
    torch.set_num_threads(n)
    t = time.perf_counter(); lo.loftr_forward({"image0": im0, "image1": im1}, sd); dt = time.perf_counter() - t
    print(n, "threads:", round(dt, 2), "s/pair")
