"""Per-stage precision budget of the backbone convolutions (verdict item 6): emulate 1 / 2 / 3 fp16 MMA passes per layer group on the CPU oracle."""
import sys, time, itertools, torch
import torch.nn.functional as F
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import loftr_oracle as lo
from tests import util, weights
torch.set_num_threads(8)

def hi(x): return x.half().float()
def split(x):  # hi + lo (22 bits), what 3 passes see
    h = hi(x); return h + (x - h).half().float()

MODES = {
    3: (split, split),          # hh + hl + lh
    "2w": (split, hi),          # weights rounded to fp16 (drop a_hi*w_lo): passes hh + lh
    "2a": (hi, split),          # activations rounded to fp16 (drop a_lo*w_hi): passes hh + hl
    1: (hi, hi),
}
state = {"i": 0, "cfg": None}
orig_conv = lo._conv
def conv(x, w, stride, pad, q):
    i = state["i"]; state["i"] += 1
    qa, qw = MODES[state["cfg"](i, w)]
    return F.conv2d(qa(x), qw(w), None, stride, pad)
lo._conv = conv

def stage_of(i, w):
    # conv call order in resnet_fpn_8_2 (coarse only): 0 stem; layer1: 1..4; layer2: 5,6,(ds 7),8,9; layer3: 10,11,(ds 12),13,14; outconv 15
    if i == 0: return "stem"
    if i <= 4: return "l1"
    if i <= 9: return "l2"
    if i <= 14: return "l3"
    return "out"

def run(pair, sd, cfg):
    state["i"] = 0; state["cfg"] = cfg
    out = lo.loftr_forward({"image0": pair[0], "image1": pair[1]}, sd, {"temperature": 0.1, "thr": 0.2}, q=split, keep=True)
    return out

sd = weights.loftr_state_dict(0, calibrated=True)
H = int(sys.argv[1]) if len(sys.argv) > 1 else 480
W = int(sys.argv[2]) if len(sys.argv) > 2 else 640
pairs = [util.synth_scene(2, H, W, seed=40 + s, noise=0.025, max_shift=64)[0] for s in range(2)]
# fp32 reference (no quantiser anywhere)
lo._conv = orig_conv
refs = [lo.loftr_forward({"image0": p[0], "image1": p[1]}, sd, {"temperature": 0.1, "thr": 0.2}, keep=True) for p in pairs]
lo._conv = conv
configs = {
    "all 3-pass": lambda i, w: 3,
    "l1 2w": lambda i, w: "2w" if stage_of(i, w) == "l1" else 3,
    "l1 2a": lambda i, w: "2a" if stage_of(i, w) == "l1" else 3,
    "l1+l2 2w": lambda i, w: "2w" if stage_of(i, w) in ("l1", "l2") else 3,
    "l1+l2+l3 2w": lambda i, w: "2w" if stage_of(i, w) in ("l1", "l2", "l3") else 3,
    "all 2w": lambda i, w: "2w",
    "l1 1-pass": lambda i, w: 1 if stage_of(i, w) == "l1" else 3,
    "l1 conv1-of-block 1-pass": lambda i, w: 1 if i in (1, 3) else 3,
}
for name, cfg in configs.items():
    res = []
    for p, r in zip(pairs, refs):
        o = run(p, sd, cfg)
        dconf = (o["conf_matrix"] - r["conf_matrix"]).abs().max().item()
        dfeat = ((o["backbone_c0"] - r["backbone_c0"]).abs().max() / r["backbone_c0"].abs().max()).item()
        same = torch.equal(o["i_ids"], r["i_ids"]) and torch.equal(o["j_ids"], r["j_ids"])
        sg = set((o["i_ids"] * 100000 + o["j_ids"]).tolist()) ^ set((r["i_ids"] * 100000 + r["j_ids"]).tolist())
        res.append((dfeat, dconf, same, len(sg), len(r["i_ids"])))
    print(f"{name:28s} " + "  ".join(f"feat {a:.1e} conf {b:.1e} ids {'same' if c else 'diff(%d/%d)' % (d, n)}" for a, b, c, d, n in res), flush=True)
