"""Oracle parity AT the BASELINE.json shapes (C1 640x480 pair, C2 832x832 pair, a 256-track slice of the C3 chunk): the
CUDA path through the C ABI vs oracle/loftr_oracle.py / oracle/multiview_oracle.py on identical seeded inputs.

Two weight sets: the BN-calibrated synthetic checkpoint (tests/weights.py; thousands of matches at the SHIPPED thr 0.2 /
temperature 0.1 on an overlapping-view pair: backbone tokens, post-transformer features, the full confidence matrix, the match
set) and the plain seeded weights at temperature 0.01 / thr 0 (dense tensors + the small match set).
Tolerances (north_star): confidences 1e-3, refined keypoints 0.1 px (asserted at 1e-2 px); match index sets identical except
where the oracle's own confidence sits within the 1e-3 tolerance of the threshold (the decision is then undefined at parity)."""
import os

import pytest
import torch

from oracle import loftr_oracle as lo
from oracle import multiview_oracle as mo
from tests import util, weights
from tests.test_refine_gpu import multiview_config, to_cuda

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-12)).item()


def _threads():
    torch.set_num_threads(max(1, min(os.cpu_count() or 1, 32)))


def _pair_parity(hw, sd, thr, temperature, images, min_matches):
    from detectorfreesfm_b200 import B200LoFTR
    _threads()
    im0, im1 = images
    ref = lo.loftr_forward({"image0": im0, "image1": im1}, sd, {"thr": thr, "temperature": temperature}, keep=True)
    m = B200LoFTR(util.loftr_config(thr=thr, temperature=temperature)).cuda().eval()
    m.load_state_dict(sd)
    # backbone tokens (x3_out + position encoding)
    h, w = ref["hw0_c"]
    assert (h, w) == (hw[0] // 8, hw[1] // 8)
    pe = lo.position_encoding_sine(256, h, w)[None]
    for im, key in ((im0, "backbone_c0"), (im1, "backbone_c1")):
        tok_ref = (ref[key] + pe).flatten(2).transpose(1, 2)[0]
        tok = m.extract_features(im.cuda()).cpu()
        assert rel_err(tok, tok_ref) < 5e-5, ("backbone", key, rel_err(tok, tok_ref))   # 14 split-fp16 conv layers: 2.3e-5 measured at 640x480
    data = {"image0": im0.cuda(), "image1": im1.cuda(), "_return_conf_matrix": True}
    m(data)
    # post-transformer features
    for k in ("feat_c0", "feat_c1"):
        e = rel_err(data[k].cpu(), ref[k])
        assert e < 5e-5, (k, e)
    # the full confidence matrix (L x S = 4800^2 / 10816^2)
    conf = data["conf_matrix"][0].cpu()
    cref = ref["conf_matrix"][0]
    assert conf.shape == cref.shape == (h * w, h * w)
    dconf = (conf - cref).abs().max().item()
    assert dconf < 1e-3, dconf
    # the match set
    L = h * w
    got = data["i_ids"].cpu() * L + data["j_ids"].cpu()
    exp = ref["i_ids"] * L + ref["j_ids"]
    assert len(exp) >= min_matches, len(exp)
    if not torch.equal(got, exp):
        sg, se = set(got.tolist()), set(exp.tolist())
        for key in sg ^ se:
            c = cref[key // L, key % L].item()
            assert abs(c - thr) < 1e-3, f"match ({key // L},{key % L}) differs and its oracle confidence {c} is not at the threshold"
        common = torch.tensor(sorted(sg & se))
    else:
        common = exp
    ci, cj = common // L, common % L
    assert (conf[ci, cj] - cref[ci, cj]).abs().max().item() < 1e-3
    if torch.equal(got, exp):
        assert (data["mconf"].cpu() - ref["mconf"]).abs().max().item() < 1e-3
        assert torch.equal(data["mkpts0_f"].cpu(), ref["mkpts0_f"]) and torch.equal(data["mkpts1_f"].cpu(), ref["mkpts1_f"])
    return len(exp)


@pytest.mark.parametrize("hw", [(480, 640), (832, 832)])
def test_c1_c2_pair_shipped_config_vs_oracle(hw):
    """BASELINE configs[0] / configs[1] shapes at the shipped thr 0.2 / temperature 0.1, BN-calibrated synthetic checkpoint, two
    overlapping views of one scene: O(10^3) matches."""
    sd = weights.loftr_state_dict(0, calibrated=True)
    images, _ = util.synth_scene(2, hw[0], hw[1], seed=40 + hw[0], noise=0.025, max_shift=64)
    n = _pair_parity(hw, sd, 0.2, 0.1, images, min_matches=500)
    print(f"{hw}: {n} matches")


@pytest.mark.parametrize("hw", [(480, 640), (832, 832)])
def test_c1_c2_pair_plain_weights_dense_tensors_vs_oracle(hw):
    """Same shapes with the plain seeded weights (full-strength random transformer) at temperature 0.01 / thr 0: the dense
    tensors are the check, the mutual-NN set is small."""
    sd = weights.loftr_state_dict(0)
    images = util.synth_pair(hw[0], hw[1], seed=3, shift=(16, 16))
    _pair_parity(hw, sd, 0.0, 0.01, images, min_matches=5)


@pytest.mark.parametrize("W,LW", [(15, 7), (11, 3)])
def test_c3_256_track_slice_vs_oracle(W, LW):
    """A 256-track slice of the C3 chunk (2000 tracks, 10 images of 600x800, <= 9 query views; both refinement iterations'
    window sizes) against oracle/multiview_oracle.py: <= 0.1 px is the bar, 1e-2 px asserted."""
    from detectorfreesfm_b200 import B200MultiviewMatcher
    _threads()
    sd = weights.multiview_state_dict(0)
    chunk = util.synth_chunk(M=2000, n_img=10, max_views=9, hw=(600, 800), seed=21, scales=torch.ones(1, 10, 2))
    sl = slice(640, 896)
    sub = dict(chunk)
    for k in ("query_points", "query_img_idxs", "query_movable_mask"):
        sub[k] = chunk[k][:, sl].contiguous()
    for k in ("reference_points_coarse", "track_valid_mask", "reference_img_idxs", "scales_relative", "view_point_vector"):
        sub[k] = chunk[k][:, :, sl].contiguous()
    ref = mo.multiview_forward(sub, sd, W, LW)
    m = B200MultiviewMatcher(multiview_config(W, LW), test=True).cuda().eval()
    m.load_state_dict(sd)
    d = to_cuda(sub)
    m(d)
    mask = sub["track_valid_mask"]
    q = d["query_points_refined"].cpu()
    r = d["reference_points_refined"][-1].cpu()
    s = d["std"][-1].cpu()
    assert (q - ref["query_points_refined"]).abs().max().item() < 1e-3
    assert (r - ref["reference_points_refined"])[mask].abs().max().item() < 1e-2
    assert (s - ref["std"])[mask].abs().max().item() < 1e-3
