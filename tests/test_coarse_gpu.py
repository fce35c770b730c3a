"""HP-1 parity on the GPU: B200LoFTR (C ABI, sm_100a kernels) vs the CPU oracle (oracle/loftr_oracle.py) on identical
seeded weights and inputs.  Tolerances: north_star -- match confidences within 1e-3; index sets identical."""
import pytest
import torch

from oracle import loftr_oracle as lo
from tests import weights
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sd():
    return weights.loftr_state_dict(0)


@pytest.fixture(scope="module")
def matcher(sd):
    from detectorfreesfm_b200 import B200LoFTR
    m = B200LoFTR(util.loftr_config(thr=0.2, temperature=0.1)).cuda().eval()
    m.load_state_dict(sd)
    return m


def rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-12)).item()


@pytest.mark.parametrize("hw", [(96, 128), (64, 64), (120, 88)])
def test_backbone_features(matcher, sd, hw):
    im = util.synth_image(hw[0], hw[1], seed=3)
    x3, _ = lo.resnet_fpn_8_2(im, sd, fine=False)
    h, w = x3.shape[2:]
    ref = (x3 + lo.position_encoding_sine(256, h, w)[None]).flatten(2).transpose(1, 2)[0]
    out = matcher.extract_features(im.cuda()).cpu()
    assert out.shape == ref.shape
    assert rel_err(out, ref) < 2e-5, rel_err(out, ref)


def test_transformer(matcher, sd):
    g = torch.Generator().manual_seed(5)
    f0 = torch.randn(1, 300, 256, generator=g)
    f1 = torch.randn(1, 417, 256, generator=g)
    r0, r1 = lo.local_feature_transformer(f0, f1, sd, "loftr_coarse", ["self", "cross"] * 4, 8)
    o0, o1 = matcher.transform(f0[0].cuda().clone(), f1[0].cuda().clone())
    assert rel_err(o0.cpu(), r0[0]) < 5e-5, rel_err(o0.cpu(), r0[0])
    assert rel_err(o1.cpu(), r1[0]) < 5e-5, rel_err(o1.cpu(), r1[0])


@pytest.mark.parametrize("hw0,hw1,thr", [((20, 24), (20, 24), 0.2), ((24, 30), (18, 26), 0.2), ((16, 16), (16, 16), 0.0),
                                         ((40, 52), (44, 48), 0.2)])
def test_coarse_matching_stage(matcher, hw0, hw1, thr):
    """dual-softmax + mutual-NN on synthetic discriminative features (thousands of candidate matches)."""
    L, S = hw0[0] * hw0[1], hw1[0] * hw1[1]
    f0, f1 = util.discriminative_features(L, S, seed=L + S)
    conf = lo.dual_softmax_conf(f0[None], f1[None], 0.1)
    ref = lo.get_coarse_match(conf, hw0, hw1, (hw0[0] * 8, hw0[1] * 8), thr, 2)
    matcher.thr = thr
    try:
        i_ids, j_ids, mconf, cm = matcher.coarse_match(f0.cuda(), hw0, f1.cuda(), hw1, return_conf=True)
    finally:
        matcher.thr = 0.2
    assert (cm.cpu() - conf[0]).abs().max().item() < 1e-3
    assert len(ref["i_ids"]) > 10
    assert torch.equal(i_ids.cpu(), ref["i_ids"]) and torch.equal(j_ids.cpu(), ref["j_ids"])
    assert (mconf.cpu() - ref["mconf"]).abs().max().item() < 1e-3


@pytest.mark.parametrize("temperature,thr", [(0.1, 0.2), (0.01, 0.2), (0.01, 0.0)])
def test_end_to_end_pair(sd, temperature, thr):
    from detectorfreesfm_b200 import B200LoFTR
    m = B200LoFTR(util.loftr_config(thr=thr, temperature=temperature)).cuda().eval()
    m.load_state_dict(sd)
    im0, im1 = util.synth_pair(96, 128, seed=1)
    scale0, scale1 = torch.tensor([[1.5, 1.25]]), torch.tensor([[1.0, 2.0]])
    ref = lo.loftr_forward({"image0": im0, "image1": im1, "scale0": scale0, "scale1": scale1}, sd,
                           {"thr": thr, "temperature": temperature}, keep=True)
    data = {"image0": im0.cuda(), "image1": im1.cuda(), "scale0": scale0.cuda(), "scale1": scale1.cuda(),
            "_return_conf_matrix": True}
    m(data)
    assert (data["conf_matrix"].cpu() - ref["conf_matrix"]).abs().max().item() < 1e-3
    assert torch.equal(data["i_ids"].cpu(), ref["i_ids"]) and torch.equal(data["j_ids"].cpu(), ref["j_ids"])
    if len(ref["mconf"]):
        assert (data["mconf"].cpu() - ref["mconf"]).abs().max().item() < 1e-3
        assert torch.equal(data["mkpts0_f"].cpu(), ref["mkpts0_f"]) and torch.equal(data["mkpts1_f"].cpu(), ref["mkpts1_f"])
    assert (data["m_bids"] == 0).all()


# ------------------------------------------------------------------------------------------- fine stage (row a11)
@pytest.mark.parametrize("hw", [(96, 128), (64, 80)])
def test_fine_feature_map(sd, hw):
    """FPN top-down path: x1_out of ResNetFPN_8_2 (1/2 resolution, 128 ch) vs the oracle."""
    from detectorfreesfm_b200 import B200LoFTR
    m = B200LoFTR(util.loftr_config(fine=True)).cuda().eval()
    m.load_state_dict(sd)
    im = util.synth_image(hw[0], hw[1], seed=7)
    x3, x1 = lo.resnet_fpn_8_2(im, sd, fine=True)
    tokens, feat_f = m.extract_features(im.cuda())
    ref_f = x1[0].flatten(1).t()
    assert feat_f.shape == ref_f.shape
    assert rel_err(feat_f.cpu(), ref_f) < 5e-5, rel_err(feat_f.cpu(), ref_f)
    h, w = x3.shape[2:]
    ref_c = (x3 + lo.position_encoding_sine(256, h, w)[None]).flatten(2).transpose(1, 2)[0]
    assert rel_err(tokens.cpu(), ref_c) < 5e-5


@pytest.mark.parametrize("temperature,thr", [(0.01, 0.2), (0.01, 0.0)])
def test_end_to_end_coarse_fine(sd, temperature, thr):
    """match type 'coarse_fine' (texturepoor config): fine-level keypoints within 0.1 px (north_star), expec_f within 1e-3."""
    from detectorfreesfm_b200 import B200LoFTR
    m = B200LoFTR(util.loftr_config(thr=thr, temperature=temperature, fine=True)).cuda().eval()
    m.load_state_dict(sd)
    im0, im1 = util.synth_pair(96, 128, seed=1)
    scale0, scale1 = torch.tensor([[1.5, 1.25]]), torch.tensor([[1.0, 2.0]])
    ref = lo.loftr_forward({"image0": im0, "image1": im1, "scale0": scale0, "scale1": scale1}, sd,
                           {"thr": thr, "temperature": temperature, "fine_enable": True}, keep=True)
    data = {"image0": im0.cuda(), "image1": im1.cuda(), "scale0": scale0.cuda(), "scale1": scale1.cuda()}
    m(data)
    assert len(ref["mconf"]) > 0
    assert torch.equal(data["i_ids"].cpu(), ref["i_ids"]) and torch.equal(data["j_ids"].cpu(), ref["j_ids"])
    assert (data["expec_f"].cpu() - ref["expec_f"]).abs().max().item() < 1e-3
    assert torch.equal(data["mkpts0_f"].cpu(), ref["mkpts0_f"])
    assert (data["mkpts1_f"].cpu() - ref["mkpts1_f"]).abs().max().item() < 0.1
    assert (data["mkpts1_f"].cpu() - ref["mkpts1_f"]).abs().max().item() < 1e-2  # in practice ~1e-4 px


# ----------------------------------------------------------------------------------------------- edge cases
def test_unequal_image_sizes_and_empty_match_set(sd):
    """Different HxW per image (every demo image has its own size, loftr.py:48-49) and a threshold nothing passes."""
    from detectorfreesfm_b200 import B200LoFTR
    im0 = util.synth_image(96, 136, seed=11)
    im1 = util.synth_image(120, 88, seed=12)
    for thr, temp in ((0.0, 0.01), (1.0, 0.1)):
        m = B200LoFTR(util.loftr_config(thr=thr, temperature=temp)).cuda().eval()
        m.load_state_dict(sd)
        ref = lo.loftr_forward({"image0": im0, "image1": im1}, sd, {"thr": thr, "temperature": temp}, keep=True)
        data = {"image0": im0.cuda(), "image1": im1.cuda(), "_return_conf_matrix": True}
        m(data)
        assert tuple(data["hw0_c"]) == (12, 17) and tuple(data["hw1_c"]) == (15, 11)
        assert (data["conf_matrix"].cpu() - ref["conf_matrix"]).abs().max().item() < 1e-3
        assert torch.equal(data["i_ids"].cpu(), ref["i_ids"]) and torch.equal(data["j_ids"].cpu(), ref["j_ids"])
        assert data["mkpts0_f"].shape == ref["mkpts0_f"].shape and data["mconf"].shape == ref["mconf"].shape
        if thr == 1.0:
            assert data["mkpts0_f"].shape == (0, 2) and data["m_bids"].shape == (0,)
        else:
            assert torch.equal(data["mkpts1_f"].cpu(), ref["mkpts1_f"])


@pytest.mark.parametrize("env", [{"DFSFM_ATTN_FOLD": "0"}, {"DFSFM_KV_EPI": "0"}, {"DFSFM_ENC_FUSED": "0"}, {}])
def test_transformer_schedules_agree_with_oracle(env):
    """The A/B schedules of the coarse transformer (round-1 GEMM-per-linear + attn_apply; folded attention with the fp32-state
    reduction; folded + KvEpi with four GEMM launches; the fused kernel) all meet the same parity bar.  The switches are read once per
    process, hence a subprocess per schedule (tests/check_transformer.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "check_transformer.py")], env=e, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-2000:]


def test_tile_stealing_schedule_meets_the_same_parity_bar():
    """DFSFM_TILE_STEAL=1 (cluster-launch-control work stealing in the engine-2 GEMM, off by default) on the C1 shape, where the conv and
    similarity launches have 4-25 tiles per CTA pair: the BASELINE-shape oracle test in a subprocess (the switch is read once per process)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e["DFSFM_TILE_STEAL"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_baseline_shapes_gpu.py"), "-m", "gpu", "-q", "-x",
                        "-k", "shipped_config and hw0"], env=e, cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
