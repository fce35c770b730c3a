"""The oracle restatement against the committed golden fixtures (outputs of the reference itself, see
tests/golden/make_golden.py) and the reference's own RoIAlign known-answer vector.  CPU only."""
import os

import torch

from oracle import build_native
from oracle import loftr_oracle as lo
from oracle import multiview_oracle as mo
from tests import weights
from tests import util

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_roialign_readme_vector():
    """third_party/RoIAlign.pytorch/README.md:42-96 -- the only hard golden vector the reference holds on the hot path."""
    g = torch.load(os.path.join(GOLD, "roialign_readme.pt"))
    out = build_native.roialign_forward(g["image"], g["boxes_norm"], g["box_index"], 4, 4)
    assert torch.equal(out, g["crops"])
    assert out[1, 0, 3].abs().max().item() == 0  # the row that falls outside the image is extrapolated with 0
    assert abs(out[1, 0, 0, 1].item() - 25.375) < 1e-4


def test_roialign_edge_cases():
    img = torch.arange(0., 2 * 3 * 5 * 6).view(2, 3, 5, 6)
    # empty box list
    assert build_native.roialign_forward(img, torch.zeros(0, 4), torch.zeros(0, dtype=torch.int32), 3, 3).shape == (0, 3, 3, 3)
    # crop size 1 samples the box centre
    out = build_native.roialign_forward(img, torch.tensor([[0., 0., 1., 1.]]), torch.tensor([1], dtype=torch.int32), 1, 1)
    assert torch.allclose(out[0, :, 0, 0], img[1, :, 2, :].mean(-1) * 0 + (img[1, :, 2, 2] + img[1, :, 2, 3]) / 2)
    # a box fully outside the image -> extrapolation value
    out = build_native.roialign_forward(img, torch.tensor([[2., 2., 3., 3.]]), torch.tensor([0], dtype=torch.int32), 2, 2, 7.0)
    assert (out == 7.0).all()


def _digest_close(t, d, tol):
    f = t.flatten().double()
    idx = torch.linspace(0, f.numel() - 1, 257).long()
    assert list(t.shape) == d["shape"]
    assert (f[idx].float() - d["samples"]).abs().max().item() <= tol
    assert abs(f.sum().item() - d["sum"]) <= tol * f.numel() ** 0.5 + 1e-6 * abs(d["sum"])


def test_loftr_oracle_vs_golden():
    gold = torch.load(os.path.join(GOLD, "loftr_small.pt"))
    sd = weights.loftr_state_dict(0)
    im0, im1 = util.synth_pair(64, 80, seed=1)
    for name, g in gold.items():
        out = lo.loftr_forward({"image0": im0, "image1": im1, "scale0": torch.tensor([[1.5, 1.25]]), "scale1": torch.tensor([[1.0, 2.0]])},
                               sd, {"thr": g["thr"], "temperature": g["temperature"]}, keep=True)
        _digest_close(out["conf_matrix"], g["conf"], 1e-6)
        assert torch.equal(out["i_ids"], g["i_ids"]) and torch.equal(out["j_ids"], g["j_ids"]), name
        if len(g["mconf"]):
            assert (out["mconf"] - g["mconf"]).abs().max().item() < 1e-6
            assert torch.equal(out["mkpts0_f"], g["mkpts0_f"]) and torch.equal(out["mkpts1_f"], g["mkpts1_f"])


def test_multiview_oracle_vs_golden():
    gold = torch.load(os.path.join(GOLD, "multiview_small.pt"))
    sd = weights.multiview_state_dict(0)
    for name, g in gold.items():
        data = util.synth_chunk(M=24, n_img=4, max_views=3, seed=4)
        out = mo.multiview_forward(data, sd, g["W"], g["LW"])
        mask = data["track_valid_mask"]
        assert (out["query_points_refined"] - g["query_points_refined"]).abs().max().item() < 1e-4, name
        assert (out["reference_points_refined"] - g["reference_points_refined"])[mask].abs().max().item() < 1e-3, name
        assert (out["std"] - g["std"])[mask].abs().max().item() < 1e-4, name
        assert out["reference_points_refined"][~mask].abs().max().item() == 0


def test_mask_border_quirk_and_position_encoding():
    """Appendix A: mask_border only removes the LEADING rows/cols; PE uses div_term = exp(-2k) with 1-based positions."""
    conf = torch.full((1, 36, 36), 0.9)  # 6x6 grids, every entry above threshold
    conf += torch.arange(36 * 36).view(1, 36, 36) * 1e-6
    r = lo.get_coarse_match(conf, (6, 6), (6, 6), (48, 48), 0.2, 2)
    # the global row/col maxima sit at the last row/col -> the trailing border is NOT masked
    assert len(r["i_ids"]) == 1 and int(r["i_ids"][0]) == 35 and int(r["j_ids"][0]) == 35
    pe = lo.position_encoding_sine(256, 4, 5)
    k = 3
    assert abs(pe[4 * k, 0, 1].item() - torch.sin(torch.tensor(2.0) * torch.exp(torch.tensor(-2.0 * k))).item()) < 1e-6
    assert abs(pe[4 * k + 3, 2, 0].item() - torch.cos(torch.tensor(3.0) * torch.exp(torch.tensor(-2.0 * k))).item()) < 1e-6


def test_empty_match_set():
    sd = weights.loftr_state_dict(0)
    im0, im1 = util.synth_pair(32, 32, seed=2)
    out = lo.loftr_forward({"image0": im0, "image1": im1}, sd, {"thr": 1.0, "temperature": 0.1})
    assert out["mkpts0_f"].shape == (0, 2) and out["mconf"].shape == (0,)


def test_postprocess_oracle_vs_golden():
    """oracle/postprocess_oracle.py reproduces the reference's own outputs stored in tests/golden/postprocess_small.pt."""
    import numpy as np
    from oracle import postprocess_oracle as po
    g = torch.load(os.path.join(GOLD, "postprocess_small.pt"), weights_only=False)
    fk, fs, upd = po.merge_keypoints(g["matches"], g["names"], " ")
    n_kp = 0
    for name in g["names"]:
        assert np.array_equal(fk[name], g["final_keypoints"][name]) and fk[name].dtype == g["final_keypoints"][name].dtype
        assert np.array_equal(fs[name], g["final_scores"][name]) and fs[name].dtype == np.float32
        n_kp += fk[name].shape[0]
    for k, v in g["updated_matches"].items():
        assert np.array_equal(upd[k], v) and upd[k].shape == v.shape
    assert n_kp > 100 and g["final_keypoints"][g["names"][4]].shape == (0, 2)


def test_image_oracle_vs_golden_and_pillow_restatement():
    """read_grayscale outputs stored from the reference; the numpy restatement of Pillow's 8-bit resampler equals PIL itself."""
    import numpy as np
    from PIL import Image
    from oracle import image_oracle as io
    for g in torch.load(os.path.join(GOLD, "image_small.pt"), weights_only=False):
        t, s, hw = io.read_grayscale_from_array(g["image"].numpy(), g["resize"], df=g["df"])
        assert torch.equal(t, g["tensor"]) and torch.equal(s, g["scales"]) and torch.equal(hw, g["original_hw"])
    for (H, W, oh, ow) in [(120, 160, 90, 120), (100, 37, 64, 24), (33, 50, 99, 120), (64, 64, 64, 32), (17, 17, 17, 17), (5, 7, 1, 1),
                           (700, 900, 208, 264)]:
        img = util.synth_photo(H, W, seed=H + W)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.LANCZOS))
        assert np.array_equal(ref, io.resample_8bpc(img, ow, oh)), (H, W, oh, ow)


def test_host_lanczos_tables_equal_oracle_tables():
    """detectorfreesfm_b200.image_pipeline builds the same fixed-point tables (and the same sizes) as the oracle restatement."""
    import numpy as np
    from detectorfreesfm_b200 import image_pipeline as ip
    from oracle import image_oracle as io
    for a, b in [(640, 480), (480, 640), (4000, 832), (37, 24), (50, 120), (7, 1), (832, 832)]:
        b1, k1 = io.coeffs(a, b)
        b2, k2 = ip.lanczos_coeffs(a, b)
        assert np.array_equal(b1, b2) and np.array_equal(k1, k2)
    for args in [(800, 600, (512,), 8, False), (600, 800, (1200,), 8, False), (640, 480, (-1,), 8, False), (640, 480, (320, 200), None, False),
                 (300, 200, (1200,), 8, True)]:
        assert ip.process_resize(*args) == io.process_resize(*args)


def test_refine_worker_loop_vs_golden():
    """Row b1 host loop (refine_stage.match_worker) vs the [K,4] arrays the reference's matchWorker produced for the same chunks
    and stand-in matcher; with freeze=True (the evident intent of UpdatedQueryPts) later chunks hold frozen nodes."""
    import numpy as np
    from detectorfreesfm_b200 import refine_stage as rs
    from tests import util
    ref = torch.load(os.path.join(GOLD, "refine_worker_small.pt"), weights_only=False)["results"]
    loader = torch.utils.data.DataLoader(util.worker_chunks(), num_workers=0)
    got = rs.match_worker(loader, util.StandInRefiner(), range(4), device=torch.device("cpu"))
    assert len(got) == len(ref) == 3 and all(np.array_equal(a, b) for a, b in zip(ref, got))
    frozen = rs.match_worker(torch.utils.data.DataLoader(util.worker_chunks(), num_workers=0), util.StandInRefiner(), range(4),
                             device=torch.device("cpu"), freeze=True)
    assert frozen[0].shape == got[0].shape and frozen[1].shape[0] < got[1].shape[0]
