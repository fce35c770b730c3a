"""Pin the oracle restatement against the UNMODIFIED reference modules imported from /root/reference (build container
only; skipped on the GPU box where that tree does not exist -- the committed golden fixtures cover it there)."""
import pytest
import torch

from tests import util  # noqa: E402

from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.available(), reason="/root/reference not present")


def test_loftr_coarse_and_fine_match_reference():
    from oracle import loftr_oracle as lo
    from tests import weights
    from tests import util
    LoFTR, _ = ref_shims.import_loftr()
    sd = weights.loftr_state_dict(0)
    for fine in (False, True):
        m = LoFTR(ref_shims.loftr_config(thr=0.0, fine=fine, temperature=0.01)).eval()
        m.load_state_dict(sd, strict=True)
        im0, im1 = util.synth_pair(64, 96, seed=3)
        d = {"image0": im0, "image1": im1, "scale0": torch.tensor([[1.5, 1.25]]), "scale1": torch.tensor([[1.0, 2.0]])}
        with torch.no_grad():
            m(d)
        out = lo.loftr_forward({k: d[k] for k in ("image0", "image1", "scale0", "scale1")}, sd,
                               {"thr": 0.0, "temperature": 0.01, "fine_enable": fine}, keep=True)
        assert (d["conf_matrix"] - out["conf_matrix"]).abs().max().item() < 1e-6
        assert torch.equal(d["i_ids"], out["i_ids"]) and torch.equal(d["j_ids"], out["j_ids"])
        assert (d["mconf"] - out["mconf"]).abs().max().item() < 1e-6
        assert (d["mkpts0_f"] - out["mkpts0_f"]).abs().max().item() < 1e-3
        assert (d["mkpts1_f"] - out["mkpts1_f"]).abs().max().item() < 1e-3


def test_multiview_matches_reference():
    from oracle import multiview_oracle as mo
    from tests import weights
    from tests import util
    MM = ref_shims.import_multiview()
    sd = weights.multiview_state_dict(0)
    m = MM(config=ref_shims.multiview_config(15, 7), test=True).eval()
    m.load_state_dict(sd, strict=True)
    data = util.synth_chunk(M=20, n_img=4, max_views=3, seed=9)
    d2 = {k: (v.clone() if torch.is_tensor(v) else list(v)) for k, v in data.items()}
    with torch.no_grad():
        m(d2)
    out = mo.multiview_forward(data, sd, 15, 7)
    mask = data["track_valid_mask"]
    assert (d2["query_points_refined"] - out["query_points_refined"]).abs().max().item() < 1e-4
    assert (d2["reference_points_refined"][-1] - out["reference_points_refined"])[mask].abs().max().item() < 1e-3
    assert (d2["std"][-1] - out["std"])[mask].abs().max().item() < 1e-4


def test_c_roialign_matches_reference_cpp():
    from oracle import build_native
    ext = ref_shims.build_ref_roialign()
    g = torch.Generator().manual_seed(0)
    image = torch.rand(2, 3, 40, 56, generator=g)
    nb = torch.rand(64, 4, generator=g) * 1.4 - 0.2
    nb[:, 2:] = nb[:, :2] + torch.rand(64, 2, generator=g) * 0.6
    bi = torch.randint(0, 2, (64,), generator=g, dtype=torch.int32)
    crops = torch.zeros(1)
    ext.forward(image, nb.contiguous(), bi, 0.0, 35, 35, crops)
    assert torch.equal(build_native.roialign_forward(image, nb, bi, 35, 35), crops)


def test_postprocess_oracle_matches_reference():
    """Match2Kpts / keypoint_worker / update_matches / transform_keypoints themselves vs oracle/postprocess_oracle.py: equal
    arrays, dtypes and shapes, including pairs without matches and an image that never appears."""
    import itertools
    import numpy as np
    from oracle import postprocess_oracle as po
    from tests.golden.make_golden import reference_postprocess
    for seed, (n, m) in enumerate([(4, 50), (6, 300), (5, [0, 10, 200]), (3, 1)]):
        pairs = list(itertools.combinations(range(n), 2))
        if seed == 2:
            pairs = [p for p in pairs if 4 not in p]
        matches, names = util.synth_matches(n, pairs, m, seed=seed)
        ref = reference_postprocess(matches, names)
        ora = po.merge_keypoints(matches, names, " ")
        for name in names:
            for a, b in ((ref[0][name], ora[0][name]), (ref[1][name], ora[1][name])):
                assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
        for k in matches:
            assert ref[2][k].dtype == ora[2][k].dtype and ref[2][k].shape == ora[2][k].shape and np.array_equal(ref[2][k], ora[2][k])


def test_image_oracle_matches_reference_read_grayscale(tmp_path):
    """the reference's read_grayscale (cv2 decode of a PNG, PIL LANCZOS, /255) vs oracle/image_oracle.py: equal tensors, scales."""
    from oracle import image_oracle as io
    from tests.golden.make_golden import reference_read_grayscale
    for seed, (h, w, resize, df) in enumerate([(150, 200, (96,), 8), (97, 61, (128,), 8), (64, 80, None, None), (300, 200, (64, 48), None)]):
        img = util.synth_photo(h, w, seed)
        t, scales, hw = reference_read_grayscale(img, resize, df, str(tmp_path))
        to, so, ho = io.read_grayscale_from_array(img, resize, df=df)
        assert t.dtype == to.dtype and torch.equal(t, to) and torch.equal(scales, so) and torch.equal(hw, ho)


def test_refine_worker_loop_matches_reference_match_worker():
    """Row b1: the reference's matchWorker (its real code, with the chunk dataset replaced by a list and dict_to_cuda by the
    identity) and detectorfreesfm_b200.refine_stage.match_worker give identical [K,4] arrays for the same chunks and the same
    (deterministic stand-in) matcher -- including the fact that UpdatedQueryPts never freezes anything in the reference."""
    import numpy as np
    from detectorfreesfm_b200 import refine_stage as rs
    from tests import util
    from tests.golden.make_golden import reference_refine_worker
    ref = reference_refine_worker()
    got = rs.match_worker(torch.utils.data.DataLoader(util.worker_chunks(), num_workers=0), util.StandInRefiner(), range(4),
                          device=torch.device("cpu"))
    assert len(ref) == len(got) == 3
    for a, b in zip(ref, got):
        assert a.shape == b.shape and a.shape[1] == 4 and np.array_equal(a, b)


def test_plugin_install_rebinds_both_hooks_and_builds_b200_models(tmp_path, monkeypatch):
    """plugin.install() behind the reference's own hook modules (imported where they lie, third-party deps stubbed):
    * the HP-1 name 'loftr_b200' builds (DetectorWrapper, B200LoFTR) from the reference's yacs config + a checkpoint file, with
      thr / temp_bug_fix overwritten like coarse_match_worker.py:31-35, and other matcher names still reach the original;
    * the HP-2 builder is rebound in BOTH modules that hold the name (multiview_match.py star-imports it, :7) and applies the
      per-iteration window rescale of multiview_match_worker.py:20-34."""
    import os
    from detectorfreesfm_b200 import B200LoFTR, B200MultiviewMatcher
    from detectorfreesfm_b200 import plugin
    from tests import weights
    cm, cmw, mm, mmw = ref_shims.import_hook_modules()
    orig_coarse, orig_refine = cmw.build_model, mmw.build_model
    assert mm.build_model is orig_refine                      # the star import the advisor pointed at
    plugin.install()
    assert cmw.build_model is not orig_coarse and cm.build_model is cmw.build_model
    assert mmw.build_model is not orig_refine and mm.build_model is mmw.build_model
    assert "loftr_b200" in cm.cfgs["matcher"]["model"]
    # ---- HP-1
    ckpt = tmp_path / "outdoor_ds.ckpt"
    torch.save({"state_dict": {"matcher." + k: v for k, v in weights.loftr_state_dict(0).items()}}, ckpt)
    monkeypatch.chdir(ref_shims.REF)                          # the yacs .py configs use cwd-relative paths, like eval_dataset.py
    args = dict(cm.cfgs["matcher"]["model"])
    args.update({"matcher": "loftr_b200", "type": "coarse_only", "match_thr": 0.35})
    args["loftr_b200"] = {**args["loftr_b200"], "weight_path": str(ckpt)}
    detector, matcher = cmw.build_model(args)
    assert isinstance(detector, cmw.DetectorWrapper) and isinstance(matcher, B200LoFTR)
    assert matcher.thr == 0.35 and matcher.fine is False and matcher.temperature == 0.1 and matcher._packed is not None
    args["type"] = "coarse_fine"
    _, matcher_f = cmw.build_model(args)
    assert matcher_f.fine is True
    with pytest.raises(NotImplementedError):                  # unknown names still fall through to the reference's if/elif chain
        cmw.build_model({**args, "matcher": "no_such_matcher", "seed": 0})
    # ---- HP-2
    ckpt2 = tmp_path / "multiview_matcher.ckpt"
    sdm = {"matcher." + k.replace("fine_transformer", "loftr_fine"): v for k, v in weights.multiview_state_dict(0).items()}
    sdm["matcher.loftr_coarse.layers.0.q_proj.weight"] = torch.zeros(4, 4)   # dropped by the reference (:47-49) and by the packer
    sdm["loss.whatever"] = torch.zeros(1)
    torch.save({"state_dict": sdm}, ckpt2)
    yaml_path = os.path.join(ref_shims.REF, "hydra_training_configs", "experiment", "multiview_refinement_matching.yaml")
    margs = {"cfg_path": [yaml_path], "weight_path": [str(ckpt2)], "seed": 666}
    for factor, (w, lw) in {None: (15, 7), 1: (13, 5), 2: (11, 3), 5: (7, 3)}.items():
        m = mm.build_model(margs, rewindow_size_factor=factor, model_idx=0)
        assert isinstance(m, B200MultiviewMatcher) and (m.W, m.LW) == (w, lw) and m._packed is not None
