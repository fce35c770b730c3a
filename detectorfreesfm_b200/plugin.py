"""Registration behind the reference's two hooks, without touching eval_dataset.py.

    import detectorfreesfm_b200.plugin as plugin; plugin.install()        # e.g. from sitecustomize / a conftest

* HP-1: wraps ``src.coarse_match.coarse_match_worker.build_model`` (coarse_match_worker.py:21-81) so that the matcher
  name ``loftr_b200`` (selected with ``neuralsfm.NEUSFM_coarse_matcher=loftr_b200``; the name contains 'loftr', hence the
  loader keeps ``df=8, pad_to=None``, coarse_match.py:82-84) returns ``(DetectorWrapper(), B200LoFTR)`` built from the
  same yacs config and checkpoint as ``loftr_official``.
* HP-2: wraps ``src.post_optimization.matcher_model.multiview_match_worker.build_model`` (multiview_match_worker.py:16-54)
  to return ``B200MultiviewMatcher`` with the per-iteration window rescale applied exactly as the reference does.

The reference tree is only imported inside ``install()``; nothing here runs at package import time.  INTEGRATION.md shows
the two-line alternative a maintainer would add to the reference instead of this monkey patch.
"""
import copy


def rescale_windows(matcher_cfg, rewindow_size_factor):
    """multiview_match_worker.py:20-34 (window 15 -> 11 -> 7 floor; left window 7 -> 3 floor)."""
    cfg = copy.deepcopy(matcher_cfg)
    if rewindow_size_factor is None:
        return cfg
    cur = cfg["multiview_transform"]["window_size"]
    w = ((cur // 2) - 1 * rewindow_size_factor) * 2 + 1
    w = max(w, 7)
    cfg["backbone"]["s2dnet"]["window_size"] = w
    cfg["multiview_transform"]["window_size"] = w
    cfg["multiview_matching_test"]["window_size"] = w
    lw = cfg["multiview_matching_test"]["left_point_movement_window_size"]
    if lw is not None:
        lw = max(((lw // 2) - 1 * rewindow_size_factor) * 2 + 1, 3)
        cfg["multiview_matching_test"]["left_point_movement_window_size"] = lw
    return cfg


def install(coarse=True, refine=True):
    import torch
    if coarse:
        from src.coarse_match import coarse_match as cm
        from src.coarse_match import coarse_match_worker as cmw
        from .coarse_matcher import B200LoFTR
        cm.cfgs["matcher"]["model"].setdefault("loftr_b200", dict(cm.cfgs["matcher"]["model"]["loftr_official"]))
        orig = cmw.build_model

        def build_model(args):
            if args["matcher"] != "loftr_b200":
                return orig(args)
            from third_party.LoFTR.src.config.default import get_cfg_defaults
            from src.utils.misc import lower_config
            margs = args["loftr_b200"]
            cfg = get_cfg_defaults()
            cfg.merge_from_file(margs[f"cfg_path_{args['type']}"])
            mcfg = lower_config(cfg)
            mcfg["loftr"]["match_coarse"]["thr"] = args["match_thr"]
            mcfg["loftr"]["coarse"]["temp_bug_fix"] = False
            matcher = B200LoFTR(config=mcfg["loftr"])
            matcher.load_state_dict(torch.load(margs["weight_path"], map_location="cpu")["state_dict"])
            return cmw.DetectorWrapper().eval(), matcher.eval()
        cmw.build_model = build_model
        cm.build_model = build_model   # coarse_match.py:11 star-imported the original name (match_worker itself resolves it in cmw)
    if refine:
        from omegaconf import OmegaConf
        from src.post_optimization.matcher_model import multiview_match as mm
        from src.post_optimization.matcher_model import multiview_match_worker as mmw
        from .refine_matcher import B200MultiviewMatcher

        def build_model(args, rewindow_size_factor=None, model_idx=None):
            cfg = OmegaConf.to_container(OmegaConf.load(args["cfg_path"][model_idx if model_idx is not None else 0]))
            mcfg = rescale_windows(cfg["model"]["multiview_refinement"], rewindow_size_factor)
            matcher = B200MultiviewMatcher(config=mcfg, test=True).eval()
            path = args["weight_path"][model_idx if model_idx is not None else 0]
            if path is not None:
                sd = torch.load(path, map_location="cpu")["state_dict"]
                matcher.load_state_dict({k: v for k, v in sd.items() if "matcher." in k})
            return matcher
        mmw.build_model = build_model
        # multiview_match.py:7 does ``from .multiview_match_worker import *``: ``multiview_matcher`` (:11) resolves ``build_model`` in
        # ITS module globals, so the name must be rebound there too or the reference MultiviewMatcher keeps running silently.
        mm.build_model = build_model
