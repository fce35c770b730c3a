"""The coarse-matching stage end to end on the GPU: rows a1-a2 of SURVEY.md section 8 around HP-1, plus 8(f) rows 1 and 3.

Mirrors ``match_worker`` (src/coarse_match/coarse_match_worker.py:103-145: dataset -> per-pair ``detector(data); matcher(data)``
-> ``(M,5)`` arrays keyed ``"path0<split>path1"``, optional grid rounding :133-136) and the part of
``detector_free_coarse_matching`` between the matcher and the h5 files (src/coarse_match/coarse_match.py:190-237), with

  * the image pipeline of image_pipeline.py (every image decoded and resized once, device resident),
  * any matcher honouring the HP-1 contract (``matcher(data)`` adds mkpts0_f / mkpts1_f / mconf / m_bids; B200LoFTR here),
  * the post-processing of postprocess.py on the match arrays while they are still on the device.

Nothing leaves the GPU between the decoded uint8 image and the final keypoint / match-index arrays except the dict
bookkeeping; the outputs have the reference's types, so ``save_h5`` (coarse_match.py:239-256) applies unchanged.
"""
import torch

from .image_pipeline import B200CoarseMatchingDataset
from .postprocess import KeypointMerger


@torch.no_grad()
def match_worker(subset_ids, image_lists, covis_pairs_out, cfgs, matcher, detector=None, keep_on_device=False, dataset=None):
    """coarse_match_worker.py:103-145 with a ready-built ``matcher`` (build_model is the caller's, see plugin.py).

    -> ``{pair_key: (M,5)}`` of [x0, y0, x1, y1, conf]: numpy float32 arrays like the reference, or CUDA tensors when
    ``keep_on_device`` (what ``merge_keypoints`` consumes without a host round trip)."""
    args = cfgs["matcher"]
    if dataset is None:
        dataset = B200CoarseMatchingDataset(cfgs["data"], image_lists, covis_pairs_out, subset_ids)
    loader = torch.utils.data.DataLoader(dataset, num_workers=0)      # items are CUDA tensors (no worker processes)
    # :134 tests ``args['model']['type'] is not 'coarse_only'`` -- an IDENTITY test against a literal, true for every run-time
    # string (hydra hands over a fresh str object), so the reference rounds whenever a ratio is configured, whatever the match
    # type.  The shipped coarse_only configs set the ratio to null; coarse points are multiples of 8 px anyway.
    rounding = args.get("round_matches_ratio") is not None
    matches = {}
    for data in loader:
        f_name0, f_name1 = data["pair_key"][0][0], data["pair_key"][1][0]
        data_c = {k: v.cuda() if isinstance(v, torch.Tensor) else v for k, v in data.items()}
        if detector is not None:
            detector(data_c)
        matcher(data_c)
        assert bool((data_c["m_bids"] == 0).all())                    # extract_preds: bs == 1
        mkpts0, mkpts1, mconf = data_c["mkpts0_f"], data_c["mkpts1_f"], data_c["mconf"]
        if rounding:                                                  # round to the grid for the later feature tracks
            r = args["round_matches_ratio"]
            s0, s1 = data_c["scale0"][:, [1, 0]], data_c["scale1"][:, [1, 0]]
            mkpts0 = torch.round((mkpts0 / s0) / r) * r * s0
            mkpts1 = torch.round((mkpts1 / s1) / r) * r * s1
        m = torch.cat([mkpts0, mkpts1, mconf[:, None]], -1).float()   # (N, 5)
        matches[args["pair_name_split"].join([f_name0, f_name1])] = m if keep_on_device else m.cpu().numpy()
    return matches


def coarse_matching_stage(image_lists, covis_pairs, cfgs, matcher, detector=None, merger=None):
    """coarse_match.py:190-237 without Ray and without the h5 cache: all pairs -> (final_keypoints, final_scores,
    updated_matches, raw matches).  ``covis_pairs``: list of "path0 path1" strings or a file of them."""
    if isinstance(covis_pairs, list):
        pair_list = covis_pairs
    else:
        with open(covis_pairs, "r") as f:
            pair_list = f.read().rstrip("\n").split("\n")
    matches = match_worker(list(range(len(pair_list))), image_lists, pair_list, cfgs, matcher, detector=detector, keep_on_device=True)
    merger = merger if merger is not None else KeypointMerger()
    final_keypoints, final_scores, updated_matches = merger(matches, image_lists, cfgs["matcher"]["pair_name_split"])
    return final_keypoints, final_scores, updated_matches, matches
