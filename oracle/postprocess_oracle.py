"""CPU oracle for the match -> keypoint -> index post-processing.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

Restates, with numpy, what the reference does between the matcher and keypoints.h5 / matches.h5
(src/coarse_match/coarse_match.py:203-237):

  * Match2Kpts (src/coarse_match/utils/merge_kpts.py:19-44): for every image, the (x, y, conf) rows of all pairs it takes
    part in -- columns (0,1,4) when it is the pair's first image, (2,3,4) when it is the second -- concatenated in the
    order of the matches dict;
  * keypoint_worker (src/coarse_match/coarse_match_worker.py:151-173) with agg_groupby_2d (merge_kpts.py:4-17): keys are
    the coordinates truncated to int, np.unique(axis=0) orders them by (x, y), np.bincount sums the confidences in
    float64 in input order; the keypoints are then ranked by summed score, descending, Python's stable sort keeping
    (x, y) order between equal scores; rank = keypoint id;
  * update_matches (coarse_match_worker.py:180-241, merge=False): every match becomes the pair of keypoint ids of its
    two truncated end points -> int array (M, 2);
  * transform_keypoints (coarse_match_worker.py:248-270): keypoints as float32 (n, 2) in id order, scores as float32.

Pinned against the reference functions themselves in tests/test_oracle_vs_reference.py (imported behind stubs, build
container only) and against tests/golden/postprocess_small.npz.
"""
import numpy as np


def split_pair(key, pair_name_split):
    """merge_kpts.py:27-30: split on the configured separator, fall back to '-'."""
    try:
        name0, name1 = key.split(pair_name_split)
    except ValueError:
        name0, name1 = key.split("-")
    return name0, name1


def merge_keypoints(matches, image_lists, pair_name_split=" "):
    """matches: ordered {pair_key: (M,5) float32 [x0,y0,x1,y1,conf]} -> (final_keypoints, final_scores, updated_matches).

    final_keypoints[name]: float32 (n,2) (np.empty((0,2)) float64 for an image without matches, as the reference),
    final_scores[name]: float32 (n,), updated_matches[pair_key]: int64 (M,2).
    """
    names = list(image_lists)
    per_image = {n: [] for n in names}      # name -> list of (pair_key, side)
    for k in matches.keys():
        n0, n1 = split_pair(k, pair_name_split)
        per_image[n0].append((k, 0))
        per_image[n1].append((k, 1))

    final_kpts, final_scores = {}, {}
    ids = {k: np.empty((np.asarray(v).shape[0], 2), dtype=np.int64) for k, v in matches.items()}
    for name in names:
        parts = [np.asarray(matches[k])[:, [2 * s, 2 * s + 1, 4]] for k, s in per_image[name]]
        parts = [p for p in parts]
        if len(parts) == 0 or sum(p.shape[0] for p in parts) == 0:
            final_kpts[name] = np.empty((0, 2))
            final_scores[name] = np.empty((0,), dtype=np.float32)
            continue
        kpts = np.concatenate(parts, 0)
        keys = kpts[:, :2].astype(int)
        uniq, group, _ = np.unique(keys, axis=0, return_inverse=True, return_counts=True)
        group = group.reshape(-1)
        sums = np.bincount(group, weights=kpts[:, -1])               # float64, accumulated in input order
        order = np.argsort(-sums, kind="stable")                     # sorted(..., reverse=True): stable, ties keep (x, y) order
        rank = np.empty_like(order)
        rank[order] = np.arange(order.shape[0])
        final_kpts[name] = uniq[order].astype(np.float32)
        final_scores[name] = sums[order].astype(np.float32)
        pos = 0
        for k, s in per_image[name]:
            m = np.asarray(matches[k]).shape[0]
            ids[k][:, s] = rank[group[pos:pos + m]]
            pos += m
    return final_kpts, final_scores, ids
