"""The refinement worker loop around HP-2: ``matchWorker`` / ``extract_results`` /
``UpdatedQueryPts`` of src/post_optimization/matcher_model/multiview_match_worker.py:59-150, for any matcher honouring the HP-2
contract (``matcher(data)`` adds ``query_points_refined`` [1,M,2] and ``reference_points_refined`` (list, last [1,N-1,M,2])).

The chunk dataset (row b2, ``MatchingMultiviewData``) stays the reference's: ``match_worker`` takes any iterable of its chunk
dicts (what its DataLoader yields, leading batch dimension 1).

A note on ``UpdatedQueryPts`` (:85-109).  Its purpose is to freeze a reference key point once a chunk has moved it, so that
later chunks sharing the node see the moved position and ``query_movable_mask=False``.  As written, the membership test
``query_pt2d_idx in self.updated_dict[query_img_id]`` (:95) receives a 0-dim torch tensor (iterating ``data['query_pt2d_idxs'][0]``)
whose hash is its identity, so it never finds the numpy-int keys stored by ``update_query_pts`` (:107-109): every point is always
movable and ``query_points`` pass through unchanged (checked against the reference class in tests/test_oracle_vs_reference.py).
``freeze=False`` (default) reproduces that observable behaviour; ``freeze=True`` implements the evident intent with integer keys
-- chunks that share a reference node then become order dependent, exactly the cross-chunk state SURVEY.md section 7 describes.
"""
import numpy as np
import torch


def dict_to_device(data, device):
    """src/utils/data_io.py:8-33 (dict_to_cuda / list_to_cuda) for an explicit device."""
    if isinstance(data, torch.Tensor):
        return data.to(device)
    if isinstance(data, dict):
        return {k: dict_to_device(v, device) for k, v in data.items()}
    if isinstance(data, list):
        return [dict_to_device(v, device) for v in data]
    return data


class UpdatedQueryPts:
    def __init__(self, colmap_image_ids, freeze=False):
        self.updated_dict = {int(i): {} for i in colmap_image_ids}
        self.freeze = freeze

    def find_movable_and_update(self, data):
        img_ids = data["query_img_ids"][0].cpu().numpy()
        pt_idxs = data["query_pt2d_idxs"][0].cpu().numpy()
        pts = data["query_points"][0].to(torch.float32).clone()
        movable = np.ones((len(img_ids),), dtype=bool)
        if self.freeze:
            for idx, (img_id, pt_idx) in enumerate(zip(img_ids.tolist(), pt_idxs.tolist())):
                hit = self.updated_dict[int(img_id)].get(int(pt_idx))
                if hit is not None:                       # already moved by an earlier chunk
                    movable[idx] = False
                    pts[idx] = torch.as_tensor(hit, dtype=torch.float32)
        data.update({"query_points": pts[None], "query_movable_mask": torch.from_numpy(movable)[None]})

    def update_query_pts(self, kpts_refined, image_ids, pt2D_idxs):
        for kpt2D, image_id, pt2D_idx in zip(kpts_refined, image_ids, pt2D_idxs):
            self.updated_dict[int(image_id)][int(pt2D_idx)] = kpt2D


@torch.no_grad()
def extract_results(data, matcher):
    """multiview_match_worker.py:59-82."""
    matcher(data)
    reference_points_refined = data["query_points_refined"].cpu().numpy()      # 1 * n_track * 2
    reference_img_ids = data["query_img_ids"].cpu().numpy()
    reference_pt2D_idxs = data["query_pt2d_idxs"].cpu().numpy()
    ref_movable_mask = data["query_movable_mask"].cpu().numpy()
    query_points_refined = data["reference_points_refined"][-1].cpu().numpy()  # 1 * n_view-1 * n_track * 2
    query_img_ids = data["reference_img_ids"].cpu().numpy()
    query_pt2D_idxs = data["reference_pt2d_idxs"].cpu().numpy()
    mask = data["track_valid_mask"].cpu().numpy()
    assert query_points_refined.shape[0] == 1
    return ([query_points_refined[mask], query_img_ids[mask], query_pt2D_idxs[mask]],
            [reference_points_refined[ref_movable_mask], reference_img_ids[ref_movable_mask], reference_pt2D_idxs[ref_movable_mask]],
            data.get("time"))


@torch.no_grad()
def match_worker(chunks, matcher, colmap_image_ids, device=None, freeze=False):
    """multiview_match_worker.py:111-150: -> list of [K,4] arrays (x, y, image id, point2D index), one per chunk."""
    if device is None and torch.cuda.is_available():
        device = torch.device("cuda", torch.cuda.current_device())
    buf = UpdatedQueryPts(colmap_image_ids, freeze=freeze)
    results = []
    for data in chunks:
        buf.find_movable_and_update(data)
        data_c = dict_to_device(data, device) if device is not None else data
        (q_pts, q_img, q_idx), (r_pts, r_img, r_idx), _ = extract_results(data_c, matcher)
        buf.update_query_pts(r_pts, r_img, r_idx)
        pts = np.concatenate([q_pts, r_pts], axis=0)
        img_ids = np.concatenate([q_img, r_img], axis=0)
        pt_idxs = np.concatenate([q_idx, r_idx], axis=0)
        results.append(np.concatenate([pts, img_ids[:, None], pt_idxs[:, None]], axis=1))   # M * 4
    return results


def update_refined_kpts_to_colmap_multiview(colmap_images, fine_match_results):
    """``CoarseColmapDataset.update_refined_kpts_to_colmap_multiview`` (src/dataset/coarse_sfm_refinement_dataset.py:333-341): write the
    refined key points of ``match_worker``'s ``[K,4]`` arrays (x, y, image id, point2D index) back into ``colmap_images[id].xys`` (+0.5 px),
    to every key point of the image that observes the same 3-D point (``point3D_ids`` duplicates).  The reference walks the results one by
    one with an ``np.where`` over the image's key points each (quadratic); here the rows are grouped by image and applied with one
    vectorised assignment per image, keeping the reference's order semantics: a later row for the same 3-D point overwrites an earlier one."""
    rows = [np.asarray(r, dtype=np.float64).reshape(-1, 4) for r in fine_match_results]
    if not rows:
        return
    rows = np.concatenate(rows, 0)
    if rows.shape[0] == 0:
        return
    img = rows[:, 2].astype(np.int64)
    order = np.argsort(img, kind="stable")                       # per image, rows stay in result order
    img_sorted = img[order]
    starts = np.flatnonzero(np.r_[True, img_sorted[1:] != img_sorted[:-1]])
    ends = np.r_[starts[1:], len(order)]
    for a, b in zip(starts.tolist(), ends.tolist()):
        image = colmap_images[int(img_sorted[a])]
        sel = order[a:b]
        p3d = image.point3D_ids[rows[sel, 3].astype(np.int64)]
        uniq, first_rev = np.unique(p3d[::-1], return_index=True)      # last occurrence of every 3-D point id wins
        last = len(p3d) - 1 - first_rev
        k = np.searchsorted(uniq, image.point3D_ids)
        k = np.minimum(k, len(uniq) - 1)
        hit = uniq[k] == image.point3D_ids
        image.xys[hit] = rows[sel[last[k[hit]]], :2] + 0.5
