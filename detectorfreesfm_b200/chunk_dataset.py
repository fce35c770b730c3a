"""SURVEY.md section 8(a) row b2 + 8(f) row 2: the refinement-chunk dataset -- bag assignment, chunking and the per-chunk dict.

``B200MatchingMultiviewData(colmap_image_dataset, config, worker_split_idxs=None)`` is a drop-in for
``MatchingMultiviewData`` (src/post_optimization/data_construct/construct_matching_data.py:165-476): same constructor, ``__len__``
and ``__getitem__`` dict (keys, dtypes, shapes, track order), consumed unchanged by ``matchWorker``
(src/post_optimization/matcher_model/multiview_match_worker.py:111-150) through a DataLoader.

* ``assign_bags`` + ``chunk_bags`` (:202-261, the greedy ``FeatureTrackStatus`` loop with one ``np.argmax`` over all tracks per bag:
  quadratic in Python) run natively in libdfsfm_b200.so (csrc/bag_assign.cpp): same bags in the same order, including the CPython
  set iteration order the reference's result depends on.
* ``__getitem__`` (:317-476, Python loops per track x view with a matrix product per node) is vectorised numpy over flat
  per-model arrays built once; ``scales_relative`` / ``view_point_vector`` (dead inputs of the shipped matcher config)
  are still produced, in float64 like the reference.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def assign_bags(track_ids, ref_img_ids, obs_offset, obs_img_ids, frame_img_ids, frame_offset, frame_track_ids, max_track_length, chunk,
                max_num_img_in_bag=16):
    """Flat-array front end of dfsfm_assign_bags -> dict of CSR arrays over the CHUNKED bags:
    bag_img_off/bag_img (bag image ids), bag_trk_off/trk_id/trk_ref (tracks of a bag, their reference image), trk_q_off/trk_q
    (query image ids of a track inside its bag)."""
    lib = _lib.load_library()
    track_ids, ref_img_ids, obs_offset, obs_img_ids = _i64(track_ids), _i64(ref_img_ids), _i64(obs_offset), _i64(obs_img_ids)
    frame_img_ids, frame_offset, frame_track_ids = _i64(frame_img_ids), _i64(frame_offset), _i64(frame_track_ids)
    h = ctypes.c_void_p()
    _lib.check(lib.dfsfm_assign_bags(ctypes.byref(h), len(track_ids), _p(track_ids), _p(ref_img_ids), _p(obs_offset), _p(obs_img_ids),
                                     len(frame_img_ids), _p(frame_img_ids), _p(frame_offset), _p(frame_track_ids), int(max_track_length),
                                     int(max_num_img_in_bag), int(chunk)))
    try:
        n = [ctypes.c_int64(0) for _ in range(4)]
        lib.dfsfm_bags_sizes(h, *[ctypes.byref(x) for x in n])
        n_bags, n_img, n_trk, n_q = [x.value for x in n]
        out = {"bag_img_off": np.zeros(n_bags + 1, np.int64), "bag_img": np.zeros(n_img, np.int64), "bag_trk_off": np.zeros(n_bags + 1, np.int64),
               "trk_id": np.zeros(n_trk, np.int64), "trk_ref": np.zeros(n_trk, np.int64), "trk_q_off": np.zeros(n_trk + 1, np.int64),
               "trk_q": np.zeros(n_q, np.int64)}
        lib.dfsfm_bags_export(h, _p(out["bag_img_off"]), _p(out["bag_img"]), _p(out["bag_trk_off"]), _p(out["trk_id"]), _p(out["trk_ref"]),
                              _p(out["trk_q_off"]), _p(out["trk_q"]))
    finally:
        lib.dfsfm_bags_destroy(h)
    return out


class B200MatchingMultiviewData(torch.utils.data.Dataset):
    """Construct image bags for MultiviewMatcher (same interface as MatchingMultiviewData)."""

    def __init__(self, colmap_image_dataset, config, worker_split_idxs=None):
        super().__init__()
        self.max_track_length = config["max_track_length"]
        self.chunk = config["chunk"]
        ds = self.colmap_image_dataset = colmap_image_dataset
        self.colmap_3ds = ds.colmap_3ds
        self.colmap_images = ds.colmap_images
        self.colmap_intrin_extrin = ds.image_intrin_extrins
        assignment_all = ds.point_cloud_assigned_imgID_kptID
        if worker_split_idxs is None:
            self.point3d_assignment = assignment_all
        else:                                       # :180-188: the worker's share, in the order of its index list
            items = list(assignment_all.items())
            self.point3d_assignment = {items[i][0]: items[i][1] for i in worker_split_idxs}

        # ---- flat model arrays (built once)
        tids = _i64(list(self.point3d_assignment.keys()))
        self._track_index = {int(t): i for i, t in enumerate(tids.tolist())}
        ref_img = _i64([self.point3d_assignment[int(t)][0] for t in tids.tolist()])
        self._ref_p2d = _i64([self.point3d_assignment[int(t)][1] for t in tids.tolist()])
        obs_img = [np.asarray(self.colmap_3ds[int(t)].image_ids, dtype=np.int64) for t in tids.tolist()]
        obs_p2d = [np.asarray(self.colmap_3ds[int(t)].point2D_idxs, dtype=np.int64) for t in tids.tolist()]
        self._obs_off = np.concatenate([[0], np.cumsum([len(a) for a in obs_img])]).astype(np.int64)
        self._obs_img = np.concatenate(obs_img) if obs_img else np.zeros(0, np.int64)
        self._obs_p2d = np.concatenate(obs_p2d) if obs_p2d else np.zeros(0, np.int64)
        self._xyz = np.stack([np.asarray(self.colmap_3ds[int(t)].xyz, dtype=np.float64) for t in tids.tolist()]) if len(tids) else np.zeros((0, 3))
        frame_ids = list(ds.keyframe_dict.keys())
        frame_tracks = [np.asarray(ds.keyframe_dict[f], dtype=np.int64) for f in frame_ids]
        frame_off = np.concatenate([[0], np.cumsum([len(a) for a in frame_tracks])]).astype(np.int64)
        frame_cat = np.concatenate(frame_tracks) if frame_tracks else np.zeros(0, np.int64)
        # cameras: image id -> row of stacked intrinsics / extrinsics
        img_ids = list(self.colmap_intrin_extrin.keys())
        self._cam_row = {int(i): r for r, i in enumerate(img_ids)}
        self._K = np.stack([np.asarray(self.colmap_intrin_extrin[i]["intrin"], dtype=np.float64) for i in img_ids])
        self._R = np.stack([np.asarray(self.colmap_intrin_extrin[i]["extrin"][0], dtype=np.float64) for i in img_ids])
        self._t = np.stack([np.asarray(self.colmap_intrin_extrin[i]["extrin"][1], dtype=np.float64) for i in img_ids])
        self._cam_lut = np.full(max(img_ids) + 1 if img_ids else 1, -1, dtype=np.int64)
        for i, r in self._cam_row.items():
            self._cam_lut[i] = r

        # (track, image) -> first observation / number of observations of that image in the track (np.where(image_ids == id), :397-400)
        self._img_span = int(self._obs_img.max()) + 1 if len(self._obs_img) else 1
        okey = np.repeat(np.arange(len(tids), dtype=np.int64), np.diff(self._obs_off)) * self._img_span + self._obs_img
        self._ukey, first, self._ucount = np.unique(okey, return_index=True, return_counts=True)
        self._ufirst_p2d = self._obs_p2d[first]

        # ---- assign_bags + chunk_bags, natively
        self._bags = assign_bags(tids, ref_img, self._obs_off, self._obs_img, frame_ids, frame_off, frame_cat, self.max_track_length, self.chunk)
        self._tids = tids

    # the reference's list-of-dicts view (same content, same order), for inspection and for the parity tests
    @property
    def image_bags(self):
        b = self._bags
        out = []
        for i in range(len(self)):
            t0, t1 = b["bag_trk_off"][i], b["bag_trk_off"][i + 1]
            out.append({
                "bag_image_ids": b["bag_img"][b["bag_img_off"][i]:b["bag_img_off"][i + 1]].tolist(),
                "track_ids": b["trk_id"][t0:t1].tolist(),
                "track_corresponding_imgs": [[int(b["trk_ref"][t]), b["trk_q"][b["trk_q_off"][t]:b["trk_q_off"][t + 1]].tolist()] for t in range(t0, t1)],
            })
        return out

    def __len__(self):
        return len(self._bags["bag_img_off"]) - 1

    # ------------------------------------------------------------------------------------------------ geometry (float64)
    def _point_scale(self, img_ids, xyz):
        """get_point_scale (:283-291): f / (depth + 1e-4), depth = (K (R p + t))_z."""
        r = self._cam_lut[img_ids]
        pc = np.einsum("nij,nj->ni", self._R[r], xyz) + self._t[r]
        depth = np.einsum("nj,nj->n", self._K[r][:, 2, :], pc)
        return self._K[r][:, 0, 0] / (depth + 1e-4)

    def _view_point(self, src_ids, dst_ids, xyz):
        """get_relative_view_point (:293-311) for arrays of (reference image, query image, 3D point)."""
        rs, rd = self._cam_lut[src_ids], self._cam_lut[dst_ids]
        n = len(rs)
        Ts = np.tile(np.eye(4), (n, 1, 1))
        Td = np.tile(np.eye(4), (n, 1, 1))
        Ts[:, :3, :3], Ts[:, :3, 3] = self._R[rs], self._t[rs]
        Td[:, :3, :3], Td[:, :3, 3] = self._R[rd], self._t[rd]
        f = np.einsum("nij,nj->ni", self._R[rs], xyz) + self._t[rs]
        rel = Ts @ np.linalg.inv(Td)
        t = rel[:, :3, 3]
        a = f - t
        fn, tn, an = (np.linalg.norm(x, axis=-1, keepdims=True) for x in (f, t, a))
        alpha = np.arccos(np.sum(f * t, -1, keepdims=True) / (fn * tn + 1e-6))
        beta = np.arccos(np.sum(a * (-t), -1, keepdims=True) / (an * tn + 1e-6))
        gamma = np.pi - alpha - beta
        return (t / (tn + 1e-6)) * gamma

    def _gather_xy(self, img_ids, p2d):
        """colmap_images[img].xys[p2d] for arrays of (image id, key-point index): one vectorised gather per distinct image (live arrays:
        refined key points written back between iterations are seen)."""
        out = np.zeros((len(img_ids), 2), dtype=np.float64)
        for i in np.unique(img_ids).tolist():
            sel = img_ids == i
            out[sel] = self.colmap_images[int(i)].xys[p2d[sel]]
        return out

    # ------------------------------------------------------------------------------------------------ one chunk
    def __getitem__(self, index):
        b = self._bags
        bag_image_ids = b["bag_img"][b["bag_img_off"][index]:b["bag_img_off"][index + 1]]
        ds = self.colmap_image_dataset
        singles = [ds[ds.colmapID2frameID_dict[int(i)]] for i in bag_image_ids.tolist()]
        data = {"images": [s["image"] for s in singles]}                       # buildDataBag (:263-281)
        if "scale" in singles[0]:
            data["scales"] = torch.stack([s["scale"] for s in singles], dim=0)
        bag_idx = {int(i): k for k, i in enumerate(bag_image_ids.tolist())}
        t0, t1 = int(b["bag_trk_off"][index]), int(b["bag_trk_off"][index + 1])
        M, Nq = t1 - t0, len(bag_image_ids) - 1
        nq = (b["trk_q_off"][t0 + 1:t1 + 1] - b["trk_q_off"][t0:t1]).astype(np.int64)
        order = np.argsort(-nq, kind="stable")                                  # sorted(..., key=len(query), reverse=True) is stable (:350-352)
        trk = np.arange(t0, t1)[order]
        nq = nq[order]
        tid = b["trk_id"][trk]
        tix = np.asarray([self._track_index[int(t)] for t in tid.tolist()], dtype=np.int64)
        ref_img = b["trk_ref"][trk]
        ref_p2d = self._ref_p2d[tix]
        xyz = self._xyz[tix]
        reference_node = self._gather_xy(ref_img, ref_p2d)
        ref_scale = self._point_scale(ref_img, xyz) if M else np.zeros(0)

        query_nodes = np.ones((M, Nq, 2), dtype=np.float64)                     # padding values of :386-393
        query_mask = np.zeros((M, Nq), dtype=bool)
        query_img_idx = np.full((M, Nq), -1, dtype=np.int64)
        query_img_ids = np.full((M, Nq), -1, dtype=np.int64)
        query_p2d = np.full((M, Nq), -1, dtype=np.int64)
        query_scale = np.repeat(ref_scale[:, None], Nq, axis=1) if Nq else np.zeros((M, 0))
        view_pts = np.zeros((M, Nq, 3), dtype=np.float64)
        if M and Nq:
            rows = np.repeat(np.arange(M), nq)
            cols = np.concatenate([np.arange(k) for k in nq.tolist()]) if len(rows) else np.zeros(0, np.int64)
            qimg = np.concatenate([b["trk_q"][b["trk_q_off"][t]:b["trk_q_off"][t + 1]] for t in trk.tolist()]) if len(rows) else np.zeros(0, np.int64)
            u = np.searchsorted(self._ukey, tix[rows] * self._img_span + qimg)
            first_p2d = self._ufirst_p2d[u]
            node = self._gather_xy(qimg, first_p2d)
            for k in np.nonzero(self._ucount[u] > 1)[0].tolist():              # an image observed twice by one track: mean of its key points (:401-404)
                r, qi = int(rows[k]), int(qimg[k])
                o0, o1 = self._obs_off[tix[r]], self._obs_off[tix[r] + 1]
                hit = self._obs_p2d[o0:o1][self._obs_img[o0:o1] == qi]
                node[k] = np.mean(self.colmap_images[qi].xys[hit], axis=0)
            query_nodes[rows, cols] = node
            query_mask[rows, cols] = True
            query_img_idx[rows, cols] = [bag_idx[int(q)] for q in qimg.tolist()]
            query_img_ids[rows, cols] = qimg
            query_p2d[rows, cols] = first_p2d
            if len(rows):
                query_scale[rows, cols] = self._point_scale(qimg, xyz[rows])
                view_pts[rows, cols] = self._view_point(ref_img[rows], qimg, xyz[rows])

        ref_img_idx = np.asarray([bag_idx[int(i)] for i in ref_img.tolist()], dtype=np.int64)
        scales_absolute = torch.cat([torch.from_numpy(ref_scale)[..., None], torch.from_numpy(query_scale)], dim=-1)       # M x N
        if "scales" in data:
            scales_absolute /= data["scales"][..., 0][None]
        scales_relative = scales_absolute / (scales_absolute[..., [0]])
        relative_view_points = torch.cat([torch.zeros((M, 1, 3)), torch.from_numpy(view_pts)], dim=-2)
        tr = lambda a: torch.from_numpy(a).transpose(0, 1)  # noqa: E731
        data.update({
            "query_points": torch.from_numpy(reference_node).to(torch.float32) - 0.5,            # [M, 2]
            "reference_points_coarse": tr(query_nodes).to(torch.float32) - 0.5,                  # [N-1, M, 2]
            "track_valid_mask": tr(query_mask),                                                  # [N-1, M]
            "query_img_idxs": torch.from_numpy(ref_img_idx),                                     # [M]
            "reference_img_idxs": tr(query_img_idx),                                             # [N-1, M]
            "scales_relative": scales_relative.transpose(0, 1),                                  # [N, M]
            "view_point_vector": relative_view_points.transpose(0, 1),                           # [N, M, 3]
            "query_img_ids": torch.from_numpy(np.asarray(ref_img)),                              # [M]
            "query_pt2d_idxs": torch.from_numpy(np.asarray(ref_p2d)),                            # [M]
            "reference_img_ids": tr(query_img_ids),                                              # [N-1, M]
            "reference_pt2d_idxs": tr(query_p2d),                                                # [N-1, M]
        })
        return data
