"""Seeded synthetic inputs shared by tests, smoke() and bench.py."""
import numpy as np
import torch


def synth_image(h, w, seed, channels=1):
    """Low-pass filtered uniform noise in [0,1]: [1, channels, h, w]."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(1, channels, h + 4, w + 4, generator=g)
    x = torch.nn.functional.avg_pool2d(x, 5, 1)
    return ((x - x.min()) / (x.max() - x.min())).contiguous()


def synth_pair(h, w, seed, shift=(8, 8)):
    im0 = synth_image(h, w, seed)
    im1 = torch.roll(im0, shift, (2, 3)).contiguous()
    return im0, im1


def discriminative_features(L, S, d=256, seed=0, gain=1.1, noise=2.0):
    """feat1 = permuted feat0 + per-row noise, row-normalised to |f| = 16*gain: hundreds..thousands of mutual-NN
    matches with a wide confidence spread, many entries near the threshold (SURVEY.md section 8c recipe)."""
    g = torch.Generator().manual_seed(seed)
    n = max(L, S)
    base = torch.randn(n, d, generator=g)
    perm = torch.randperm(n, generator=g)
    f0 = base[:L].clone()
    f1 = (base[perm] + noise * torch.randn(n, d, generator=g) * torch.rand(n, 1, generator=g))[:S]
    f0 = f0 / f0.norm(dim=1, keepdim=True) * (16 * gain)
    f1 = f1 / f1.norm(dim=1, keepdim=True) * (16 * gain)
    return f0.contiguous(), f1.contiguous()


def loftr_config(thr=0.2, temperature=0.1, fine=False):
    """lower_config(get_cfg_defaults())['loftr'] after coarse_match_worker.py:31-35 (the keys the engine reads)."""
    return {
        "backbone_type": "ResNetFPN", "resolution": (8, 2), "fine_window_size": 5, "fine_concat_coarse_feat": True,
        "resnetfpn": {"initial_dim": 128, "block_dims": [128, 196, 256]},
        "coarse": {"d_model": 256, "d_ffn": 256, "nhead": 8, "layer_names": ["self", "cross"] * 4, "attention": "linear",
                   "temp_bug_fix": False},
        "match_coarse": {"thr": thr, "border_rm": 2, "match_type": "dual_softmax", "dsmax_temperature": temperature,
                         "skh_iters": 3, "skh_init_bin_score": 1.0, "skh_prefilter": False, "train_coarse_percent": 0.2,
                         "train_pad_num_gt_min": 200, "sparse_spvs": True},
        "fine": {"enable": fine, "d_model": 128, "d_ffn": 128, "nhead": 8, "layer_names": ["self", "cross"], "attention": "linear"},
    }


def synth_chunk(M=64, n_img=6, max_views=5, hw=(120, 160), seed=0, scales=None, frozen_frac=0.2):
    """One refinement chunk dict like MatchingMultiviewData.__getitem__ (construct_matching_data.py:317-476) after the
    DataLoader added the batch dim: tracks sorted by #valid query views (descending, :350-352), valid views first,
    -1 image index in the padded slots.  max_views = n_view - 1 (query slots)."""
    g = torch.Generator().manual_seed(seed)
    Nq = max_views
    H, W = hw
    images = [synth_image(H + 8 * (i % 3), W + 8 * ((i + 1) % 2), seed * 100 + i, channels=3) for i in range(n_img)]
    if scales is None:
        scales = torch.stack([torch.tensor([1.0 + 0.25 * (i % 3), 1.0 + 0.5 * (i % 2)]) for i in range(n_img)])[None]
    counts = torch.randint(1, Nq + 1, (M,), generator=g).sort(descending=True)[0]
    counts[0] = Nq
    q_img = torch.randint(0, n_img, (M,), generator=g)
    r_img = torch.full((Nq, M), -1, dtype=torch.long)
    valid = torch.zeros(Nq, M, dtype=torch.bool)
    for t in range(M):
        others = [i for i in torch.randperm(n_img, generator=g).tolist() if i != int(q_img[t])]
        k = min(int(counts[t]), len(others))
        counts[t] = k
        for v in range(k):
            r_img[v, t] = others[v]
            valid[v, t] = True
    order = counts.sort(descending=True, stable=True)[1]
    counts, q_img, r_img, valid = counts[order], q_img[order], r_img[:, order], valid[:, order]

    def pts(img_idx):
        """random points in ORIGINAL-image px, some close to the border so that crops leave the image"""
        idx = img_idx.clamp(min=0)
        hh = torch.tensor([im.shape[2] for im in images], dtype=torch.float32)[idx]
        ww = torch.tensor([im.shape[3] for im in images], dtype=torch.float32)[idx]
        u = torch.rand(*img_idx.shape, 2, generator=g)
        x = (4 + u[..., 0] * (ww - 8)) * scales[0][idx][..., 1]
        y = (4 + u[..., 1] * (hh - 8)) * scales[0][idx][..., 0]
        return torch.stack([x, y], -1)

    movable = torch.rand(M, generator=g) >= frozen_frac
    return {
        "images": images, "scales": scales.float(),
        "query_points": pts(q_img)[None], "reference_points_coarse": pts(r_img)[None],
        "track_valid_mask": valid[None], "query_img_idxs": q_img[None], "reference_img_idxs": r_img[None],
        "query_movable_mask": movable[None],
        "scales_relative": torch.ones(1, Nq + 1, M), "view_point_vector": torch.zeros(1, Nq + 1, M, 3),
    }


def worker_chunks(seeds=(4, 5, 6)):
    """Chunk items as MatchingMultiviewData.__getitem__ yields them (no batch dimension; the DataLoader adds it) for the
    refinement-worker loop tests: colmap ids included, the same reference nodes recur in every chunk."""
    out = []
    for seed in seeds:
        c = synth_chunk(M=12, n_img=4, max_views=3, seed=seed)
        item = {k: (v[0] if torch.is_tensor(v) else v) for k, v in c.items() if k != "images"}
        item["images"] = [im[0] for im in c["images"]]
        item["query_img_ids"] = item["query_img_idxs"].clone() % 2
        item["query_pt2d_idxs"] = torch.arange(12) % 5
        item["reference_img_ids"] = item["reference_img_idxs"].clamp(min=0)
        item["reference_pt2d_idxs"] = (torch.arange(item["reference_img_idxs"].numel()).view_as(item["reference_img_idxs"]) * 7 + seed) % 101
        out.append(item)
    return out


class StandInRefiner:
    """Deterministic stand-in honouring the HP-2 contract (host-loop tests only; no kernels involved)."""

    def cuda(self):
        return self

    def __call__(self, data):
        data["query_points_refined"] = data["query_points"] + 0.25
        data["reference_points_refined"] = [data["reference_points_coarse"] * 0 + 1.0, data["reference_points_coarse"] - 0.5]


def synth_matches(n_images, pairs, m_per_pair, hw=(880, 1200), seed=0, grid=8, dup=0.3):
    """Synthetic matcher output: grid-aligned coordinates (coarse matches sit on the 1/8 grid times a scale), confidences in
    (0.2, 1], a fraction of repeated key points so that the groupby and the ties matter.  -> (matches dict, names)."""
    rng = np.random.default_rng(seed)
    names = [f"/data/scene/img_{i:05d}.jpg" for i in range(n_images)]
    H, W = hw
    scale = np.array([1.25, 1.25], dtype=np.float32)
    matches = {}
    for (i, j) in pairs:
        m = int(m_per_pair if np.isscalar(m_per_pair) else m_per_pair[len(matches) % len(m_per_pair)])
        gx = rng.integers(0, W // grid, size=(m, 2))
        gy = rng.integers(0, H // grid, size=(m, 2))
        if dup > 0 and m > 0:
            # pull a fraction of the points onto a small set of popular cells
            hot = rng.random(m) < dup
            gx[hot] = gx[hot] % 7
            gy[hot] = gy[hot] % 5
        xy0 = np.stack([gx[:, 0] * grid * scale[0], gy[:, 0] * grid * scale[1]], 1).astype(np.float32)
        xy1 = np.stack([gx[:, 1] * grid * scale[0], gy[:, 1] * grid * scale[1]], 1).astype(np.float32)
        conf = (0.2 + 0.8 * rng.random(m)).astype(np.float32)
        if m > 4:
            conf[: m // 4] = np.float32(0.5)   # exact ties in the summed scores
        matches[f"{names[i]} {names[j]}"] = np.concatenate([xy0, xy1, conf[:, None]], 1).astype(np.float32)
    return matches, names


def synth_photo(h, w, seed=0):
    """uint8 test image with smooth structure, edges and noise (full 0..255 range so that the clipping matters)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = 127 + 90 * np.sin(x / 17.0 + seed) * np.cos(y / 23.0) + 60 * ((x // 31 + y // 19) % 2) + rng.normal(0, 25, (h, w))
    img[: h // 8] = 255 * ((x[: h // 8] // 3) % 2)     # hard 0/255 stripes: overshoot of the negative lobes gets clipped
    return np.clip(img, 0, 255).astype(np.uint8)


def synth_scene(n_images, h, w, seed, noise=0.025, max_shift=256, align=8):
    """An overlapping-view scene: ``n_images`` crops [1,1,h,w] of one big low-pass noise image at seeded offsets that are
    multiples of ``align`` px (views related by translation, coarse cells aligned), plus independent per-view Gaussian
    noise -- with the BN-calibrated synthetic weights (tests/weights.py) every pair yields O(10^3) mutual-NN matches with
    confidences spread over (0.2, 1], like a real scene.  Returns (images, offsets[(dy, dx)])."""
    g = torch.Generator().manual_seed(seed)
    base = synth_image(h + max_shift, w + max_shift, seed)
    n_off = max_shift // align + 1
    offs = [(0, 0)] + [(int(torch.randint(0, n_off, (1,), generator=g)) * align, int(torch.randint(0, n_off, (1,), generator=g)) * align)
                       for _ in range(n_images - 1)]
    images = []
    for dy, dx in offs:
        crop = base[:, :, dy:dy + h, dx:dx + w]
        images.append((crop + noise * torch.randn(crop.shape, generator=g)).clamp(0, 1).contiguous())
    return images, offs


def multiview_config(window=15, left_window=7):
    mm = {"enable": True, "type": "s2d", "detector": "OnGrid", "window_size": window, "best_left_strategy": "smallest_mean_std",
          "s2d": {"type": "heatmap", "obtain_offset_method": "argsoftmax"}}
    return {"n_matching_steps": 1, "enable_multiview_scale_align": False,
            "backbone": {"type": "S2DNet", "resolution": [4, 1], "s2dnet": {"window_size": window}, "pretrained": None},
            "multiview_transform": {"sparse": True, "crop_size": 35, "window_size": window, "enable": True, "type": "LoFTR", "d_model": 128,
                                    "nhead": 8, "layer_names": ["self", "cross"], "layer_iter_n": 2, "attention": "linear"},
            "multiview_matching_test": {**mm, "left_point_movement_window_size": left_window}}


def to_cuda(data):
    out = {}
    for k, v in data.items():
        if isinstance(v, list):
            out[k] = [x.cuda() for x in v]
        elif torch.is_tensor(v):
            out[k] = v.cuda()
        else:
            out[k] = v
    return out


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class SynthColmapDataset:
    """The attributes and item protocol of ``CoarseColmapDataset`` (src/dataset/coarse_sfm_refinement_dataset.py) that the chunk
    dataset reads -- colmap_images / colmap_3ds / image_intrin_extrins / keyframe_dict / point_cloud_assigned_imgID_kptID /
    colmapID2frameID_dict and ``ds[frame] -> {'image', 'scale'}`` -- over a seeded synthetic reconstruction: cameras on a ring looking
    at a point cloud, every 3-D point observed by 2..max_obs images (sometimes twice by the same image, as COLMAP tracks do), the
    reference node of a track chosen by the 'middle scale' rule of get_keyframes_by_scale (:236-297)."""

    def __init__(self, n_images=12, n_points=400, max_obs=9, seed=0, hw=(48, 64), dup_frac=0.05, first_image_id=1):
        rng = np.random.default_rng(seed)
        self.img_list = [f"img_{i:04d}.jpg" for i in range(n_images)]
        img_ids = [first_image_id + 3 * i for i in range(n_images)]            # non-contiguous ids, like a filtered COLMAP model
        self.colmapID2frameID_dict = {cid: f for f, cid in enumerate(img_ids)}
        self.image_intrin_extrins = {}
        for k, cid in enumerate(img_ids):
            a = 2 * np.pi * k / n_images
            c = np.array([4 * np.cos(a), 4 * np.sin(a), 0.3 * np.sin(3 * a)])
            z = -c / np.linalg.norm(c)
            x = np.cross([0, 0, 1.0], z); x /= np.linalg.norm(x)
            y = np.cross(z, x)
            R = np.stack([x, y, z])
            f = 500.0 + 20 * k
            self.image_intrin_extrins[cid] = {"intrin": np.array([[f, 0, hw[1] / 2], [0, f, hw[0] / 2], [0, 0, 1.0]]), "extrin": [R, -R @ c]}
        xyz = rng.normal(0, 0.6, (n_points, 3))
        n_kpts = {cid: 0 for cid in img_ids}
        xys = {cid: [] for cid in img_ids}
        self.colmap_3ds, self.point_cloud_assigned_imgID_kptID = {}, {}
        state = {cid: [] for cid in img_ids}
        for p in range(n_points):
            pid = 10 + 7 * p
            k = int(rng.integers(2, max_obs + 1))
            obs = rng.choice(img_ids, size=min(k, n_images), replace=False).tolist()
            if rng.random() < dup_frac:
                obs.append(obs[int(rng.integers(0, len(obs)))])              # the same image twice in one track
            p2d = []
            for cid in obs:
                K, (R, t) = self.image_intrin_extrins[cid]["intrin"], self.image_intrin_extrins[cid]["extrin"]
                pc = K @ (R @ xyz[p] + t)
                xys[cid].append(pc[:2] / pc[2] + rng.normal(0, 0.5, 2))
                state[cid].append(-3)
                p2d.append(n_kpts[cid])
                n_kpts[cid] += 1
            scales = [self.image_intrin_extrins[c]["intrin"][0, 0] / ((self.image_intrin_extrins[c]["intrin"] @ (self.image_intrin_extrins[c]["extrin"][0] @ xyz[p] + self.image_intrin_extrins[c]["extrin"][1]))[2] + 1e-4) for c in obs]
            a_idx = int(np.argsort(np.array(scales))[len(obs) // 2])
            self.colmap_3ds[pid] = _NS(xyz=xyz[p], image_ids=np.array(obs, dtype=np.int64), point2D_idxs=np.array(p2d, dtype=np.int64))
            self.point_cloud_assigned_imgID_kptID[pid] = (np.int64(obs[a_idx]), np.int64(p2d[a_idx]))
            state[obs[a_idx]][p2d[a_idx]] = pid
        self.colmap_images = {cid: _NS(xys=np.array(xys[cid], dtype=np.float64).reshape(-1, 2)) for cid in img_ids}
        self.keyframe_dict = {cid: np.array([s for s in state[cid] if s >= 0], dtype=np.int32) for cid in img_ids}
        self.colmap_cameras = {}
        g = torch.Generator().manual_seed(seed)
        self._items = [{"image": torch.rand(3, hw[0] + 8 * (i % 2), hw[1], generator=g), "scale": torch.tensor([1.0 + 0.25 * (i % 3), 1.5 - 0.25 * (i % 2)]),
                        "f_name": n, "img_name": n, "frameID": i, "img_path": [n]} for i, n in enumerate(self.img_list)]

    def __len__(self):
        return len(self.img_list)

    def __getitem__(self, idx):
        return self._items[idx]
