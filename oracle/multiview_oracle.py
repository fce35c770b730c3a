"""Functional CPU restatement of the multi-view refinement matcher (HP-2).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference: src/MultiviewMatcher/ (paths below relative to
/root/reference/src/MultiviewMatcher unless stated), inference path ``forward(data, chunk_track=1000,
chunk_backbone_img=True)`` with the shipped config (hydra_training_configs/experiment/
multiview_refinement_matching.yaml:22-90): S2DNet backbone, sparse 35x35 RoIAlign crops, 4-layer d=128
multiview linear-attention transformer, s2d heatmap matching with a 7x7 movable reference point.

Weights: plain dict with the reference's state_dict keys (oracle/weights.py::multiview_state_dict).
``q``: optional operand quantiser for the precision study (see loftr_oracle.py).
"""
import torch
import torch.nn.functional as F

from . import build_native
from .loftr_oracle import _id, create_meshgrid, encoder_layer, spatial_expectation2d

MEAN = [0.485, 0.456, 0.406]
STD = [0.229, 0.224, 0.225]
VGG_CONV_IDX = [0, 2, 5, 7, 10, 12, 14]
CROP = 35


# ------------------------------------------------------------------------------------------ patches
def extract_patches(image, points_xy, crop=CROP):
    """matcher_module/fine_preprocess.py:92-106 _extract_local_patches + third_party/RoIAlign.pytorch/roi_align/
    roi_align.py:17-48 (transform_fpcoor=False): boxes kp +- crop//2, normalised by (W-1),(H-1), reordered to
    (y1,x1,y2,x2), then the native crop_and_resize.  image [1,3,H,W]; points [P,2] (x,y) in resized-image px."""
    H, W = image.shape[2:]
    r = crop // 2
    boxes = torch.cat([points_xy - r, points_xy + r], dim=-1).to(torch.float32)
    x1, y1, x2, y2 = torch.split(boxes, 1, dim=1)
    x1 = x1 / float(W - 1)
    x2 = x2 / float(W - 1)
    y1 = y1 / float(H - 1)
    y2 = y2 / float(H - 1)
    nb = torch.cat((y1, x1, y2, x2), 1)
    return build_native.roialign_forward(image, nb, torch.zeros(nb.shape[0], dtype=torch.int32), crop, crop, 0.0)


# ------------------------------------------------------------------------------------------ S2DNet
def s2dnet_forward(patches, sd, window, q=_id, p="backbone"):
    """backbone/S2DNet/s2dnet.py:127-193 with num_layers=2, substitute_pooling_layers (MaxPool2d(3,2,1), :89-92),
    combine=True (bicubic align_corners upsample + add, :164-171), zoomin_strategy='post' with scales=None (centre
    crop window x window, :193).  patches [P,3,35,35] -> [P, window*window, 128]."""
    mean = patches.new_tensor(MEAN)[:, None, None]
    std = patches.new_tensor(STD)[:, None, None]
    x = (patches - mean) / std

    def conv(x, i):
        return F.conv2d(q(x), q(sd[f"{p}.encoder.{i}.weight"]), sd[f"{p}.encoder.{i}.bias"], 1, 1)

    x = F.relu(conv(x, 0))
    f0 = F.relu(conv(x, 2))                       # relu1_2 (inplace in the reference: the tap sees the relu)
    x = F.max_pool2d(f0, 3, 2, 1)
    x = F.relu(conv(x, 5))
    x = F.relu(conv(x, 7))
    x = F.max_pool2d(x, 3, 2, 1)
    x = F.relu(conv(x, 10))
    x = F.relu(conv(x, 12))
    f1 = F.relu(conv(x, 14))                      # relu3_3

    def adap(f, i):
        a = f"{p}.adaptation_layers.adap_layer_{i}"
        y = F.relu(F.conv2d(q(f), q(sd[a + ".0.weight"]), sd[a + ".0.bias"]))
        y = F.conv2d(q(y), q(sd[a + ".2.weight"]), sd[a + ".2.bias"], 1, 2)
        return F.batch_norm(y, sd[a + ".3.running_mean"], sd[a + ".3.running_var"], sd[a + ".3.weight"], sd[a + ".3.bias"],
                            False, 0.0, 1e-5)

    fmap = adap(f0, 0)
    fmap = fmap + F.interpolate(adap(f1, 1), size=fmap.shape[2:], mode="bicubic", align_corners=True)
    c = fmap.shape[-1] // 2
    r = window // 2
    fmap = fmap[..., c - r:c + r + 1, c - r:c + r + 1]
    return fmap.flatten(2).transpose(1, 2).contiguous()   # 'm c h w -> m (h w) c'


# --------------------------------------------------------------------------------------- transformer
def multiview_transformer(ref, query, query_mask, sd, p="fine_transformer", nhead=8, layer_names=("self", "cross") * 2, q=_id):
    """matcher_module/transformer.py:132-177 (attention_type 'multiview').  ref [m,WW,C], query [m,n,WW,C],
    query_mask [m,n] bool.  NOTE the cross layer feeds the PRE-update tensors to both directions (:162-167)."""
    m, n, WW, C = query.shape
    qf = query.reshape(m, n * WW, C)
    qm = query_mask[:, :, None].expand(m, n, WW).reshape(m, n * WW) if query_mask is not None else None
    for i, name in enumerate(layer_names):
        lp = f"{p}.layers.{i}"
        if name == "self":
            ref, qf = (encoder_layer(ref, ref, sd, lp, nhead, None, None, q),
                       encoder_layer(qf, qf, sd, lp, nhead, qm, qm, q))
        else:
            src0, src1 = ref, qf
            qf, ref = (encoder_layer(qf, src0, sd, lp, nhead, qm, None, q),
                       encoder_layer(ref, src1, sd, lp, nhead, None, qm, q))
    return ref, qf.reshape(m, n, WW, C)


# ------------------------------------------------------------------------------------------ matching
def fine_matching(ref, query, query_pts, ref_pts, scale_q, scale_r, track_mask, movable, W, left_W, q=_id):
    """utils/fine_matching.py:36-98 with left_point_movement (test config).  ref [m,WW,C], query [m,n,WW,C],
    query_pts [m,2], ref_pts [m,n,2] (orig px), scale_q [m,2], scale_r [m,n,2], track_mask [m,n], movable [m].
    -> query_refined [m,2], ref_refined [m,n,2], std [m,n]."""
    m, n, WW, C = query.shape
    r = left_W // 2
    cgrid = ref.view(m, W, W, C)
    picked = cgrid[:, W // 2 - r:W // 2 + r + 1, W // 2 - r:W // 2 + r + 1, :].flatten(1, 2)      # [m, L, C] (:100-119)
    L = picked.shape[1]
    sim = torch.einsum("mlc,mnrc->mlnr", q(picked), q(query))                                       # (:205)
    heat = torch.softmax(sim / C ** .5, dim=-1).view(m, L * n, W, W)
    coords = spatial_expectation2d(heat)                                                            # [m, L*n, 2] (:258-284)
    grid = create_meshgrid(W, W).reshape(1, 1, -1, 2)
    var = torch.sum(grid ** 2 * heat.view(m, L * n, WW, 1), dim=-2) - coords ** 2
    std = torch.sum(torch.sqrt(torch.clamp(var, min=1e-10)), -1)
    coords = coords.view(m, L, n, 2)
    std = std.view(m, L, n)
    tm = track_mask[:, None, :].expand(m, L, n).float()
    score = (tm * std).sum(-1) / tm.sum(-1).clamp(min=1)                                            # masked_mean (:254-256)
    best = torch.min(score, dim=-1)[1]
    best = torch.where(movable, best, torch.full_like(best, L // 2))                                # (:167)
    left = torch.stack([best % left_W, best // left_W], dim=-1)
    left_norm = (left / (left_W - 1)) * 2 - 1
    ids = torch.arange(m)
    coords_sel, std_sel = coords[ids, best], std[ids, best]
    query_refined = query_pts + left_norm * (left_W // 2) * scale_q                                 # build_moved_query (:221-232)
    ref_refined = ref_pts + coords_sel * (W // 2) * scale_r                                         # build_mkpts (:234-252)
    return query_refined, ref_refined, std_sel


# -------------------------------------------------------------------------------------------- forward
@torch.no_grad()
def multiview_forward(data, sd, window=15, left_window=7, chunk_track=1000, q=None, keep=False):
    """MultiviewMatcher.forward (MultiviewMatcher.py:59-405), n_steps=1, images as a list (no padding).

    data: images list of [1,3,H,W]; scales [1,N_img,2] (h,w); query_points [1,M,2]; reference_points_coarse
    [1,N-1,M,2]; track_valid_mask [1,N-1,M]; query_img_idxs [1,M]; reference_img_idxs [1,N-1,M];
    query_movable_mask [1,M].  Returns dict with query_points_refined [1,M,2], reference_points_refined
    [1,N-1,M,2], std [1,N-1,M] (zeros in the padded slots, like the F.pad at :376-377)."""
    q = q or _id
    images = data["images"]
    N_img = len(images)
    fine_scale = torch.full((1, N_img, 2), 1.0)               # backbone.resolution[-1] == 1 (:70-74)
    scales = fine_scale * data["scales"][:, :, [1, 0]] if "scales" in data else fine_scale
    ref_loc = data["reference_points_coarse"]
    pts = torch.cat([data["query_points"][:, None], ref_loc], dim=1)                  # [1, n_view, M, 2]
    img_idxs = torch.cat([data["query_img_idxs"][:, None], data["reference_img_idxs"]], dim=1)
    _, n_view, M = img_idxs.shape
    pt_scales = scales.view(-1, 2)[img_idxs.view(-1)].view(1, n_view, M, 2)           # idx -1 -> last image (:99-101)
    pts = pts / pt_scales
    flat_idx = img_idxs.reshape(-1)
    flat_pts = pts.reshape(-1, 2)
    WW = window * window
    feats = torch.zeros(n_view * M, WW, 128)
    patches_keep = {}
    for i in range(N_img):                                                             # (:188-261)
        sel = flat_idx == i
        if sel.sum() == 0:
            continue
        patches = extract_patches(images[i], flat_pts[sel])
        if keep:
            patches_keep[i] = patches
        feats[sel] = s2dnet_forward(patches, sd, window, q)
    # slots with img idx -1 keep zeros here (the reference indexes row -1 of the concatenation: masked garbage)
    feats = feats.view(n_view, M, WW, 128).permute(1, 0, 2, 3)                         # [M, n_view, WW, C]
    valid = data["track_valid_mask"][0].transpose(0, 1)                                # [M, n_view-1]
    counts = valid.sum(1)
    movable = data["query_movable_mask"][0] if "query_movable_mask" in data else torch.ones(M, dtype=torch.bool)
    out_q = torch.zeros(M, 2)
    out_r = torch.zeros(M, n_view - 1, 2)
    out_s = torch.zeros(M, n_view - 1)
    max_view_tracks = 16 * chunk_track
    i = 0
    while i < M:                                                                       # grouping (:118-133)
        k = int(counts[i])
        assert k >= 1, "a track needs at least one valid query view"
        j = i
        while j < M and int(counts[j]) == k:
            j += 1
        step = max_view_tracks // (k + 1)
        for a in range(i, j, step):
            b = min(a + step, j)
            sl = slice(a, b)
            assert bool(valid[sl, :k].all()) and not bool(valid[sl, k:].any()), "valid views must be a prefix, tracks sorted desc"
            ref_f, qry_f = multiview_transformer(feats[sl, 0], feats[sl, 1:k + 1], valid[sl, :k], sd, q=q)
            qr, rr, ss = fine_matching(ref_f, qry_f, data["query_points"][0, sl], ref_loc[0].transpose(0, 1)[sl, :k],
                                       pt_scales[0, 0, sl], pt_scales[0, 1:k + 1, sl].transpose(0, 1), valid[sl, :k],
                                       movable[sl], window, left_window, q)
            out_q[sl] = qr
            out_r[sl, :k] = rr
            out_s[sl, :k] = ss
        i = j
    out = {"query_points_refined": out_q[None], "reference_points_refined": out_r.transpose(0, 1)[None].contiguous(),
           "std": out_s.transpose(0, 1)[None].contiguous()}
    if keep:
        out["features"] = feats
        out["patches"] = patches_keep
    return out
