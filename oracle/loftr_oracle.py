"""Functional CPU restatement of the LoFTR coarse(-to-fine) matcher (HP-1).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Weights are a plain ``dict`` with the
reference's ``state_dict`` key names, so the very same dict
can be loaded into the reference ``LoFTR`` module when /root/reference is present.

Reference: third_party/LoFTR/src/loftr/ (paths below are relative to that dir).

``q`` is an optional operand quantiser applied to both operands of every dense
contraction (conv / linear / einsum).  It exists only so that the precision study in
DESIGN.md (fp16-operand / fp32-accumulate tensor-core arithmetic) can be emulated on
the CPU; the oracle proper uses ``q=None`` (pure fp32).
"""
import math

import torch
import torch.nn.functional as F

DEFAULT_CFG = {
    # src/config/default.py:5-45 as overwritten by coarse_match_worker.py:31-35
    "d_model": 256, "nhead": 8, "layer_names": ["self", "cross"] * 4,
    "thr": 0.2, "border_rm": 2, "temperature": 0.1,
    "fine_enable": False, "fine_window": 5, "fine_d_model": 128, "fine_nhead": 8,
    "fine_layer_names": ["self", "cross"],
    # The reference always evaluates the 1/2-resolution FPN branch, even when fine.enable=False and its output is unused
    # (resnet_fpn.py:110-118).  The oracle skips it by default (same results); the CPU-baseline legs of bench.py switch
    # this on so that the timed work is exactly the reference's.
    "compute_unused_fine_branch": False,
}


def _id(x):
    return x


# ----------------------------------------------------------------------------- backbone
def _conv(x, w, stride, pad, q):
    return F.conv2d(q(x), q(w), None, stride, pad)


def _bn(x, sd, p):
    # nn.BatchNorm2d in eval mode, eps=1e-5 (backbone/resnet_fpn.py:21-22)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)


def basic_block(x, sd, p, stride, q=_id):
    """backbone/resnet_fpn.py:32-40 BasicBlock.forward."""
    y = F.relu(_bn(_conv(x, sd[p + ".conv1.weight"], stride, 1, q), sd, p + ".bn1"))
    y = _bn(_conv(y, sd[p + ".conv2.weight"], 1, 1, q), sd, p + ".bn2")
    if stride != 1:
        x = _bn(_conv(x, sd[p + ".downsample.0.weight"], stride, 0, q), sd, p + ".downsample.1")
    return F.relu(x + y)


def resnet_fpn_8_2(x, sd, fine=True, q=_id, p="backbone"):
    """backbone/resnet_fpn.py:100-118 ResNetFPN_8_2.forward -> (x3_out, x1_out|None)."""
    x0 = F.relu(_bn(_conv(x, sd[p + ".conv1.weight"], 2, 3, q), sd, p + ".bn1"))
    x1 = basic_block(basic_block(x0, sd, p + ".layer1.0", 1, q), sd, p + ".layer1.1", 1, q)
    x2 = basic_block(basic_block(x1, sd, p + ".layer2.0", 2, q), sd, p + ".layer2.1", 1, q)
    x3 = basic_block(basic_block(x2, sd, p + ".layer3.0", 2, q), sd, p + ".layer3.1", 1, q)
    x3_out = _conv(x3, sd[p + ".layer3_outconv.weight"], 1, 0, q)
    if not fine:
        return x3_out, None
    x3_out_2x = F.interpolate(x3_out, scale_factor=2.0, mode="bilinear", align_corners=True)
    x2_out = _conv(x2, sd[p + ".layer2_outconv.weight"], 1, 0, q)
    t = x2_out + x3_out_2x
    t = F.leaky_relu(_bn(_conv(t, sd[p + ".layer2_outconv2.0.weight"], 1, 1, q), sd, p + ".layer2_outconv2.1"), 0.01)
    x2_out = _conv(t, sd[p + ".layer2_outconv2.3.weight"], 1, 1, q)
    x2_out_2x = F.interpolate(x2_out, scale_factor=2.0, mode="bilinear", align_corners=True)
    x1_out = _conv(x1, sd[p + ".layer1_outconv.weight"], 1, 0, q)
    t = x1_out + x2_out_2x
    t = F.leaky_relu(_bn(_conv(t, sd[p + ".layer1_outconv2.0.weight"], 1, 1, q), sd, p + ".layer1_outconv2.1"), 0.01)
    x1_out = _conv(t, sd[p + ".layer1_outconv2.3.weight"], 1, 1, q)
    return x3_out, x1_out


# ------------------------------------------------------------------- position encoding
def position_encoding_sine(d_model, h, w, temp_bug_fix=False):
    """utils/position_encoding.py:20-35.  Returns pe [d_model, h, w] (1-based positions).

    With temp_bug_fix=False (forced at coarse_match_worker.py:35) the exponent is
    ``-math.log(10000.0) / d_model // 2`` == floor(-0.03598) // ... == -1.0, i.e.
    div_term[k] = exp(-2k).
    """
    y_position = torch.ones(h, w).cumsum(0).float().unsqueeze(0)
    x_position = torch.ones(h, w).cumsum(1).float().unsqueeze(0)
    if temp_bug_fix:
        div_term = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / (d_model // 2)))
    else:
        div_term = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / d_model // 2))
    div_term = div_term[:, None, None]
    pe = torch.zeros(d_model, h, w)
    pe[0::4] = torch.sin(x_position * div_term)
    pe[1::4] = torch.cos(x_position * div_term)
    pe[2::4] = torch.sin(y_position * div_term)
    pe[3::4] = torch.cos(y_position * div_term)
    return pe


# --------------------------------------------------------------------------- transformer
def linear_attention(queries, keys, values, q_mask=None, kv_mask=None, eps=1e-6, q=_id):
    """loftr_module/linear_attention.py:20-47 (identical maths in
    src/MultiviewMatcher/matcher_module/linear_attention.py:28-60)."""
    Q = F.elu(queries) + 1
    K = F.elu(keys) + 1
    if q_mask is not None:
        Q = Q * q_mask[:, :, None, None]
    if kv_mask is not None:
        K = K * kv_mask[:, :, None, None]
        values = values * kv_mask[:, :, None, None]
    v_length = values.size(1)
    values = values / v_length
    KV = torch.einsum("nshd,nshv->nhdv", q(K), q(values))
    Z = 1 / (torch.einsum("nlhd,nhd->nlh", Q, K.sum(dim=1)) + eps)
    return (torch.einsum("nlhd,nhdv,nlh->nlhv", q(Q), q(KV), Z) * v_length).contiguous()


def encoder_layer(x, source, sd, p, nhead, x_mask=None, source_mask=None, q=_id):
    """loftr_module/transformer.py:35-58 LoFTREncoderLayer.forward."""
    bs, d = x.size(0), x.size(2)
    dim = d // nhead
    query = F.linear(q(x), q(sd[p + ".q_proj.weight"])).view(bs, -1, nhead, dim)
    key = F.linear(q(source), q(sd[p + ".k_proj.weight"])).view(bs, -1, nhead, dim)
    value = F.linear(q(source), q(sd[p + ".v_proj.weight"])).view(bs, -1, nhead, dim)
    message = linear_attention(query, key, value, x_mask, source_mask, q=q)
    message = F.linear(q(message.view(bs, -1, d)), q(sd[p + ".merge.weight"]))
    message = F.layer_norm(message, (d,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5)
    message = F.linear(q(torch.cat([x, message], dim=2)), q(sd[p + ".mlp.0.weight"]))
    message = F.linear(q(F.relu(message)), q(sd[p + ".mlp.2.weight"]))
    message = F.layer_norm(message, (d,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-5)
    return x + message


def local_feature_transformer(feat0, feat1, sd, p, layer_names, nhead, q=_id, taps=None):
    """loftr_module/transformer.py:80-101.  NOTE the cross layer updates feat0 first and
    feat1 then attends to the *updated* feat0 (:96-97)."""
    for i, name in enumerate(layer_names):
        lp = f"{p}.layers.{i}"
        if name == "self":
            feat0 = encoder_layer(feat0, feat0, sd, lp, nhead, q=q)
            feat1 = encoder_layer(feat1, feat1, sd, lp, nhead, q=q)
        elif name == "cross":
            feat0 = encoder_layer(feat0, feat1, sd, lp, nhead, q=q)
            feat1 = encoder_layer(feat1, feat0, sd, lp, nhead, q=q)
        else:
            raise KeyError(name)
        if taps is not None:
            taps.append((feat0, feat1))
    return feat0, feat1


# ----------------------------------------------------------------------- coarse matching
def dual_softmax_conf(feat_c0, feat_c1, temperature, q=_id):
    """utils/coarse_matching.py:100-116 (dual_softmax branch, no masks)."""
    feat_c0, feat_c1 = feat_c0 / feat_c0.shape[-1] ** .5, feat_c1 / feat_c1.shape[-1] ** .5
    sim = torch.einsum("nlc,nsc->nls", q(feat_c0), q(feat_c1)) / temperature
    return F.softmax(sim, 1) * F.softmax(sim, 2)


def get_coarse_match(conf, hw0_c, hw1_c, hw0_i, thr, border_rm, scale0=None, scale1=None):
    """utils/coarse_matching.py:148-258 (inference branch, no padding masks).

    ``mask_border`` (:8-22) only zeroes the LEADING ``b`` rows/cols of each of the four
    grid axes: its trailing ``-b:0`` slices are empty.
    """
    N = conf.size(0)
    h0c, w0c = hw0_c
    h1c, w1c = hw1_c
    mask = (conf > thr).view(N, h0c, w0c, h1c, w1c).clone()
    b = border_rm
    if b > 0:
        mask[:, :b] = False
        mask[:, :, :b] = False
        mask[:, :, :, :b] = False
        mask[:, :, :, :, :b] = False
    mask = mask.view(N, h0c * w0c, h1c * w1c)
    mask = mask * (conf == conf.max(dim=2, keepdim=True)[0]) * (conf == conf.max(dim=1, keepdim=True)[0])
    mask_v, all_j_ids = mask.max(dim=2)
    b_ids, i_ids = torch.where(mask_v)
    j_ids = all_j_ids[b_ids, i_ids]
    mconf = conf[b_ids, i_ids, j_ids]
    scale = hw0_i[0] / hw0_c[0]
    s0 = scale * scale0[b_ids][:, [1, 0]] if scale0 is not None else scale
    s1 = scale * scale1[b_ids][:, [1, 0]] if scale1 is not None else scale
    mkpts0_c = torch.stack([i_ids % w0c, i_ids // w0c], dim=1) * s0
    mkpts1_c = torch.stack([j_ids % w1c, j_ids // w1c], dim=1) * s1
    keep = mconf != 0
    return {"b_ids": b_ids, "i_ids": i_ids, "j_ids": j_ids, "gt_mask": mconf == 0,
            "m_bids": b_ids[keep], "mkpts0_c": mkpts0_c[keep], "mkpts1_c": mkpts1_c[keep],
            "mconf": mconf[keep]}


# ---------------------------------------------------------------------------- fine stage
def create_meshgrid(h, w):
    """kornia 0.4.1 ``create_meshgrid(h, w, normalized_coordinates=True)`` (un-vendored
    dependency, requirements.txt:8): [1,h,w,2], [...,0]=x in [-1,1], [...,1]=y."""
    xs = (torch.linspace(0, w - 1, w) / (w - 1) - 0.5) * 2
    ys = (torch.linspace(0, h - 1, h) / (h - 1) - 0.5) * 2
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack([gx, gy], -1)[None]


def spatial_expectation2d(heat):
    """kornia 0.4.1 ``dsnt.spatial_expectation2d(input[B,N,h,w], True)`` -> [B,N,2]."""
    B, N, h, w = heat.shape
    grid = create_meshgrid(h, w).to(heat)
    pos_x = grid[..., 0].reshape(-1)
    pos_y = grid[..., 1].reshape(-1)
    flat = heat.reshape(B, N, -1)
    ex = torch.sum(pos_x * flat, -1, keepdim=True)
    ey = torch.sum(pos_y * flat, -1, keepdim=True)
    return torch.cat([ex, ey], -1)


def fine_preprocess(feat_f0, feat_f1, feat_c0, feat_c1, cm, sd, W, stride, q=_id):
    """loftr_module/fine_preprocess.py:29-59 (restated as a gather of the M selected WxW
    windows instead of unfold-then-select; zero padding W//2 like F.unfold)."""
    b_ids, i_ids, j_ids = cm["b_ids"], cm["i_ids"], cm["j_ids"]
    if b_ids.shape[0] == 0:
        e = torch.empty(0, W * W, feat_f0.shape[1])
        return e, e.clone()

    def gather(feat_f, ids):
        n, c, hf, wf = feat_f.shape
        wc = wf // stride
        fp = F.pad(feat_f, (W // 2,) * 4)
        out = []
        for b, i in zip(b_ids.tolist(), ids.tolist()):
            y, x = (i // wc) * stride, (i % wc) * stride
            out.append(fp[b, :, y:y + W, x:x + W].reshape(c, W * W).t())
        return torch.stack(out)  # [M, WW, C]

    f0, f1 = gather(feat_f0, i_ids), gather(feat_f1, j_ids)
    c_win = F.linear(q(torch.cat([feat_c0[b_ids, i_ids], feat_c1[b_ids, j_ids]], 0)),
                     q(sd["fine_preprocess.down_proj.weight"]), sd["fine_preprocess.down_proj.bias"])
    cat = torch.cat([torch.cat([f0, f1], 0), c_win[:, None].expand(-1, W * W, -1)], -1)
    cf = F.linear(q(cat), q(sd["fine_preprocess.merge_feat.weight"]), sd["fine_preprocess.merge_feat.bias"])
    return torch.chunk(cf, 2, dim=0)


def fine_matching(feat_f0, feat_f1, cm, hw0_i, hw0_f, scale1=None, q=_id):
    """utils/fine_matching.py:15-74."""
    M, WW, C = feat_f0.shape
    W = int(math.sqrt(WW))
    scale = hw0_i[0] / hw0_f[0]
    if M == 0:
        return {"expec_f": torch.empty(0, 3), "mkpts0_f": cm["mkpts0_c"], "mkpts1_f": cm["mkpts1_c"]}
    picked = feat_f0[:, WW // 2, :]
    sim = torch.einsum("mc,mrc->mr", q(picked), q(feat_f1))
    heat = torch.softmax(sim / C ** .5, dim=1).view(-1, W, W)
    coords = spatial_expectation2d(heat[None])[0]
    grid = create_meshgrid(W, W).reshape(1, -1, 2)
    var = torch.sum(grid ** 2 * heat.view(-1, WW, 1), dim=1) - coords ** 2
    std = torch.sum(torch.sqrt(torch.clamp(var, min=1e-10)), -1)
    s1 = scale * scale1[cm["b_ids"]][:, [1, 0]] if scale1 is not None else scale
    mkpts1_f = cm["mkpts1_c"] + (coords * (W // 2) * s1)[:len(cm["mconf"])]
    return {"expec_f": torch.cat([coords, std[:, None]], -1), "mkpts0_f": cm["mkpts0_c"], "mkpts1_f": mkpts1_f}


# ------------------------------------------------------------------------------ forward
@torch.no_grad()
def loftr_forward(data, sd, cfg=None, q=None, keep=False):
    """loftr.py:29-81 LoFTR.forward.  ``data`` holds image0/image1 [N,1,H,W] and optional
    scale0/scale1 [N,2]; returns a NEW dict with the keys the reference adds (the
    reference mutates ``data`` in place; the caller may ``data.update(out)``)."""
    cfg = {**DEFAULT_CFG, **(cfg or {})}
    q = q or _id
    im0, im1 = data["image0"], data["image1"]
    out = {"bs": im0.size(0), "hw0_i": tuple(im0.shape[2:]), "hw1_i": tuple(im1.shape[2:])}
    fine = cfg["fine_enable"]
    run_fpn = fine or cfg["compute_unused_fine_branch"]
    if out["hw0_i"] == out["hw1_i"]:
        fc, ff = resnet_fpn_8_2(torch.cat([im0, im1], 0), sd, run_fpn, q)
        feat_c0, feat_c1 = fc.split(out["bs"])
        feat_f0, feat_f1 = ff.split(out["bs"]) if fine else (None, None)
    else:
        feat_c0, feat_f0 = resnet_fpn_8_2(im0, sd, run_fpn, q)
        feat_c1, feat_f1 = resnet_fpn_8_2(im1, sd, run_fpn, q)
    out["hw0_c"], out["hw1_c"] = tuple(feat_c0.shape[2:]), tuple(feat_c1.shape[2:])
    if keep:
        out["backbone_c0"], out["backbone_c1"] = feat_c0, feat_c1
    d = cfg["d_model"]
    f0 = (feat_c0 + position_encoding_sine(d, *out["hw0_c"])[None]).flatten(2).transpose(1, 2)
    f1 = (feat_c1 + position_encoding_sine(d, *out["hw1_c"])[None]).flatten(2).transpose(1, 2)
    taps = [] if keep else None
    f0, f1 = local_feature_transformer(f0, f1, sd, "loftr_coarse", cfg["layer_names"], cfg["nhead"], q, taps)
    if keep:
        out["feat_c0"], out["feat_c1"], out["layer_taps"] = f0, f1, taps
    conf = dual_softmax_conf(f0, f1, cfg["temperature"], q)
    if keep:
        out["conf_matrix"] = conf
    cm = get_coarse_match(conf, out["hw0_c"], out["hw1_c"], out["hw0_i"], cfg["thr"], cfg["border_rm"],
                          data.get("scale0"), data.get("scale1"))
    out.update(cm)
    if not fine:
        out["mkpts0_f"], out["mkpts1_f"] = cm["mkpts0_c"], cm["mkpts1_c"]
        return out
    out["hw0_f"], out["hw1_f"] = tuple(feat_f0.shape[2:]), tuple(feat_f1.shape[2:])
    W = cfg["fine_window"]
    stride = out["hw0_f"][0] // out["hw0_c"][0]
    u0, u1 = fine_preprocess(feat_f0, feat_f1, f0, f1, cm, sd, W, stride, q)
    if u0.size(0) != 0:
        u0, u1 = local_feature_transformer(u0, u1, sd, "loftr_fine", cfg["fine_layer_names"], cfg["fine_nhead"], q)
    out.update(fine_matching(u0, u1, cm, out["hw0_i"], out["hw0_f"], data.get("scale1"), q))
    return out
