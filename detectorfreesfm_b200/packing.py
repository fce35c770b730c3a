"""state_dict -> packed engine parameters (BatchNorm folded, channel padding, tap-major K layout).

Conv weights [cout, cin, kh, kw] become GEMM operands W[n, t*cpad + c] with tap t = ky*kw + kx, rows padded to
`npad`, channels to `cpad` (zeros); eval-mode BatchNorm (eps 1e-5) is folded in float64:
    w' = w * gamma / sqrt(var + eps),   b' = beta - mean * gamma / sqrt(var + eps)  (+ conv bias * scale).
Key names follow the reference checkpoints:
  LoFTR:            third_party/LoFTR/src/loftr/loftr.py:12-27, backbone/resnet_fpn.py:42-75
  MultiviewMatcher: src/MultiviewMatcher/backbone/S2DNet/s2dnet.py:24-110, matcher_module/transformer.py:8-60
"""
import torch


def _fold_bn(sd, p, eps=1e-5):
    g = sd[p + ".weight"].double()
    b = sd[p + ".bias"].double()
    m = sd[p + ".running_mean"].double()
    v = sd[p + ".running_var"].double()
    s = g / torch.sqrt(v + eps)
    return s, b - m * s


def conv_matrix(w, scale=None, npad=None, cpad=None):
    """[cout,cin,kh,kw] -> float64 [npad, kh*kw*cpad]."""
    cout, cin, kh, kw = w.shape
    npad = npad or cout
    cpad = cpad or cin
    w = w.double()
    if scale is not None:
        w = w * scale.view(-1, 1, 1, 1)
    m = torch.zeros(npad, kh * kw, cpad, dtype=torch.float64)
    m[:cout, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    return m.reshape(npad, kh * kw * cpad)


def _padvec(v, n):
    out = torch.zeros(n, dtype=torch.float64)
    out[:v.numel()] = v.double()
    return out


def pack_loftr(sd, fine=False):
    """-> dict name -> (tensor fp32 2-D, kind) for dfsfm_coarse_set_param (kind 0 = GEMM operand, 1 = fp32)."""
    sd = {k.replace("matcher.", "", 1) if k.startswith("matcher.") else k: v for k, v in sd.items()}
    out = {}

    def put(name, t, kind):
        t = t.float().contiguous()
        if t.dim() == 1:
            t = t.view(1, -1)
        out[name] = (t, kind)

    s, b = _fold_bn(sd, "backbone.bn1")
    put("stem.w", (sd["backbone.conv1.weight"].double() * s.view(-1, 1, 1, 1)).reshape(128, 49), 1)
    put("stem.b", b, 1)
    pad = {128: 128, 196: 208, 256: 256}
    cin = 128
    for li, d in ((1, 128), (2, 196), (3, 256)):
        for bi in range(2):
            p = f"backbone.layer{li}.{bi}"
            c_in = cin if bi == 0 else d
            s1, b1 = _fold_bn(sd, p + ".bn1")
            put(f"l{li}.{bi}.c1.w", conv_matrix(sd[p + ".conv1.weight"], s1, pad[d], pad[c_in]), 0)
            put(f"l{li}.{bi}.c1.b", _padvec(b1, pad[d]), 1)
            s2, b2 = _fold_bn(sd, p + ".bn2")
            m2 = conv_matrix(sd[p + ".conv2.weight"], s2, pad[d], pad[d])
            bias2 = _padvec(b2, pad[d])
            if bi == 0 and li > 1:
                # relu(bn2(conv2(y)) + bn_ds(conv_ds(x)))  ==  one GEMM with the 1x1 stride-2 downsample as a 10th tap
                sds, bds = _fold_bn(sd, p + ".downsample.1")
                mds = conv_matrix(sd[p + ".downsample.0.weight"], sds, pad[d], pad[d])  # cin padded up to the tap width
                m2 = torch.cat([m2, mds], dim=1)
                bias2 = bias2 + _padvec(bds, pad[d])
            put(f"l{li}.{bi}.c2.w", m2, 0)
            put(f"l{li}.{bi}.c2.b", bias2, 1)
        cin = d
    put("out3.w", conv_matrix(sd["backbone.layer3_outconv.weight"]), 0)
    if fine:
        # FPN top-down path (resnet_fpn.py:56-73): 1x1 laterals on the stored (padded) channel counts, 3x3 + BN + LeakyReLU, 3x3
        put("fpn.l2o.w", conv_matrix(sd["backbone.layer2_outconv.weight"], None, 256, 208), 0)
        s_, b_ = _fold_bn(sd, "backbone.layer2_outconv2.1")
        put("fpn.l2o2a.w", conv_matrix(sd["backbone.layer2_outconv2.0.weight"], s_, 256, 256), 0)
        put("fpn.l2o2a.b", b_, 1)
        put("fpn.l2o2b.w", conv_matrix(sd["backbone.layer2_outconv2.3.weight"], None, 208, 256), 0)
        put("fpn.l1o.w", conv_matrix(sd["backbone.layer1_outconv.weight"], None, 208, 128), 0)
        s_, b_ = _fold_bn(sd, "backbone.layer1_outconv2.1")
        put("fpn.l1o2a.w", conv_matrix(sd["backbone.layer1_outconv2.0.weight"], s_, 208, 208), 0)
        put("fpn.l1o2a.b", _padvec(b_, 208), 1)
        put("fpn.l1o2b.w", conv_matrix(sd["backbone.layer1_outconv2.3.weight"], None, 128, 208), 0)
        # FinePreprocess (fine_preprocess.py:18-20): merge_feat acts on cat[window (128), down_proj(coarse) (128)]
        put("fine.down.w", sd["fine_preprocess.down_proj.weight"], 0)
        put("fine.down.b", sd["fine_preprocess.down_proj.bias"], 1)
        put("fine.merge_f.w", sd["fine_preprocess.merge_feat.weight"][:, :128], 0)
        put("fine.merge_c.w", sd["fine_preprocess.merge_feat.weight"][:, 128:], 0)
        put("fine.merge.b", sd["fine_preprocess.merge_feat.bias"], 1)
        for i in range(2):
            p = f"loftr_fine.layers.{i}"
            put(f"fine.tr.{i}.qkv", torch.cat([sd[p + ".q_proj.weight"], sd[p + ".k_proj.weight"], sd[p + ".v_proj.weight"]], 0), 0)
            put(f"fine.tr.{i}.merge", sd[p + ".merge.weight"], 0)
            put(f"fine.tr.{i}.mlp0", sd[p + ".mlp.0.weight"], 0)
            put(f"fine.tr.{i}.mlp2", sd[p + ".mlp.2.weight"], 0)
            for n in ("1", "2"):
                put(f"fine.tr.{i}.ln{n}.g", sd[p + f".norm{n}.weight"], 1)
                put(f"fine.tr.{i}.ln{n}.b", sd[p + f".norm{n}.bias"], 1)
    for i in range(8):
        p = f"loftr_coarse.layers.{i}"
        put(f"tr.{i}.qkv", torch.cat([sd[p + ".q_proj.weight"], sd[p + ".k_proj.weight"], sd[p + ".v_proj.weight"]], 0), 0)
        # k/v projection with the state reduction in the GEMM epilogue (KvEpi): each 256-row weight tile = [K rows of heads 4t..4t+3 |
        # V rows of the same heads], so that one output tile holds both factors of its heads' K^T V products
        wk, wv = sd[p + ".k_proj.weight"], sd[p + ".v_proj.weight"]
        put(f"tr.{i}.kvp", torch.cat([wk[:128], wv[:128], wk[128:], wv[128:]], 0), 0)
        put(f"tr.{i}.merge", sd[p + ".merge.weight"], 0)
        put(f"tr.{i}.mlp0", sd[p + ".mlp.0.weight"], 0)
        put(f"tr.{i}.mlp2", sd[p + ".mlp.2.weight"], 0)
        for n in ("1", "2"):
            put(f"tr.{i}.ln{n}.g", sd[p + f".norm{n}.weight"], 1)
            put(f"tr.{i}.ln{n}.b", sd[p + f".norm{n}.bias"], 1)
    return out


def position_encoding(h, w, d_model=256):
    """PositionEncodingSine(d_model, temp_bug_fix=False) as [h*w, d_model] fp32 tokens
    (third_party/LoFTR/src/loftr/utils/position_encoding.py:20-35; coarse_match_worker.py:35 forces the legacy
    formula: div_term = exp(arange(0, d/2, 2) * (-ln(1e4) / d // 2)) = exp(-k), k = 0, 2, 4, ...).
    Built with the same torch ops as the reference so the table is bit-identical to its buffer."""
    import math
    y_position = torch.ones(h, w).cumsum(0).float().unsqueeze(0)
    x_position = torch.ones(h, w).cumsum(1).float().unsqueeze(0)
    div_term = torch.exp(torch.arange(0, d_model // 2, 2).float() * (-math.log(10000.0) / d_model // 2))[:, None, None]
    pe = torch.zeros(d_model, h, w)
    pe[0::4] = torch.sin(x_position * div_term)
    pe[1::4] = torch.cos(x_position * div_term)
    pe[2::4] = torch.sin(y_position * div_term)
    pe[3::4] = torch.cos(y_position * div_term)
    return pe.flatten(1).t().contiguous()


VGG_CONVS = ((0, "c11"), (2, "c12"), (5, "c21"), (7, "c22"), (10, "c31"), (12, "c32"), (14, "c33"))


def pack_multiview(sd):
    """MultiviewMatcher checkpoint -> engine parameters for dfsfm_refine_set_param.  Accepts the keys after
    multiview_match_worker.py:42-52 (``matcher.`` stripped, loftr_fine -> fine_transformer) or the raw checkpoint keys."""
    fixed = {}
    for k, v in sd.items():
        if k.startswith("matcher."):
            k = k[len("matcher."):]
        if "loftr_coarse" in k:
            continue
        fixed[k.replace("loftr_fine", "fine_transformer")] = v
    sd = fixed
    out = {}

    def put(name, t, kind):
        t = t.float().contiguous()
        if t.dim() == 1:
            t = t.view(1, -1)
        out[name] = (t, kind)

    for idx, name in VGG_CONVS:
        w = sd[f"backbone.encoder.{idx}.weight"]
        b = sd[f"backbone.encoder.{idx}.bias"]
        if name == "c11":
            put("c11.w", w.reshape(64, 27), 1)   # [cout][cin*3*3] for the fused RoIAlign + conv1_1 kernel
        else:
            put(name + ".w", conv_matrix(w), 0)
        put(name + ".b", b, 1)
    for i in range(2):
        a = f"backbone.adaptation_layers.adap_layer_{i}"
        put(f"a{i}.0.w", conv_matrix(sd[a + ".0.weight"]), 0)
        put(f"a{i}.0.b", sd[a + ".0.bias"], 1)
        s, b = _fold_bn(sd, a + ".3")
        put(f"a{i}.2.w", conv_matrix(sd[a + ".2.weight"], s), 0)
        put(f"a{i}.2.b", sd[a + ".2.bias"].double() * s + b, 1)
    for i in range(4):
        p = f"fine_transformer.layers.{i}"
        put(f"tr.{i}.qkv", torch.cat([sd[p + ".q_proj.weight"], sd[p + ".k_proj.weight"], sd[p + ".v_proj.weight"]], 0), 0)
        put(f"tr.{i}.merge", sd[p + ".merge.weight"], 0)
        put(f"tr.{i}.mlp0", sd[p + ".mlp.0.weight"], 0)
        put(f"tr.{i}.mlp2", sd[p + ".mlp.2.weight"], 0)
        for n in ("1", "2"):
            put(f"tr.{i}.ln{n}.g", sd[p + f".norm{n}.weight"], 1)
            put(f"tr.{i}.ln{n}.b", sd[p + f".norm{n}.bias"], 1)
    return out
