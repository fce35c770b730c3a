"""Seeded synthetic weights with the reference's state_dict key names and shapes.

TEST INFRASTRUCTURE (see oracle/__init__.py).  No pretrained checkpoint is available
offline, so parity is checked on seeded random weights that exercise every code path:
BatchNorm gets non-trivial running stats / affine terms (so that BN folding is tested)
and biased layers get non-zero biases.

Shapes: SURVEY.md appendix B (read off the instantiated reference modules).

``loftr_state_dict(seed, calibrated=True)`` additionally sets every backbone BatchNorm's running statistics to the
statistics of its own input on a seeded calibration image -- what training does to a real checkpoint.  With arbitrary
running stats a random ReLU network collapses onto a handful of directions (participation ratio of the 1/8 features
~4 of 256) and the dual-softmax never exceeds 1e-3; with data-matched statistics the features stay high-dimensional
(~110) and the default thr 0.2 / temperature 0.1 yields thousands of mutual-NN matches on overlapping views, like the
real model.  Pure generator code (plain torch ops, no oracle import): bench.py's product arm uses it.
"""
import torch
import torch.nn.functional as F


def _conv_w(g, cout, cin, k, gain=2.0):
    # kaiming_normal_(mode='fan_out', relu) like backbone/resnet_fpn.py:77-79
    std = (gain / (cout * k * k)) ** 0.5
    return torch.randn(cout, cin, k, k, generator=g) * std


def _lin_w(g, cout, cin):
    # xavier_uniform_ like loftr_module/transformer.py:75-78
    a = (6.0 / (cin + cout)) ** 0.5
    return (torch.rand(cout, cin, generator=g) * 2 - 1) * a


def _bn(g, sd, p, c):
    sd[p + ".weight"] = 0.75 + 0.5 * torch.rand(c, generator=g)
    sd[p + ".bias"] = 0.1 * torch.randn(c, generator=g)
    sd[p + ".running_mean"] = 0.1 * torch.randn(c, generator=g)
    sd[p + ".running_var"] = 0.5 + torch.rand(c, generator=g)
    sd[p + ".num_batches_tracked"] = torch.tensor(1)


def _encoder(g, sd, p, d):
    for n in ("q_proj", "k_proj", "v_proj", "merge"):
        sd[f"{p}.{n}.weight"] = _lin_w(g, d, d)
    sd[f"{p}.mlp.0.weight"] = _lin_w(g, 2 * d, 2 * d)
    sd[f"{p}.mlp.2.weight"] = _lin_w(g, d, 2 * d)
    for n in ("norm1", "norm2"):
        sd[f"{p}.{n}.weight"] = 0.75 + 0.5 * torch.rand(d, generator=g)
        sd[f"{p}.{n}.bias"] = 0.1 * torch.randn(d, generator=g)


def calibrate_backbone_bn(sd, image, p="backbone"):
    """Returns a copy of ``sd`` whose ResNetFPN_8_2 BatchNorms (coarse sub-graph, resnet_fpn.py:100-108) carry the
    per-channel mean / biased variance of their own inputs on ``image`` [1,1,H,W], computed layer by layer."""
    sd = dict(sd)

    def bn(x, name):
        sd[name + ".running_mean"] = x.mean((0, 2, 3))
        sd[name + ".running_var"] = x.var((0, 2, 3), unbiased=False)
        return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"], sd[name + ".bias"], False, 0.0, 1e-5)

    def block(x, q, stride):
        y = F.relu(bn(F.conv2d(x, sd[q + ".conv1.weight"], None, stride, 1), q + ".bn1"))
        y = bn(F.conv2d(y, sd[q + ".conv2.weight"], None, 1, 1), q + ".bn2")
        if stride != 1:
            x = bn(F.conv2d(x, sd[q + ".downsample.0.weight"], None, stride, 0), q + ".downsample.1")
        return F.relu(x + y)

    x = F.relu(bn(F.conv2d(image, sd[p + ".conv1.weight"], None, 2, 3), p + ".bn1"))
    for li, stride in ((1, 1), (2, 2), (3, 2)):
        x = block(x, f"{p}.layer{li}.0", stride)
        x = block(x, f"{p}.layer{li}.1", 1)
    return sd


_CAL_CACHE = {}


def loftr_state_dict(seed=0, calibrated=False):
    """LoFTR (outdoor_ds layout): third_party/LoFTR/src/loftr/loftr.py:12-27."""
    if calibrated:
        if seed not in _CAL_CACHE:
            g = torch.Generator().manual_seed(1000 + seed)
            cal = F.avg_pool2d(torch.rand(1, 1, 260, 260, generator=g), 5, 1)
            cal = (cal - cal.min()) / (cal.max() - cal.min())
            with torch.no_grad():
                _CAL_CACHE[seed] = calibrate_backbone_bn(loftr_state_dict(seed), cal)
        return dict(_CAL_CACHE[seed])
    g = torch.Generator().manual_seed(seed)
    sd = {}
    dims = [128, 196, 256]
    sd["backbone.conv1.weight"] = _conv_w(g, 128, 1, 7)
    _bn(g, sd, "backbone.bn1", 128)
    cin = 128
    for li, d in enumerate(dims, start=1):
        for bi in range(2):
            p = f"backbone.layer{li}.{bi}"
            sd[p + ".conv1.weight"] = _conv_w(g, d, cin if bi == 0 else d, 3)
            sd[p + ".conv2.weight"] = _conv_w(g, d, d, 3)
            _bn(g, sd, p + ".bn1", d)
            _bn(g, sd, p + ".bn2", d)
            if bi == 0 and li > 1:
                sd[p + ".downsample.0.weight"] = _conv_w(g, d, cin, 1)
                _bn(g, sd, p + ".downsample.1", d)
        cin = d
    sd["backbone.layer3_outconv.weight"] = _conv_w(g, 256, 256, 1)
    sd["backbone.layer2_outconv.weight"] = _conv_w(g, 256, 196, 1)
    sd["backbone.layer2_outconv2.0.weight"] = _conv_w(g, 256, 256, 3)
    _bn(g, sd, "backbone.layer2_outconv2.1", 256)
    sd["backbone.layer2_outconv2.3.weight"] = _conv_w(g, 196, 256, 3)
    sd["backbone.layer1_outconv.weight"] = _conv_w(g, 196, 128, 1)
    sd["backbone.layer1_outconv2.0.weight"] = _conv_w(g, 196, 196, 3)
    _bn(g, sd, "backbone.layer1_outconv2.1", 196)
    sd["backbone.layer1_outconv2.3.weight"] = _conv_w(g, 128, 196, 3)
    for i in range(8):
        _encoder(g, sd, f"loftr_coarse.layers.{i}", 256)
    sd["fine_preprocess.down_proj.weight"] = _lin_w(g, 128, 256)
    sd["fine_preprocess.down_proj.bias"] = 0.05 * torch.randn(128, generator=g)
    sd["fine_preprocess.merge_feat.weight"] = _lin_w(g, 128, 256)
    sd["fine_preprocess.merge_feat.bias"] = 0.05 * torch.randn(128, generator=g)
    for i in range(2):
        _encoder(g, sd, f"loftr_fine.layers.{i}", 128)
    return sd


VGG_CONVS = [(0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256), (14, 256, 256)]


def multiview_state_dict(seed=0):
    """MultiviewMatcher (S2DNet backbone + 4-layer d=128 transformer):
    src/MultiviewMatcher/MultiviewMatcher.py:17-40, backbone/S2DNet/s2dnet.py:24-110."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for idx, cin, cout in VGG_CONVS:
        sd[f"backbone.encoder.{idx}.weight"] = _conv_w(g, cout, cin, 3)
        sd[f"backbone.encoder.{idx}.bias"] = 0.05 * torch.randn(cout, generator=g)
    for i, cin in enumerate((64, 256)):
        p = f"backbone.adaptation_layers.adap_layer_{i}"
        sd[p + ".0.weight"] = _conv_w(g, 64, cin, 1)
        sd[p + ".0.bias"] = 0.05 * torch.randn(64, generator=g)
        sd[p + ".2.weight"] = _conv_w(g, 128, 64, 5)
        sd[p + ".2.bias"] = 0.05 * torch.randn(128, generator=g)
        _bn(g, sd, p + ".3", 128)
    for i in range(4):
        _encoder(g, sd, f"fine_transformer.layers.{i}", 128)
    return sd
