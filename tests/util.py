"""Seeded synthetic inputs shared by tests, smoke() and bench.py."""
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
