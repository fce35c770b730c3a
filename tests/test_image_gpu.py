"""SURVEY 8(f) row 3 on the GPU: csrc/image_ops.cu behind dfsfm_resize_lanczos_gray and the host mirror of read_grayscale /
CoarseMatchingDataset vs PIL itself (the reference's resize) and the CPU oracle -- bit-exact (uint8 arithmetic, one float32
division)."""
import os

import numpy as np
import pytest
import torch

from tests import util  # noqa: E402

from oracle import image_oracle as io

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def reader():
    from detectorfreesfm_b200.image_pipeline import GpuImageReader
    return GpuImageReader()


@pytest.mark.parametrize("H,W,oh,ow", [(480, 640, 360, 480), (100, 37, 64, 24), (33, 50, 99, 120), (64, 64, 64, 32), (64, 64, 40, 64),
                                       (17, 17, 17, 17), (5, 7, 1, 1), (200, 300, 208, 304), (3000, 4000, 624, 832), (880, 1200, 880, 1200)])
def test_resize_matches_pil(reader, H, W, oh, ow):
    from PIL import Image
    img = util.synth_photo(H, W, seed=H + W)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=Image.LANCZOS)).astype("float32") / 255.
    out = reader.resize_gray(img, (ow, oh)).cpu().numpy()
    assert out.shape == ref.shape and out.dtype == np.float32
    assert np.array_equal(out, ref), f"max diff {np.abs(out - ref).max() * 255:.0f} grey levels"
    # a tensor input (already on the device) takes the same path
    out2 = reader.resize_gray(torch.from_numpy(img).cuda(), (ow, oh)).cpu().numpy()
    assert np.array_equal(out2, ref)


def test_read_grayscale_matches_oracle_and_golden(reader, tmp_path):
    import cv2
    for seed, (h, w, resize, df) in enumerate([(150, 200, (96,), 8), (97, 61, (128,), 8), (64, 80, None, None), (300, 200, (64, 48), None)]):
        img = util.synth_photo(h, w, seed)
        path = str(tmp_path / f"im{seed}.png")
        assert cv2.imwrite(path, img)
        t, s, hw = reader.read_grayscale(path, resize, df=df, ret_scales=True)
        to, so, ho = io.read_grayscale_from_array(img, resize, df=df)
        assert t.is_cuda and t.dtype == torch.float32 and torch.equal(t.cpu(), to) and torch.equal(s, so) and torch.equal(hw, ho)
    for i, g in enumerate(torch.load(os.path.join(GOLD, "image_small.pt"), weights_only=False)):   # outputs of the reference itself
        path = str(tmp_path / f"g{i}.png")
        assert cv2.imwrite(path, g["image"].numpy())
        t, s, hw = reader.read_grayscale(path, g["resize"], df=g["df"], ret_scales=True)
        assert torch.equal(t.cpu(), g["tensor"]) and torch.equal(s, g["scales"]) and torch.equal(hw, g["original_hw"])
    # pad_to: zeros to the bottom / right, mask of the valid area (utils.py:33-52)
    img = util.synth_photo(90, 120, 5)
    path = str(tmp_path / "pad.png")
    cv2.imwrite(path, img)
    t, s, hw, mask = reader.read_grayscale(path, (64,), df=8, pad_to=-1, ret_scales=True, ret_pad_mask=True)
    to, _, _ = io.read_grayscale_from_array(img, (64,), df=8)
    assert t.shape == (1, 64, 64) and torch.equal(t[0, :48, :64].cpu(), to[0]) and float(t[0, 48:].abs().sum()) == 0.0
    assert float(mask.sum()) == 48 * 64
    with pytest.raises(FileNotFoundError):
        reader.read_grayscale(str(tmp_path / "missing.png"), (64,), df=8)


def test_dataset_mirror_caches_and_feeds_the_matcher(tmp_path):
    import cv2
    from detectorfreesfm_b200.image_pipeline import B200CoarseMatchingDataset
    paths = []
    imgs = []
    for i in range(3):
        img = util.synth_photo(120 + 8 * i, 160, 10 + i)
        p = str(tmp_path / f"scene_{i}.png")
        cv2.imwrite(p, img)
        paths.append(p)
        imgs.append(img)
    pairs = [f"{paths[a]} {paths[b]}" for a, b in ((0, 1), (0, 2), (1, 2))]
    args = {"img_resize": 128, "df": 8, "pad_to": None, "img_preload": False, "img_type": "grayscale"}
    ds = B200CoarseMatchingDataset(args, paths, pairs, subset_ids=[0, 1, 2])
    loader = torch.utils.data.DataLoader(ds, batch_size=1, shuffle=False, num_workers=0)
    seen = 0
    for item, (a, b) in zip(loader, ((0, 1), (0, 2), (1, 2))):
        for side, idx in (("0", a), ("1", b)):
            to, so, _ = io.read_grayscale_from_array(imgs[idx], (128,), df=8)
            assert item["image" + side].is_cuda and torch.equal(item["image" + side][0].cpu(), to)
            assert torch.equal(item["scale" + side][0], so)
            assert item["f_name" + side][0] == f"scene_{idx}"
        assert item["pair_key"][0][0] == paths[a] and item["pair_key"][1][0] == paths[b] and int(item["frameID"][0]) == seen
        seen += 1
    assert ds.decodes == 3          # three images, six uses: every image decoded and resized once
    # preload fills the cache up front, like img_preload in the reference
    ds2 = B200CoarseMatchingDataset(dict(args, img_preload=True), paths, pairs, subset_ids=[2])
    assert ds2.decodes == 3 and torch.equal(ds2[0]["image0"].cpu(), io.read_grayscale_from_array(imgs[1], (128,), df=8)[0])
    with pytest.raises(NotImplementedError):
        B200CoarseMatchingDataset(dict(args, img_type="rgb"), paths, pairs, subset_ids=[0])


@pytest.mark.parametrize("H,W,resize,df", [(150, 200, (96,), 8), (97, 61, (128,), 8), (300, 420, (256,), 8), (64, 80, None, None)])
def test_read_rgb_matches_pil_rgb_resample(reader, tmp_path, H, W, resize, df):
    """HP-2's loader (src/dataset/utils.py:80-118, called with resize_no_larger_than=True at coarse_sfm_refinement_dataset.py:372-380):
    cv2 decode -> RGB -> PIL LANCZOS on the 3-band image -> /255 -> [3,h,w]; the GPU path runs the grayscale kernel per band."""
    import cv2
    from PIL import Image
    from detectorfreesfm_b200.image_pipeline import process_resize
    rgb = np.stack([util.synth_photo(H, W, seed=H + W + c) for c in range(3)], -1)
    path = str(tmp_path / "rgb.png")
    assert cv2.imwrite(path, cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR))
    t, scales, hw = reader.read_rgb(path, resize, resize_no_larger_than=True, df=df, ret_scales=True)
    w_new, h_new = process_resize(W, H, resize if resize is not None else (W, H), df, resize_no_larger_than=True)
    ref = np.asarray(Image.fromarray(rgb).resize((w_new, h_new), resample=Image.LANCZOS), dtype=np.uint8).astype("float32")
    ref = torch.from_numpy(ref / 255.).float().permute(2, 0, 1).contiguous()
    assert t.is_cuda and t.shape == ref.shape and torch.equal(t.cpu(), ref)
    assert torch.equal(scales, torch.tensor([float(H) / float(h_new), float(W) / float(w_new)])) and torch.equal(hw, torch.tensor([H, W]))
