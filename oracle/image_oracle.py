"""CPU oracle for the host image pipeline.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

* ``read_grayscale_from_array`` restates src/dataset/utils.py:121-159 after the decode (process_resize :14-30, resize_image
  :161-177 with "pil_LANCZOS", grayscale2tensor :56-57) and, like the reference, calls PIL for the resize -- Pillow is the
  reference's own third-party dependency (this image: Pillow 12.2.0), not part of /root/reference.
* ``resample_8bpc`` / ``coeffs`` restate Pillow's published algorithm (src/libImaging/Resample.c: precompute_coeffs,
  normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc, ImagingResampleVertical_8bpc, lanczos_filter) in numpy; the tests pin
  it bit-for-bit against PIL itself over up/down-scaling, identity and degenerate sizes, so the GPU kernels are checked against
  two independent statements of the same arithmetic.

Pinned against the reference's own ``read_grayscale`` (imported behind stubs, build container only) in
tests/test_oracle_vs_reference.py and against tests/golden/image_small.pt.
"""
import math

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2


def process_resize(w, h, resize, df=None, resize_no_larger_than=False):
    """utils.py:14-30"""
    assert len(resize) > 0 and len(resize) <= 2
    if resize_no_larger_than and (max(h, w) <= max(resize)):
        w_new, h_new = w, h
    elif len(resize) == 1 and resize[0] > -1:
        scale = resize[0] / max(h, w)
        w_new, h_new = int(round(w * scale)), int(round(h * scale))
    elif len(resize) == 1 and resize[0] == -1:
        w_new, h_new = w, h
    else:
        w_new, h_new = resize[0], resize[1]
    if df is not None:
        w_new, h_new = int(w_new // df * df), int(h_new // df * df)
    return w_new, h_new


def read_grayscale_from_array(image, resize=None, resize_no_larger_than=False, df=None):
    """image: uint8 (H, W) as cv2.imread(..., IMREAD_GRAYSCALE) returns it -> (tensor [1,h,w] float32, scales [2], original_hw [2])."""
    import PIL.Image
    resize = tuple(resize) if resize is not None else None
    w, h = image.shape[1], image.shape[0]
    w_new, h_new = process_resize(w, h, resize if resize is not None else (w, h), df, resize_no_larger_than)
    scales = torch.tensor([float(h) / float(h_new), float(w) / float(w_new)])
    original_hw = torch.tensor([h, w])
    resized = PIL.Image.fromarray(image.astype(np.uint8)).resize((w_new, h_new), resample=PIL.Image.LANCZOS)   # resize_image :169-173
    resized = np.asarray(resized, dtype=image.dtype).astype("float32")                                          # :146
    return torch.from_numpy(resized / 255.).float()[None], scales, original_hw                                  # grayscale2tensor


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _lanczos(x):
    return _sinc(x) * _sinc(x / 3) if -3.0 <= x < 3.0 else 0.0


def coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for the whole-image box -> (bounds [out,2], kk [out,ksize] int32)."""
    in0, in1 = 0.0, float(in_size)
    scale = filterscale = (in1 - in0) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = in0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(v * (1 << PRECISION_BITS) - 0.5) if v < 0 else int(v * (1 << PRECISION_BITS) + 0.5)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def resample_8bpc(img, out_w, out_h):
    """ImagingResample for mode 'L': horizontal pass, then vertical pass, uint8 in between; unchanged axes are skipped."""
    H, W = img.shape
    cur = img
    half = 1 << (PRECISION_BITS - 1)
    if out_w != W:
        b, kk = coeffs(W, out_w)
        tmp = np.zeros((H, out_w), np.uint8)
        for xx in range(out_w):
            xmin, n = b[xx]
            acc = half + cur[:, xmin:xmin + n].astype(np.int64) @ kk[xx, :n].astype(np.int64)
            tmp[:, xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
        cur = tmp
    if out_h != H:
        b, kk = coeffs(H, out_h)
        out = np.zeros((out_h, cur.shape[1]), np.uint8)
        for yy in range(out_h):
            ymin, n = b[yy]
            acc = half + kk[yy, :n].astype(np.int64) @ cur[ymin:ymin + n].astype(np.int64)
            out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
        cur = out
    return cur
