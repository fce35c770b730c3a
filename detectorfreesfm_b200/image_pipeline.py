"""Host image pipeline for HP-1: the reference's ``read_grayscale`` / ``CoarseMatchingDataset`` with the
resize on the GPU and every image decoded once.

Reference behaviour mirrored (src/dataset/utils.py:121-177, src/dataset/coarse_matching_dataset.py:10-98): decode with
``cv2.imread(path, IMREAD_GRAYSCALE)``; new size from ``process_resize`` (longest side -> ``img_resize``, both sides floored to a
multiple of ``df``); ``PIL.Image.resize((w_new, h_new), LANCZOS)`` on the uint8 image; ``/255`` -> float32 ``[1, h, w]``;
``scales = (h / h_new, w / w_new)``.  The reference does this twice per pair on the CPU unless ``img_preload`` is set.

Here the decoded uint8 image goes through a pinned staging buffer to the GPU and is resized there by
``dfsfm_resize_lanczos_gray`` (csrc/image_ops.cu) -- Pillow's 8-bit resampler is integer arithmetic once its fixed-point
coefficient tables exist, so the result is bit-identical to PIL's; the tables are built here exactly as Pillow's
``precompute_coeffs`` / ``normalize_coeffs_8bpc`` do (Pillow 12.2, src/libImaging/Resample.c; ``math.sin`` is the same libm call).
Results are cached per image path on the device, so a scene's N images are decoded and resized N times, not 2 x pairs.
There is no CPU resize path: without the CUDA library the reader raises.
"""
import collections
import ctypes
import math
import os.path as osp

import numpy as np
import torch

from . import _lib

PRECISION_BITS = 32 - 8 - 2  # Resample.c


def process_resize(w, h, resize, df=None, resize_no_larger_than=False):
    """src/dataset/utils.py:14-30."""
    assert len(resize) > 0 and len(resize) <= 2
    if resize_no_larger_than and (max(h, w) <= max(resize)):
        w_new, h_new = w, h
    else:
        if len(resize) == 1 and resize[0] > -1:  # resize the larger side
            scale = resize[0] / max(h, w)
            w_new, h_new = int(round(w * scale)), int(round(h * scale))
        elif len(resize) == 1 and resize[0] == -1:
            w_new, h_new = w, h
        else:
            w_new, h_new = resize[0], resize[1]
    if df is not None:
        w_new, h_new = map(lambda x: int(x // df * df), [w_new, h_new])
    return w_new, h_new


def _lanczos(x):
    """Resample.c: lanczos_filter / sinc_filter (a = 3)."""
    if -3.0 <= x < 3.0:
        if x == 0.0:
            return 1.0
        a = x * math.pi
        b = (x / 3.0) * math.pi
        return (math.sin(a) / a) * (math.sin(b) / b)
    return 0.0


def lanczos_coeffs(in_size, out_size):
    """Pillow's fixed-point tables for one axis -> (bounds int32 [out,2] = (first input index, taps), coef int32 [out,ksize]).

    precompute_coeffs (whole-image box: in0 = 0, in1 = in_size) followed by normalize_coeffs_8bpc, double arithmetic in the
    same order."""
    scale = filterscale = float(in_size) / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = 3.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coef = np.zeros((out_size, ksize), dtype=np.int32)
    one = float(1 << PRECISION_BITS)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        k = [_lanczos((x + xmin - center + 0.5) * ss) for x in range(n)]
        ww = 0.0
        for w in k:
            ww += w
        for x in range(n):
            v = k[x] / ww if ww != 0.0 else k[x]
            coef[xx, x] = int(-0.5 + v * one) if v < 0 else int(0.5 + v * one)
        bounds[xx, 0], bounds[xx, 1] = xmin, n
    return bounds, coef


class GpuImageReader:
    """``read_grayscale`` with the resize on the GPU.  One instance per process / device."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise _lib.DfsfmError("GpuImageReader needs a CUDA device (there is no CPU resize path)")
        self._lib = _lib.load_library()
        idx = torch.cuda.current_device() if device is None else (torch.device(device).index or 0)
        self.device = torch.device("cuda", idx)
        self._tables = {}                      # (in, out) -> (bounds_dev, coef_dev, ksize)
        self._pinned = [None, None]            # two staging buffers, alternated; an event guards reuse
        self._events = [None, None]
        self._turn = 0

    def _axis(self, in_size, out_size):
        key = (in_size, out_size)
        t = self._tables.get(key)
        if t is None:
            b, c = lanczos_coeffs(in_size, out_size)
            t = (torch.from_numpy(b).to(self.device), torch.from_numpy(c).to(self.device), int(c.shape[1]))
            self._tables[key] = t
        return t

    def _stage(self, image_u8):
        """numpy uint8 (H, W) -> device uint8 tensor through a reusable pinned buffer."""
        h, w = image_u8.shape
        i = self._turn
        self._turn ^= 1
        if self._events[i] is not None:
            self._events[i].synchronize()      # the copy that last used this buffer has finished
        n = h * w
        if self._pinned[i] is None or self._pinned[i].numel() < n:
            self._pinned[i] = torch.empty(max(n, 1 << 20), dtype=torch.uint8, pin_memory=True)
        host = self._pinned[i][:n].view(h, w)
        host.numpy()[...] = image_u8
        dev = host.to(self.device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._events[i] = ev
        return dev

    def resize_gray(self, image_u8, size):
        """uint8 (H, W) numpy array or tensor -> CUDA float32 (h_new, w_new) = PIL-LANCZOS(image) / 255; size = (w_new, h_new)."""
        if isinstance(image_u8, np.ndarray):
            assert image_u8.dtype == np.uint8 and image_u8.ndim == 2
            img = self._stage(np.ascontiguousarray(image_u8))
        else:
            assert image_u8.dtype == torch.uint8 and image_u8.dim() == 2
            img = image_u8.to(self.device).contiguous()
        h, w = int(img.shape[0]), int(img.shape[1])
        w_new, h_new = int(size[0]), int(size[1])
        out = torch.empty((h_new, w_new), dtype=torch.float32, device=self.device)
        xb = xc = yb = yc = None
        xk = yk = 0
        if w_new != w:
            xb, xc, xk = self._axis(w, w_new)
        if h_new != h:
            yb, yc, yk = self._axis(h, h_new)
        tmp = torch.empty((h, w_new), dtype=torch.uint8, device=self.device) if (xb is not None and yb is not None) else None
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.dfsfm_resize_lanczos_gray(p(img), h, w, w, p(xb), p(xc), xk, p(yb), p(yc), yk, h_new, w_new, p(tmp), p(out),
                                                       ctypes.c_void_p(st)))
        return out

    def resize_rgb(self, image_u8, size):
        """uint8 (H, W, 3) RGB numpy array or tensor -> CUDA float32 (3, h_new, w_new) = PIL-LANCZOS(image) / 255.
        Pillow resamples the bands of an RGB image independently with the same coefficient tables (Resample.c, 8bpc, bands == 3), so
        three passes of the grayscale kernel over the de-interleaved planes are bit-identical to PIL's result."""
        if isinstance(image_u8, np.ndarray):
            assert image_u8.dtype == np.uint8 and image_u8.ndim == 3 and image_u8.shape[2] == 3
            h, w = image_u8.shape[:2]
            img = self._stage(np.ascontiguousarray(image_u8).reshape(h, w * 3)).view(h, w, 3)
        else:
            assert image_u8.dtype == torch.uint8 and image_u8.dim() == 3 and image_u8.shape[2] == 3
            img = image_u8.to(self.device)
        planes = img.permute(2, 0, 1).contiguous()
        return torch.stack([self.resize_gray(planes[c], size) for c in range(3)], 0)

    def read_rgb(self, path, resize=None, resize_no_larger_than=False, df=None, pad_to=None, ret_scales=False):
        """src/dataset/utils.py:80-118 (client=None, augmentor=None; the loader of the refinement stage,
        src/dataset/coarse_sfm_refinement_dataset.py:365-380): image [3,h,w] float32 in [0,1] on the GPU (+ scales, original_hw)."""
        import cv2
        resize = tuple(resize) if resize is not None else None
        image = cv2.imread(str(path), cv2.IMREAD_COLOR)
        if image is None:
            raise FileNotFoundError(f"Problem exists when loading image: {path}")
        image = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
        w, h = image.shape[1], image.shape[0]
        w_new, h_new = process_resize(w, h, resize if resize is not None else (w, h), df, resize_no_larger_than=resize_no_larger_than)
        scales = torch.tensor([float(h) / float(h_new), float(w) / float(w_new)])
        original_hw = torch.tensor([h, w])
        img = self.resize_rgb(image, (w_new, h_new))
        if pad_to is not None:
            if pad_to == -1:
                pad_to = max(w_new, h_new)
            assert pad_to >= max(h_new, w_new)
            padded = torch.zeros((3, pad_to, pad_to), dtype=torch.float32, device=self.device)
            padded[:, :h_new, :w_new] = img
            img = padded
        return [img, scales, original_hw] if ret_scales else img

    def read_grayscale(self, path, resize=None, resize_no_larger_than=False, df=None, pad_to=None, ret_scales=False, ret_pad_mask=False):
        """src/dataset/utils.py:121-159 (client=None, augmentor=None): image [1,h,w] float32 on the GPU (+ scales, original_hw)."""
        import cv2
        resize = tuple(resize) if resize is not None else None
        image = cv2.imread(str(path), cv2.IMREAD_GRAYSCALE)
        if image is None:
            raise FileNotFoundError(f"Problem exists when loading image: {path}")
        w, h = image.shape[1], image.shape[0]
        w_new, h_new = process_resize(w, h, resize if resize is not None else (w, h), df, resize_no_larger_than=resize_no_larger_than)
        scales = torch.tensor([float(h) / float(h_new), float(w) / float(w_new)])
        original_hw = torch.tensor([h, w])
        img = self.resize_gray(image, (w_new, h_new))
        mask = None
        if pad_to is not None:  # pad_bottom_right, utils.py:33-52 (zeros stay zeros through the /255)
            if pad_to == -1:
                pad_to = max(w_new, h_new)
            assert pad_to >= max(h_new, w_new)
            padded = torch.zeros((pad_to, pad_to), dtype=torch.float32, device=self.device)
            padded[:h_new, :w_new] = img
            if ret_pad_mask:
                mask = torch.zeros((pad_to, pad_to), dtype=torch.float32, device=self.device)
                mask[:h_new, :w_new] = 1
            img = padded
        ret = [img[None]]
        if ret_scales:
            ret += [scales, original_hw]
        if ret_pad_mask:
            ret.append(mask if pad_to else None)
        return ret[0] if len(ret) == 1 else ret


class B200CoarseMatchingDataset(torch.utils.data.Dataset):
    """``CoarseMatchingDataset`` (src/dataset/coarse_matching_dataset.py:10-98) with device-resident, per-image cached inputs.

    Same constructor arguments and the same item dict (``image0/1`` [1,h,w] float32, ``scale0/1`` [2], ``f_name0/1``, ``frameID``,
    ``pair_key``); images live on the GPU, so use it with ``num_workers=0`` (the default collate adds the batch dimension).
    ``cache_images`` bounds the number of cached images (LRU); ``img_preload`` fills the cache up front like the reference."""

    def __init__(self, args, image_lists, covis_pairs, subset_ids, device=None, cache_images=256):
        super().__init__()
        if args["img_type"] != "grayscale":
            raise NotImplementedError("the matchers on this path take grayscale input (read_rgb is not built)")
        self.img_dir = image_lists
        self.img_resize = args["img_resize"]
        self.df = args["df"]
        self.pad_to = args["pad_to"]
        self.preload = args["img_preload"]
        self.subset_ids = subset_ids
        if isinstance(covis_pairs, list):
            self.pair_list = covis_pairs
        else:
            assert osp.exists(covis_pairs)
            with open(covis_pairs, "r") as f:
                self.pair_list = f.read().rstrip("\n").split("\n")
        self.reader = GpuImageReader(device)
        self.cache_images = max(2, int(cache_images))
        self.img_dict = collections.OrderedDict()
        self.decodes = 0
        if self.preload:
            self.cache_images = max(self.cache_images, len(self.img_dir))
            for image_path in self.img_dir:
                self._image(image_path)

    def _image(self, path):
        hit = self.img_dict.get(path)
        if hit is not None:
            self.img_dict.move_to_end(path)
            return hit
        item = self.reader.read_grayscale(path, (self.img_resize,) if self.img_resize is not None else None, df=self.df, pad_to=self.pad_to,
                                          ret_scales=True)
        self.decodes += 1
        self.img_dict[path] = item
        while len(self.img_dict) > self.cache_images:
            self.img_dict.popitem(last=False)
        return item

    def __len__(self):
        return len(self.subset_ids)

    def __getitem__(self, idx):
        pair_idx = self.subset_ids[idx]
        img_path0, img_path1 = self.pair_list[pair_idx].split(" ")
        img0, scale0, _ = self._image(img_path0)
        img1, scale1, _ = self._image(img_path1)
        return {
            "image0": img0,
            "image1": img1,
            "scale0": scale0,
            "scale1": scale1,
            "f_name0": osp.basename(img_path0).rsplit(".", 1)[0],
            "f_name1": osp.basename(img_path1).rsplit(".", 1)[0],
            "frameID": pair_idx,
            "pair_key": (img_path0, img_path1),
        }
