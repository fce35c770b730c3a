"""Match -> keypoint -> index post-processing on the GPU: the block of src/coarse_match/coarse_match.py:203-237.

``merge_keypoints(matches, image_lists, pair_name_split)`` takes what ``match_worker`` returns (an ordered dict
``"name0<split>name1" -> (M,5)`` of ``[x0, y0, x1, y1, conf]``, coarse_match_worker.py:139-141) and returns what the reference
builds with ``Match2Kpts`` + ``keypoint_worker`` + ``update_matches(merge=False)`` + ``transform_keypoints``:

    final_keypoints[name]   float32 (n, 2)  truncated coordinates in keypoint-id order (np.empty((0, 2)) if the image has none)
    final_scores[name]      float32 (n,)    summed match confidence
    updated_matches[pair]   int64   (M, 2)  keypoint ids of the two end points of every match

The dict bookkeeping (names, pair keys) stays in Python; unique / sum / ranking / id look-up run in libdfsfm_b200.so
(csrc/postprocess.cu) on all pairs at once.  Match arrays may be CUDA tensors (the matcher's own outputs: no host round trip),
CPU tensors or numpy arrays.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _split_pair(key, pair_name_split):
    try:  # src/coarse_match/utils/merge_kpts.py:27-30
        name0, name1 = key.split(pair_name_split)
    except ValueError:
        name0, name1 = key.split("-")
    return name0, name1


class KeypointMerger:
    """Owns the device workspace; ``merge`` is the device step on flat arrays, ``__call__`` the dict-level mirror."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise _lib.DfsfmError("KeypointMerger needs a CUDA device (there is no CPU fallback)")
        self._lib = _lib.load_library()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index or 0)
        self._h = ctypes.c_void_p()
        _lib.check(self._lib.dfsfm_post_create(ctypes.byref(self._h), self.device.index))

    def __del__(self):
        try:
            if self._h:
                self._lib.dfsfm_post_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    def merge(self, rows, pair_offset, pair_images, n_images):
        """rows: cuda float32 [T,5]; pair_offset: int64 [P+1]; pair_images: int32 [P,2] ->
        (kpt_xy [K,2] f32, kpt_score [K] f32, image_offset [n_images+1] i32, match_ids [T,2] i32), all CUDA tensors."""
        T = int(rows.shape[0])
        rows = rows.to(self.device, torch.float32).contiguous()
        pair_offset = torch.as_tensor(pair_offset, dtype=torch.int64).to(self.device).contiguous()
        pair_images = torch.as_tensor(pair_images, dtype=torch.int32).to(self.device).contiguous()
        P = int(pair_images.shape[0])
        assert pair_offset.numel() == P + 1
        cap = max(2 * T, 1)
        kpt_xy = torch.empty((cap, 2), dtype=torch.float32, device=self.device)
        kpt_score = torch.empty((cap,), dtype=torch.float32, device=self.device)
        img_off = torch.empty((n_images + 1,), dtype=torch.int32, device=self.device)
        ids = torch.empty((max(T, 1), 2), dtype=torch.int32, device=self.device)
        k = ctypes.c_int64(0)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.dfsfm_post_merge_keypoints(
            self._h, rows.data_ptr(), T, P, pair_offset.data_ptr(), pair_images.data_ptr(), int(n_images), kpt_xy.data_ptr(),
            kpt_score.data_ptr(), img_off.data_ptr(), ids.data_ptr(), ctypes.byref(k), ctypes.c_void_p(st)))
        K = int(k.value)
        return kpt_xy[:K], kpt_score[:K], img_off, ids[:T]

    def __call__(self, matches, image_lists, pair_name_split=" "):
        names = list(image_lists)
        index = {n: i for i, n in enumerate(names)}
        keys = list(matches.keys())
        pair_images = np.zeros((len(keys), 2), dtype=np.int32)
        counts = np.zeros((len(keys),), dtype=np.int64)
        parts = []
        for p, k in enumerate(keys):
            n0, n1 = _split_pair(k, pair_name_split)
            pair_images[p] = (index[n0], index[n1])
            v = matches[k]
            v = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
            counts[p] = v.shape[0]
            if v.shape[0]:
                parts.append(v.to(self.device, torch.float32))
        pair_offset = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        rows = torch.cat(parts, 0) if parts else torch.empty((0, 5), dtype=torch.float32, device=self.device)
        kpt_xy, kpt_score, img_off, ids = self.merge(rows, pair_offset, pair_images, len(names))
        kpt_xy, kpt_score, img_off, ids = kpt_xy.cpu().numpy(), kpt_score.cpu().numpy(), img_off.cpu().numpy(), ids.cpu().numpy()
        final_kpts, final_scores, updated = {}, {}, {}
        for i, n in enumerate(names):
            a, b = int(img_off[i]), int(img_off[i + 1])
            final_kpts[n] = kpt_xy[a:b].copy() if b > a else np.empty((0, 2))  # transform_keypoints' n_kpts=0 corner case
            final_scores[n] = kpt_score[a:b].copy()
        for p, k in enumerate(keys):
            updated[k] = ids[pair_offset[p]:pair_offset[p + 1]].astype(int)      # update_matches: mids.astype(int), (M, 2)
        return final_kpts, final_scores, updated


_default = None


def merge_keypoints(matches, image_lists, pair_name_split=" "):
    """Module-level convenience with a shared workspace (see KeypointMerger)."""
    global _default
    if _default is None:
        _default = KeypointMerger()
    return _default(matches, image_lists, pair_name_split)
