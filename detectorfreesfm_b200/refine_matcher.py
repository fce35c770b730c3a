"""HP-2 host side: a drop-in for the reference's ``MultiviewMatcher`` module behind the refinement-chunk API.

Mirrors src/MultiviewMatcher/MultiviewMatcher.py:16-59: ``B200MultiviewMatcher(config, test=True)`` is called as
``matcher(data)`` under ``torch.no_grad()`` from src/post_optimization/matcher_model/multiview_match_worker.py:59-64 with
the chunk dict of construct_matching_data.py:317-476, and writes ``query_points_refined`` [1,M,2],
``reference_points_refined`` (list, last = [1,N-1,M,2]) and ``std`` into it.  All compute runs in libdfsfm_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .packing import pack_multiview


class B200MultiviewMatcher(torch.nn.Module):
    def __init__(self, config, profiler=None, debug=False, test=False, plotting_vis=None, test_speed=False):
        super().__init__()
        self.config = config
        if not test:
            raise NotImplementedError("only the inference configuration (test=True) is built")
        if config["n_matching_steps"] != 1 or config["enable_multiview_scale_align"]:
            raise NotImplementedError("engine implements the shipped config: n_matching_steps=1, no multiview scale align")
        bb = config["backbone"]
        if bb["type"] != "S2DNet" or list(bb["resolution"]) != [4, 1]:
            raise NotImplementedError("engine implements the S2DNet backbone at resolution [4, 1]")
        mt, mm = config["multiview_transform"], config["multiview_matching_test"]
        if mt["crop_size"] != 35 or mt["d_model"] != 128 or mt["nhead"] != 8 or list(mt["layer_names"]) * mt["layer_iter_n"] != ["self", "cross"] * 2:
            raise NotImplementedError("engine implements crop 35, d_model 128, 8 heads, ['self','cross']*2")
        if mm["best_left_strategy"] != "smallest_mean_std" or mm["left_point_movement_window_size"] is None:
            raise NotImplementedError("engine implements the test-time movable reference point (smallest_mean_std)")
        self.W = int(mt["window_size"])
        self.LW = int(mm["left_point_movement_window_size"])
        assert int(bb["s2dnet"]["window_size"]) == self.W and int(mm["window_size"]) == self.W
        self._lib = _lib.load_library()
        self._h = ctypes.c_void_p()
        self._device = None
        self._packed = None

    def cuda(self, device=None):
        if device is None:
            idx = torch.cuda.current_device()
        elif isinstance(device, int):
            idx = device
        else:
            idx = torch.device(device).index
            idx = torch.cuda.current_device() if idx is None else idx
        dev = torch.device("cuda", idx)
        if self._device != dev:
            self._destroy()
            self._device = dev
            _lib.check(self._lib.dfsfm_refine_create(ctypes.byref(self._h), idx, self.W, self.LW))
            if self._packed is not None:
                self._upload()
        return self

    def load_state_dict(self, state_dict, strict=True):
        self._packed = pack_multiview(state_dict)
        if self._h:
            self._upload()
        return self

    def _upload(self):
        for name, (t, kind) in self._packed.items():
            _lib.check(self._lib.dfsfm_refine_set_param(self._h, name.encode(), ctypes.c_void_p(t.data_ptr()), t.shape[0], t.shape[1], kind))

    def _destroy(self):
        if self._h:
            self._lib.dfsfm_refine_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    @torch.no_grad()
    def forward(self, data, chunk_track=1000, chunk_backbone_img=True):
        """Same contract as MultiviewMatcher.forward (MultiviewMatcher.py:59): updates ``data`` in place.  chunk_track /
        chunk_backbone_img only bound memory in the reference; results do not depend on them."""
        if not self._h:
            raise _lib.DfsfmError("B200MultiviewMatcher: call .cuda() before forward (there is no CPU path)")
        if self._packed is None:
            raise _lib.DfsfmError("B200MultiviewMatcher: load_state_dict() has not been called")
        images = data["images"]
        if isinstance(images, torch.Tensor):
            assert images.shape[0] == 1
            images = [images[:, i] for i in range(images.shape[1])]
        images = [im[0].contiguous().float() for im in images]  # [3,H,W] on the device
        n_img = len(images)
        for im in images:
            assert im.is_cuda and im.shape[0] == 3
        dev = images[0].device
        ptrs = (ctypes.c_void_p * n_img)(*[im.data_ptr() for im in images])
        Hs = np.asarray([im.shape[1] for im in images], dtype=np.int32)
        Ws = np.asarray([im.shape[2] for im in images], dtype=np.int32)

        # The C ABI takes the small per-track arrays as HOST arrays (it builds the patch records on the host).  The reference worker hands
        # the whole dict over as CUDA tensors (dict_to_cuda, multiview_match_worker.py:127): pack them into ONE float32 buffer on the
        # device and read it back with a single copy (image indices <= 16 and the masks are exact in float32); CPU tensors are used as
        # they are.  The three outputs stay on the device.
        q_idx_t = data["query_img_idxs"][0]
        M = int(q_idx_t.shape[0])
        rpts_t = data["reference_points_coarse"][0]
        Nq = int(rpts_t.shape[0])
        parts = [data["scales"][0] if "scales" in data else torch.ones(n_img, 2), data["query_points"][0], rpts_t, data["track_valid_mask"][0],
                 q_idx_t, data["reference_img_idxs"][0],
                 data["query_movable_mask"][0] if "query_movable_mask" in data else torch.ones(M, dtype=torch.bool)]
        sizes = [p.numel() for p in parts]
        if any(p.is_cuda for p in parts):
            flat = torch.cat([p.reshape(-1).to(device=dev, dtype=torch.float32) for p in parts]).cpu().numpy()
        else:
            flat = torch.cat([p.reshape(-1).to(torch.float32) for p in parts]).numpy()
        offs = np.cumsum([0] + sizes)
        seg = [flat[offs[i]:offs[i + 1]] for i in range(len(parts))]
        scales = np.ascontiguousarray(seg[0], dtype=np.float32)
        qpts = np.ascontiguousarray(seg[1], dtype=np.float32)
        rpts = np.ascontiguousarray(seg[2], dtype=np.float32)
        valid = np.ascontiguousarray(seg[3] != 0, dtype=np.uint8)
        q_idx = np.ascontiguousarray(np.rint(seg[4]), dtype=np.int32)
        r_idx = np.ascontiguousarray(np.rint(seg[5]), dtype=np.int32)
        movable = np.ascontiguousarray(seg[6] != 0, dtype=np.uint8)
        out_q = torch.zeros((1, M, 2), dtype=torch.float32, device=dev)
        out_r = torch.zeros((1, Nq, M, 2), dtype=torch.float32, device=dev)
        out_s = torch.zeros((1, Nq, M), dtype=torch.float32, device=dev)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        with torch.cuda.device(dev):
            _lib.check(self._lib.dfsfm_refine_chunk(self._h, n_img, ctypes.cast(ptrs, ctypes.c_void_p), p(Hs), p(Ws), p(scales), M, Nq, p(qpts),
                                                    p(rpts), p(valid), p(q_idx), p(r_idx), p(movable), _lib.ptr(out_q), _lib.ptr(out_r),
                                                    _lib.ptr(out_s), _lib.stream_ptr(dev)))
        data["W"] = self.W
        data["query_points_refined"] = out_q
        rr, ss = out_r, out_s
        if "reference_points_refined" in data:
            data["reference_points_refined"].append(rr)
            data["std"].append(ss)
        else:
            data["reference_points_refined"] = [rr]
            data["std"] = [ss]
        return None
