"""Build the oracle's native pieces -- TEST INFRASTRUCTURE.

* oracle/_build/libroialign_oracle.so  <- oracle/roialign_oracle.c (gcc), the C restatement.
* oracle/_ref/crop_and_resize_cpu.so   <- the reference's OWN crop_and_resize.cpp compiled from the sources where
  they lie under /root/reference (only when that tree is present; nothing is copied into the repo).
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
SO = os.path.join(BUILD, "libroialign_oracle.so")


def build_c(verbose=False):
    src = os.path.join(HERE, "roialign_oracle.c")
    os.makedirs(BUILD, exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        cmd = ["gcc", "-O2", "-fPIC", "-shared", "-o", SO, src, "-lm"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return SO


def build_ref(verbose=False):
    from . import ref_shims
    if not ref_shims.available():
        return None
    try:
        return ref_shims.build_ref_roialign()
    except Exception as e:  # the reference build is optional evidence, never a product dependency
        if verbose:
            print("reference RoIAlign build failed:", e)
        return None


def build_all(verbose=False):
    build_c(verbose)
    build_ref(verbose)


_lib = None


def roialign_forward(image, boxes, box_index, crop_h, crop_w, extrapolation_value=0.0):
    """image [N,C,H,W] fp32, boxes [n,4] (y1,x1,y2,x2 normalised), box_index [n] int32 -> crops [n,C,ch,cw]."""
    import torch
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
        _lib.roialign_oracle_forward.restype = ctypes.c_int
    image = image.contiguous().float()
    boxes = boxes.contiguous().float()
    box_index = box_index.contiguous().to(torch.int32)
    n = boxes.shape[0]
    N, C, H, W = image.shape
    crops = torch.zeros(n, C, crop_h, crop_w)
    rc = _lib.roialign_oracle_forward(ctypes.c_void_p(image.data_ptr()), N, C, H, W, ctypes.c_void_p(boxes.data_ptr()),
                                      ctypes.c_void_p(box_index.data_ptr()), n, ctypes.c_float(extrapolation_value), crop_h, crop_w,
                                      ctypes.c_void_p(crops.data_ptr()))
    if rc != 0:
        raise ValueError("box index out of range")
    return crops
