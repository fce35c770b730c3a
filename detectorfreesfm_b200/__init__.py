"""detectorfreesfm_b200 -- B200-native engine for DetectorFreeSfM's two dense-matching hot paths.

Host side = thin Python mirroring the reference's plugin interfaces; device side = hand-written sm_100a CUDA
(libdfsfm_b200.so) behind the C ABI of include/dfsfm_b200.h.  No CPU fallback.
"""
from ._lib import DfsfmError, load_library  # noqa: F401

__all__ = ["DfsfmError", "load_library", "B200LoFTR", "B200MultiviewMatcher", "KeypointMerger", "merge_keypoints", "GpuImageReader", "B200CoarseMatchingDataset", "B200MatchingMultiviewData"]


def __getattr__(name):
    if name == "B200LoFTR":
        from .coarse_matcher import B200LoFTR
        return B200LoFTR
    if name == "B200MultiviewMatcher":
        from .refine_matcher import B200MultiviewMatcher
        return B200MultiviewMatcher
    if name in ("KeypointMerger", "merge_keypoints"):
        from . import postprocess
        return getattr(postprocess, name)
    if name in ("GpuImageReader", "B200CoarseMatchingDataset"):
        from . import image_pipeline
        return getattr(image_pipeline, name)
    if name == "B200MatchingMultiviewData":
        from .chunk_dataset import B200MatchingMultiviewData
        return B200MatchingMultiviewData
    raise AttributeError(name)
