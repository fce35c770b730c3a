"""Import the UNMODIFIED reference modules from /root/reference on CPU.

TEST INFRASTRUCTURE, build-container only (the GPU box has no /root/reference): used by
tests/test_oracle_vs_reference.py and tests/golden/make_golden.py to pin the oracle
restatement against the reference itself.  Nothing is copied; the reference's files are
imported where they lie, behind stubs for its un-installed third-party deps
(kornia 0.4.1: two functions; yacs CfgNode; omegaconf; timm registry; loguru) --
SURVEY.md section 8(c).
"""
import importlib
import importlib.util
import os
import sys
import types

REF = os.environ.get("DFSFM_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "third_party", "LoFTR"))


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _install_stubs():
    from . import loftr_oracle as lo
    import torch

    def spatial_expectation2d(inp, normalized_coordinates=True):
        return lo.spatial_expectation2d(inp)

    def create_meshgrid(h, w, normalized_coordinates=True, device=None):
        return lo.create_meshgrid(h, w).to(device) if device is not None else lo.create_meshgrid(h, w)

    try:
        import kornia  # noqa: F401
    except Exception:
        _mod("kornia")
        _mod("kornia.geometry")
        _mod("kornia.geometry.subpix")
        dsnt = _mod("kornia.geometry.subpix.dsnt", spatial_expectation2d=spatial_expectation2d)
        sys.modules["kornia.geometry.subpix"].dsnt = dsnt
        _mod("kornia.utils")
        _mod("kornia.utils.grid", create_meshgrid=create_meshgrid)
    try:
        import yacs  # noqa: F401
    except Exception:
        class CfgNode(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

            def __setattr__(self, k, v):
                self[k] = v

            def clone(self):
                import copy
                return copy.deepcopy(self)

            def _merge(self, other):
                for k, v in other.items():
                    if isinstance(v, CfgNode) and isinstance(self.get(k), CfgNode):
                        self[k]._merge(v)
                    else:
                        self[k] = v

            def merge_from_file(self, path):
                """yacs loads a ``.py`` config by importing it and reading its ``cfg`` (yacs/config.py load_cfg_py_source)"""
                import runpy
                assert path.endswith(".py"), path
                self._merge(runpy.run_path(path)["cfg"])
        _mod("yacs")
        _mod("yacs.config", CfgNode=CfgNode)
    try:
        import omegaconf  # noqa: F401
    except Exception:
        class _Conf(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError:
                    raise AttributeError(k)

        class OmegaConf:
            @staticmethod
            def merge(*cfgs):
                out = _Conf()
                for c in cfgs:
                    out.update(dict(c))
                return out

            @staticmethod
            def set_struct(c, v):
                pass

            @staticmethod
            def set_readonly(c, v):
                pass
        _mod("omegaconf", OmegaConf=OmegaConf)
    try:
        import timm  # noqa: F401
    except Exception:
        _mod("timm")
        _mod("timm.models")
        _mod("timm.models.registry", register_model=lambda f: f)
        _mod("timm.models.layers", DropPath=torch.nn.Identity, trunc_normal_=lambda *a, **k: None,
             to_2tuple=lambda x: (x, x))
    try:
        import loguru  # noqa: F401
    except Exception:
        import logging
        _mod("loguru", logger=logging.getLogger("ref"))


def import_loftr():
    """-> (LoFTR class, default_cfg dict) from third_party/LoFTR/src/loftr."""
    _install_stubs()
    root = os.path.join(REF, "third_party", "LoFTR")
    if root not in sys.path:
        sys.path.insert(0, root)
    # 'src' here is third_party/LoFTR/src; keep it out of the way of the main repo's src
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    mod = importlib.import_module("src.loftr")
    LoFTR, default_cfg = mod.LoFTR, mod.default_cfg
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        sys.modules["_loftr_" + k] = sys.modules.pop(k)
    sys.path.remove(root)
    return LoFTR, default_cfg


def loftr_config(thr=0.2, fine=False, temperature=0.1):
    """The dict coarse_match_worker.build_model produces (coarse_match_worker.py:31-36)."""
    import copy
    _, default_cfg = import_loftr()
    cfg = copy.deepcopy(default_cfg)
    cfg["coarse"]["temp_bug_fix"] = False
    cfg["match_coarse"]["thr"] = thr
    cfg["match_coarse"]["dsmax_temperature"] = temperature
    cfg["match_coarse"]["skh_prefilter"] = False
    cfg["match_coarse"]["sparse_spvs"] = True
    cfg["fine"]["enable"] = fine
    return cfg


_ROI_EXT = None


def build_ref_roialign():
    """Compile the reference's own CPU RoIAlign (crop_and_resize.cpp, OpenMP) from the
    sources where they lie into oracle/_ref/ (git-ignored; travels to the GPU box)."""
    global _ROI_EXT
    if _ROI_EXT is not None:
        return _ROI_EXT
    from torch.utils.cpp_extension import load
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    os.makedirs(out, exist_ok=True)
    src = os.path.join(REF, "third_party", "RoIAlign.pytorch", "roi_align", "src", "crop_and_resize.cpp")
    # the image has no libgomp.spec, so -fopenmp is unavailable: the OpenMP pragma (crop_and_resize.cpp:31) is
    # ignored and the reference op runs single-threaded -- same arithmetic.
    _ROI_EXT = load(name="crop_and_resize_cpu", sources=[src], build_directory=out,
                    extra_cflags=["-O2", "-w"], verbose=False)
    return _ROI_EXT


def load_prebuilt_ref_roialign():
    """Load oracle/_ref/crop_and_resize_cpu.so if it was built (works without /root/reference)."""
    global _ROI_EXT
    if _ROI_EXT is not None:
        return _ROI_EXT
    import torch  # noqa: F401  (the .so links against libtorch)
    so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "crop_and_resize_cpu.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("crop_and_resize_cpu", so)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    _ROI_EXT = m
    return m


def import_multiview():
    """-> MultiviewMatcher class from src/MultiviewMatcher (never executes src/__init__.py)."""
    _install_stubs()
    ext = build_ref_roialign()
    roi_dir = os.path.join(REF, "third_party", "RoIAlign.pytorch", "roi_align")
    pkg = _mod("roi_align")
    pkg.__path__ = [roi_dir]
    sys.modules["roi_align.crop_and_resize_cpu"] = ext
    pkg.crop_and_resize_cpu = ext
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    src = _mod("src")
    src.__path__ = [os.path.join(REF, "src")]
    utils = _mod("src.utils")
    utils.__path__ = []  # block the real src/utils (needs ray, h5py, ...)

    class PassThroughProfiler:
        def record_function(self, name):
            import contextlib
            return contextlib.nullcontext()
    _mod("src.utils.profiler", PassThroughProfiler=PassThroughProfiler)
    mod = importlib.import_module("src.MultiviewMatcher.MultiviewMatcher")
    return mod.MultiviewMatcher


def multiview_config(window=15, left_window=7):
    """model.multiview_refinement of hydra_training_configs/experiment/
    multiview_refinement_matching.yaml:22-90, with the per-iteration window rescale of
    multiview_match_worker.py:20-34 already applied and pretrained=None (no network)."""
    mm = {"enable": True, "type": "s2d", "detector": "OnGrid", "window_size": window,
          "best_left_strategy": "smallest_mean_std",
          "s2d": {"type": "heatmap", "obtain_offset_method": "argsoftmax"}}
    return {
        "n_matching_steps": 1, "enable_multiview_scale_align": False,
        "backbone": {"type": "S2DNet", "resolution": [4, 1],
                     "s2dnet": {"name": "s2dnet", "num_layers": 2, "window_size": window,
                                "checkpointing": None, "output_dim": 128, "pretrained": None,
                                "substitute_pooling_layers": True, "combine": True,
                                "zoomin_strategy": "post"},
                     "pretrained": None, "pretrained_fix": False},
        "use_fine_backbone_as_coarse": False, "interpol_type": "bilinear",
        "multiview_transform": {"sparse": True, "crop_size": 35, "window_size": window,
                                "enable_rescaled_crop": False, "enable": True, "type": "LoFTR",
                                "d_model": 128, "nhead": 8, "layer_names": ["self", "cross"],
                                "layer_iter_n": 2, "dropout": 0.0, "attention": "linear",
                                "norm_method": "layernorm", "attention_type": "multiview",
                                "kernel_fn": "elu + 1", "d_kernel": 16, "redraw_interval": 2,
                                "rezero": None, "final_proj": False},
        "multiview_matching_train": {**mm, "left_point_movement_window_size": None},
        "multiview_matching_test": {**mm, "left_point_movement_window_size": left_window},
    }


class _Remote:
    """stand-in for ray.remote: ``@ray.remote`` and ``@ray.remote(...)`` both leave the function as it is"""
    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return lambda f: f


def _stub_ray():
    """ray is not installed in this image: an inert module with ``remote`` and ``actor.ActorHandle`` (decorators / annotations only)"""
    m = sys.modules.get("ray")
    if m is not None and getattr(m, "__file__", None):
        return                                   # a real ray
    try:
        if m is None:
            import ray  # noqa: F401
            return
    except Exception:
        pass
    ray_mod = _mod("ray", remote=_Remote())
    ray_mod.__path__ = []
    ray_mod.actor = _mod("ray.actor", ActorHandle=object)


def import_postprocess():
    """-> (Match2Kpts, keypoint_worker, update_matches, transform_keypoints) from src/coarse_match.

    src/__init__.py, the dataset and the model builders are never executed: `src` is a bare namespace and the worker
    module's unrelated imports (ray, pytorch_lightning, the dataset class, src.utils helpers) are inert stubs.
    """
    _install_stubs()
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    src = _mod("src")
    src.__path__ = [os.path.join(REF, "src")]
    utils = _mod("src.utils")
    utils.__path__ = []
    _mod("src.utils.misc", lower_config=lambda c: c)
    _mod("src.utils.torch_utils", update_state_dict=lambda *a, **k: None, STATE_DICT_MAPPER={})
    ds = _mod("src.dataset")
    ds.__path__ = []
    _mod("src.dataset.coarse_matching_dataset", CoarseMatchingDataset=object)

    _stub_ray()
    try:
        import pytorch_lightning  # noqa: F401
    except Exception:
        _mod("pytorch_lightning", seed_everything=lambda s: None)
    try:
        import tqdm  # noqa: F401
    except Exception:
        _mod("tqdm", tqdm=lambda x, **k: x)
    worker = importlib.import_module("src.coarse_match.coarse_match_worker")
    merge = importlib.import_module("src.coarse_match.utils.merge_kpts")
    return merge.Match2Kpts, worker.keypoint_worker, worker.update_matches, worker.transform_keypoints


def import_image_utils():
    """-> the reference's src/dataset/utils.py module (read_grayscale, process_resize, resize_image).

    Its unrelated imports that this image lacks (h5py, albumentations) are inert stubs; cv2 and PIL are real."""
    _install_stubs()
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    src = _mod("src")
    src.__path__ = [os.path.join(REF, "src")]
    ds = _mod("src.dataset")
    ds.__path__ = [os.path.join(REF, "src", "dataset")]
    for name in ("h5py", "albumentations"):
        try:
            importlib.import_module(name)
        except Exception:
            _mod(name)
    return importlib.import_module("src.dataset.utils")


def import_refine_worker(dataset_cls=None):
    """-> the reference's src/post_optimization/matcher_model/multiview_match_worker.py module (matchWorker, extract_results,
    UpdatedQueryPts) for CPU comparisons of the host loop.

    ``MatchingMultiviewData`` is replaced by ``dataset_cls`` (the chunk construction is row b2 / 8(f)-2, not under test here),
    ``dict_to_cuda`` by the identity (no GPU in the build container); ray / pytorch_lightning / omegaconf are inert stubs."""
    _install_stubs()
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    src = _mod("src")
    src.__path__ = [os.path.join(REF, "src")]
    utils = _mod("src.utils")
    utils.__path__ = []
    _mod("src.utils.data_io", dict_to_cuda=lambda d: d)
    mm = _mod("src.MultiviewMatcher")
    mm.__path__ = []
    _mod("src.MultiviewMatcher.MultiviewMatcher", MultiviewMatcher=object)
    po_ = _mod("src.post_optimization")
    po_.__path__ = [os.path.join(REF, "src", "post_optimization")]
    _mod("src.post_optimization.data_construct", MatchingMultiviewData=dataset_cls if dataset_cls is not None else object)
    mmod = _mod("src.post_optimization.matcher_model")
    mmod.__path__ = [os.path.join(REF, "src", "post_optimization", "matcher_model")]

    _stub_ray()
    try:
        import pytorch_lightning  # noqa: F401
    except Exception:
        _mod("pytorch_lightning", seed_everything=lambda s: None)
    try:
        import omegaconf  # noqa: F401
    except Exception:
        _mod("omegaconf", OmegaConf=object)
    return importlib.import_module("src.post_optimization.matcher_model.multiview_match_worker")


def import_hook_modules():
    """-> (coarse_match, coarse_match_worker, matcher_model.multiview_match, multiview_match_worker) of the reference, imported
    where they lie behind inert stubs (ray, h5py, pytorch_lightning, the dataset classes, MultiviewMatcher) so that the two
    ``build_model`` hooks can be exercised on the CPU box.  yacs / omegaconf are replaced by small
    work-alikes (``merge_from_file`` of a .py config; ``OmegaConf.load`` = yaml.safe_load) when not installed."""
    import yaml
    _install_stubs()
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.") or k == "third_party" or k.startswith("third_party.")]:
        del sys.modules[k]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    src = _mod("src")
    src.__path__ = [os.path.join(REF, "src")]
    utils = _mod("src.utils")
    utils.__path__ = [os.path.join(REF, "src", "utils")]
    _stub_ray()
    for name in ("h5py",):
        try:
            importlib.import_module(name)
        except Exception:
            _mod(name)
    try:
        import pytorch_lightning  # noqa: F401
    except Exception:
        _mod("pytorch_lightning", seed_everything=lambda s: None)
    om = sys.modules.get("omegaconf")
    if om is None or not getattr(om, "__file__", None):
        class OmegaConf:
            @staticmethod
            def load(path):
                with open(path) as f:
                    return yaml.safe_load(f)

            @staticmethod
            def to_container(c, **k):
                return c
            merge = staticmethod(lambda *cfgs: {k: v for c in cfgs for k, v in dict(c).items()})
            set_struct = staticmethod(lambda c, v: None)
            set_readonly = staticmethod(lambda c, v: None)
        _mod("omegaconf", OmegaConf=OmegaConf)
    ds = _mod("src.dataset")
    ds.__path__ = []
    _mod("src.dataset.coarse_matching_dataset", CoarseMatchingDataset=object)
    mvm = _mod("src.MultiviewMatcher")
    mvm.__path__ = []
    _mod("src.MultiviewMatcher.MultiviewMatcher", MultiviewMatcher=object)
    po_ = _mod("src.post_optimization")
    po_.__path__ = [os.path.join(REF, "src", "post_optimization")]
    _mod("src.post_optimization.data_construct", MatchingMultiviewData=object)
    cm = importlib.import_module("src.coarse_match.coarse_match")
    cmw = importlib.import_module("src.coarse_match.coarse_match_worker")
    mm = importlib.import_module("src.post_optimization.matcher_model.multiview_match")
    mmw = importlib.import_module("src.post_optimization.matcher_model.multiview_match_worker")
    return cm, cmw, mm, mmw


def import_chunk_dataset():
    """-> the reference's ``MatchingMultiviewData`` class (src/post_optimization/data_construct/construct_matching_data.py; SURVEY 8(a) row
    b2 / 8(f) row 2), imported where it lies: its own geometry helpers and ``chunks_balance`` are the real files, ray / loguru are inert."""
    _install_stubs()
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    src = _mod("src")
    src.__path__ = [os.path.join(REF, "src")]
    utils = _mod("src.utils")
    utils.__path__ = [os.path.join(REF, "src", "utils")]
    colmap = _mod("src.utils.colmap")
    colmap.__path__ = [os.path.join(REF, "src", "utils", "colmap")]
    _stub_ray()
    po_ = _mod("src.post_optimization")
    po_.__path__ = [os.path.join(REF, "src", "post_optimization")]
    pu = _mod("src.post_optimization.utils")
    pu.__path__ = [os.path.join(REF, "src", "post_optimization", "utils")]
    dc = _mod("src.post_optimization.data_construct")
    dc.__path__ = [os.path.join(REF, "src", "post_optimization", "data_construct")]
    mod = importlib.import_module("src.post_optimization.data_construct.construct_matching_data")
    return mod.MatchingMultiviewData


def import_colmap_dataset_class():
    """-> the reference's ``CoarseColmapDataset`` class (src/dataset/coarse_sfm_refinement_dataset.py) for calling single methods unbound
    on a stand-in ``self`` (update_refined_kpts_to_colmap_multiview); h5py / albumentations / ray are inert."""
    _install_stubs()
    for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        del sys.modules[k]
    src = _mod("src")
    src.__path__ = [os.path.join(REF, "src")]
    for name in ("h5py", "albumentations"):
        try:
            importlib.import_module(name)
        except Exception:
            _mod(name)
    _stub_ray()
    return importlib.import_module("src.dataset.coarse_sfm_refinement_dataset").CoarseColmapDataset
