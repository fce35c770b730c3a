"""CPU checks of the host side: the C-ABI library loads and exports every symbol include/dfsfm_b200.h declares, weight
packing, the no-fallback behaviour, and the multi-rank partition + gather on gloo (world_size 2)."""
import os
import re

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    from detectorfreesfm_b200 import _lib
    header = open(os.path.join(ROOT, "include", "dfsfm_b200.h")).read()
    declared = set(re.findall(r"\b(dfsfm_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes and header disagree"
    assert lib.dfsfm_version() == 1


def test_no_cpu_fallback():
    """The product path must fail loudly without a CUDA device instead of computing on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from detectorfreesfm_b200 import B200LoFTR, DfsfmError
    from tests import util
    m = B200LoFTR(util.loftr_config())
    with pytest.raises(DfsfmError):
        m({"image0": torch.zeros(1, 1, 64, 64), "image1": torch.zeros(1, 1, 64, 64)})
    import ctypes
    from detectorfreesfm_b200 import _lib
    h = ctypes.c_void_p()
    rc = _lib.load_library().dfsfm_coarse_create(ctypes.byref(h), 0)
    assert rc != 0 and _lib.load_library().dfsfm_last_error()


def test_unsupported_configs_are_rejected():
    from detectorfreesfm_b200 import B200LoFTR
    from tests import util
    cfg = util.loftr_config()
    cfg["match_coarse"]["match_type"] = "sinkhorn"
    with pytest.raises(NotImplementedError):
        B200LoFTR(cfg)
    cfg = util.loftr_config(fine=True)
    cfg["fine_window_size"] = 7
    with pytest.raises(NotImplementedError):
        B200LoFTR(cfg)


def test_pack_loftr_shapes_and_bn_fold():
    from detectorfreesfm_b200.packing import pack_loftr, position_encoding
    from oracle import loftr_oracle as lo
    from oracle import weights
    sd = weights.loftr_state_dict(0)
    p = pack_loftr({("matcher." + k): v for k, v in sd.items()})  # checkpoint-style prefix is stripped
    assert p["l2.0.c2.w"][0].shape == (208, 10 * 208) and p["l3.0.c2.w"][0].shape == (256, 10 * 256)
    assert p["l2.0.c1.w"][0].shape == (208, 9 * 128) and p["tr.0.qkv"][0].shape == (768, 256)
    assert p["l2.0.c1.w"][0][196:].abs().max() == 0  # padded output channels are zero
    # folded stem == conv + BN of the oracle
    x = torch.rand(1, 1, 16, 16)
    ref = torch.relu(lo._bn(torch.nn.functional.conv2d(x, sd["backbone.conv1.weight"], None, 2, 3), sd, "backbone.bn1"))
    w = p["stem.w"][0].view(128, 1, 7, 7)
    out = torch.relu(torch.nn.functional.conv2d(x, w, p["stem.b"][0].view(-1), 2, 3))
    assert (out - ref).abs().max().item() < 1e-5
    assert torch.equal(position_encoding(5, 7).t().reshape(256, 5, 7), lo.position_encoding_sine(256, 5, 7))


def test_pack_multiview_shapes():
    from detectorfreesfm_b200.packing import pack_multiview
    from oracle import weights
    p = pack_multiview(weights.multiview_state_dict(0))
    assert p["c11.w"][0].shape == (64, 27) and p["c12.w"][0].shape == (64, 576) and p["c33.w"][0].shape == (256, 2304)
    assert p["a0.2.w"][0].shape == (128, 1600) and p["a1.0.w"][0].shape == (64, 256) and p["tr.3.mlp2"][0].shape == (128, 256)


def _gloo_worker(rank, world, port, q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank)})
    from detectorfreesfm_b200 import dist as D
    r, w, _ = D.init_from_env(backend="gloo")
    mine = D.shard(7, r, w)
    local = [torch.full((i + 1, 5), float(i)) for i in mine]  # unit i has i+1 rows filled with i
    got = D.gather_varlen(local)
    t = D.max_over_ranks(1.0 + r, torch.device("cpu"))
    if r == 0:
        flat = sorted((int(t_[0, 0]), t_.shape[0]) for per_rank in got for t_ in per_rank)
        q.put((flat, t, [len(x) for x in got]))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_and_gather_gloo_world2():
    port = 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat, t, lens = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert flat == [(i, i + 1) for i in range(7)]  # every unit arrived once, with its rows
    assert t == 2.0 and lens == [4, 3]


def test_refinement_window_rescale_matches_reference_rule():
    """multiview_match_worker.py:20-34: window 15 -> 11 -> 7 (floor 7), left window 7 -> 3 (floor 3) per refinement iteration."""
    from detectorfreesfm_b200.plugin import rescale_windows
    from tests.test_refine_gpu import multiview_config
    cfg = multiview_config(15, 7)
    got = []
    for factor in (None, 0, 2, 4, 6):
        r = rescale_windows(cfg, factor)
        got.append((r["multiview_transform"]["window_size"], r["backbone"]["s2dnet"]["window_size"],
                    r["multiview_matching_test"]["window_size"], r["multiview_matching_test"]["left_point_movement_window_size"]))
    assert got == [(15, 15, 15, 7), (15, 15, 15, 7), (11, 11, 11, 3), (7, 7, 7, 3), (7, 7, 7, 3)]
    assert cfg["multiview_transform"]["window_size"] == 15  # the input config is not mutated


def test_shard_covers_every_unit_once():
    from detectorfreesfm_b200.dist import shard
    for n, world in ((5050, 8), (28, 8), (3, 8), (500, 4)):
        seen = sorted(i for r in range(world) for i in shard(n, r, world))
        assert seen == list(range(n))
        sizes = [len(shard(n, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
