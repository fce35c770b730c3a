"""CPU checks of the host side: the C-ABI library loads and exports every symbol include/dfsfm_b200.h declares, weight
packing, the no-fallback behaviour, and the multi-rank partition + gather on gloo (world_size 2)."""
import os
import re

import pytest
import torch

from tests import util  # noqa: E402
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
    # the section-8(f) rows have no CPU path either
    import numpy as np
    from detectorfreesfm_b200 import GpuImageReader, KeypointMerger, merge_keypoints
    with pytest.raises(DfsfmError):
        KeypointMerger()
    with pytest.raises(DfsfmError):
        merge_keypoints({"a b": np.zeros((1, 5), dtype=np.float32)}, ["a", "b"], " ")
    with pytest.raises(DfsfmError):
        GpuImageReader()
    rc = _lib.load_library().dfsfm_post_create(ctypes.byref(h), 0)
    assert rc != 0


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
    from tests import weights
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
    from tests import weights
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
    got2 = D.gather_varlen_to(local + [torch.zeros((0, 5))], dst=0)       # point-to-point variant: only dst receives (incl. an empty unit)
    assert (got2 is None) == (r != 0)
    t = D.max_over_ranks(1.0 + r, torch.device("cpu"))
    if r == 0:
        flat = sorted((int(t_[0, 0]), t_.shape[0]) for per_rank in got for t_ in per_rank)
        flat2 = sorted((int(t_[0, 0]), t_.shape[0]) for per_rank in got2 for t_ in per_rank if t_.shape[0])
        assert flat2 == flat and [len(x) for x in got2] == [len(x) + 1 for x in got]
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
    from tests.util import multiview_config
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


def test_locality_partition_of_the_pair_graph():
    """dist.shard_pairs_by_image: every pair exactly once, balanced to a few %, and a rank's first images form one contiguous range"""
    import itertools
    from detectorfreesfm_b200.dist import shard_pairs_by_image
    pairs = list(itertools.combinations(range(101), 2))
    for world in (1, 2, 4, 8):
        parts = [shard_pairs_by_image(pairs, r, world) for r in range(world)]
        assert sorted(k for p in parts for k in p) == list(range(len(pairs)))
        sizes = [len(p) for p in parts]
        assert max(sizes) <= 1.15 * len(pairs) / world + 1
        for p in parts:
            firsts = sorted({pairs[k][0] for k in p})
            assert firsts == list(range(firsts[0], firsts[-1] + 1))


def _oracle_flat_merge(rows5, pair_off, pair_img, n_images):
    """Stand-in for KeypointMerger.merge on the CPU (tests only): the numpy oracle behind the same flat-array signature."""
    import numpy as np
    from oracle import postprocess_oracle as po
    rows = rows5.cpu().numpy()
    names = [f"i{i}" for i in range(n_images)]
    matches = {}
    for p in range(len(pair_img)):
        matches[f"{names[pair_img[p][0]]} {names[pair_img[p][1]]} #{p}"] = rows[pair_off[p]:pair_off[p + 1]]
    # one oracle call per image: its observations in pseudo-pair order form the first image of a two-image helper problem
    per_image = {n: [] for n in names}
    for key, v in matches.items():
        a, b, _ = key.split(" ")
        per_image[a].append((key, 0))
        per_image[b].append((key, 1))
    ids = {k: np.zeros((v.shape[0], 2), dtype=np.int32) for k, v in matches.items()}
    xy, sc, off = [], [], [0]
    for n in names:
        parts = [matches[k][:, [2 * s, 2 * s + 1, 4]] for k, s in per_image[n]]
        kp = np.concatenate(parts, 0) if parts else np.empty((0, 3), dtype=np.float32)
        if kp.shape[0]:
            single = po.merge_keypoints({"a b": np.concatenate([kp[:, :2], np.zeros_like(kp[:, :2]), kp[:, 2:3]], 1).astype(np.float32)}, ["a", "b"], " ")
            # image "a" of the helper call carries exactly this image's observations (image "b" is a sink)
            k_a, s_a, m_a = single[0]["a"], single[1]["a"], single[2]["a b"][:, 0]
            pos = 0
            for k, s in per_image[n]:
                m = matches[k].shape[0]
                ids[k][:, s] = m_a[pos:pos + m]
                pos += m
        else:
            k_a, s_a = np.empty((0, 2), dtype=np.float32), np.empty((0,), dtype=np.float32)
        xy.append(k_a.astype(np.float32).reshape(-1, 2))
        sc.append(s_a.astype(np.float32))
        off.append(off[-1] + k_a.shape[0])
    all_ids = np.concatenate([ids[k] for k in matches], 0) if matches else np.zeros((0, 2), dtype=np.int32)
    return (torch.from_numpy(np.concatenate(xy, 0)), torch.from_numpy(np.concatenate(sc, 0)), torch.tensor(off, dtype=torch.int32),
            torch.from_numpy(all_ids))


def _post_worker(rank, world, port, q):
    import itertools
    import numpy as np
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank)})
    from detectorfreesfm_b200 import dist as D
    from detectorfreesfm_b200.postprocess_dist import merge_keypoints_sharded
    from oracle import postprocess_oracle as po
    r, w, _ = D.init_from_env(backend="gloo")
    pairs = [p for p in itertools.combinations(range(6), 2) if 5 not in p]       # image 5 never matched
    matches, names = util.synth_matches(6, pairs, [0, 30, 120], seed=3, dup=0.4)
    keys = list(matches.keys())
    mine = D.shard(len(keys), r, w)                                               # strided, as the HP-1 bench shards pairs
    local = {keys[i]: matches[keys[i]] for i in mine}
    fk, fs, upd = merge_keypoints_sharded(local, mine, names, " ", local_merge=_oracle_flat_merge)
    ref = po.merge_keypoints(matches, names, " ")
    ok = all(np.array_equal(fk[n], ref[0][n]) and fk[n].shape == ref[0][n].shape and np.array_equal(fs[n], ref[1][n]) for n in names)
    ok = ok and all(np.array_equal(upd[k], ref[2][k]) and upd[k].dtype == ref[2][k].dtype for k in local)
    q.put((r, ok, len(local)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_postprocess_exchange_gloo_world2():
    """Pairs strided over two ranks, images owned in blocks: all-to-all of the (pair, side) halves, merge per owner, ids back.
    Bit-identical to the single-process oracle on every rank (scores summed in global pair order)."""
    port = 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_post_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True] and sum(r[2] for r in res) == 10


def test_sharded_postprocess_single_process_path():
    """Without torch.distributed the sharded entry point degenerates to the plain merge (same halves / sink-image plumbing)."""
    import itertools
    import numpy as np
    from detectorfreesfm_b200.postprocess_dist import image_owner, merge_keypoints_sharded
    from oracle import postprocess_oracle as po
    matches, names = util.synth_matches(4, list(itertools.combinations(range(4), 2)), 60, seed=9)
    fk, fs, upd = merge_keypoints_sharded(matches, list(range(len(matches))), names, " ", local_merge=_oracle_flat_merge)
    ref = po.merge_keypoints(matches, names, " ")
    assert all(np.array_equal(fk[n], ref[0][n]) and np.array_equal(fs[n], ref[1][n]) for n in names)
    assert all(np.array_equal(upd[k], ref[2][k]) for k in matches)
    assert image_owner(10, 4).tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3] and image_owner(3, 8).tolist() == [0, 1, 2]
