"""Full-size (BASELINE.json configs[1] / configs[2]) checks through size-independent properties: the oracle is too slow at
these sizes, so the CUDA path is checked against invariants of the algorithm and against itself."""
import pytest
import torch

from tests import weights
from tests import util
from tests.test_refine_gpu import multiview_config, to_cuda

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def matcher():
    from detectorfreesfm_b200 import B200LoFTR
    m = B200LoFTR(util.loftr_config(thr=0.002, temperature=0.01)).cuda().eval()  # random weights: low thr for a populated match set
    m.load_state_dict(weights.loftr_state_dict(0))
    return m


def test_c2_pair_832_selection_matches_dense_confidence(matcher):
    """832x832 pair (L = S = 10816): the fused dual-softmax / mutual-NN kernels must report exactly the entries that the
    reference rule selects on the dense confidence matrix the same kernels can optionally emit (coarse_matching.py:172-193),
    and that matrix must be a product of two softmaxes (rows and columns of the factors sum to one)."""
    im0, im1 = util.synth_pair(832, 832, seed=5, shift=(16, 24))
    data = {"image0": im0.cuda(), "image1": im1.cuda(), "_return_conf_matrix": True}
    matcher(data)
    conf = data["conf_matrix"][0]
    L, S = conf.shape
    assert (L, S) == (104 * 104, 104 * 104)
    # softmax structure: conf = P_row * P_col  =>  sum_j sqrt-free check via the two factors recovered from the features
    f0, f1 = data["feat_c0"][0], data["feat_c1"][0]
    sim = (f0 @ f1.t()) / 256.0 / 0.01
    ref = torch.softmax(sim, 0) * torch.softmax(sim, 1)
    assert (conf - ref).abs().max().item() < 1e-3
    del sim, ref
    # selection rule on the dense matrix
    mask = conf > matcher.thr
    mask = mask.view(104, 104, 104, 104).clone()
    mask[:2] = False; mask[:, :2] = False; mask[:, :, :2] = False; mask[:, :, :, :2] = False
    mask = mask.view(L, S)
    mask &= conf == conf.max(dim=1, keepdim=True)[0]
    mask &= conf == conf.max(dim=0, keepdim=True)[0]
    mv, jj = mask.max(dim=1)
    ii = torch.where(mv)[0]
    jj = jj[ii]
    assert len(ii) > 5, len(ii)
    assert torch.equal(data["i_ids"], ii) and torch.equal(data["j_ids"], jj)
    assert torch.equal(data["mconf"], conf[ii, jj])
    # one-to-one, inside the border, above threshold
    assert len(torch.unique(data["i_ids"])) == len(ii) and len(torch.unique(data["j_ids"])) == len(ii)
    assert (data["mconf"] > matcher.thr).all()
    assert (data["mkpts0_f"] >= 16).all() and (data["mkpts1_f"] >= 16).all()


def test_c2_determinism_and_feature_cache(matcher):
    """Same pair twice, with and without the per-image feature cache: bit-identical matches."""
    im0, im1 = util.synth_pair(832, 832, seed=6)
    outs = []
    for keyed in (False, False, True, True):
        data = {"image0": im0.cuda(), "image1": im1.cuda()}
        if keyed:
            data["pair_key"] = (("a.jpg",), ("b.jpg",))
        matcher(data)
        outs.append((data["i_ids"].clone(), data["j_ids"].clone(), data["mconf"].clone(), data["mkpts1_f"].clone()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert torch.equal(a, b)


def test_c3_chunk_2000_tracks_subset_invariance():
    """Refinement chunk at the shipped size (2000 tracks, <= 9 query views): tracks never interact, so refining a
    sub-chunk alone must reproduce the full chunk's values for those tracks bit for bit; padded slots stay zero and every
    refined point stays within the search window of its coarse location."""
    from detectorfreesfm_b200 import B200MultiviewMatcher
    m = B200MultiviewMatcher(multiview_config(15, 7), test=True).cuda().eval()
    m.load_state_dict(weights.multiview_state_dict(0))
    chunk = util.synth_chunk(M=2000, n_img=10, max_views=9, hw=(600, 800), seed=21, scales=torch.ones(1, 10, 2))
    full = to_cuda(chunk)
    m(full)
    q = full["query_points_refined"].cpu()
    r = full["reference_points_refined"][-1].cpu()
    s = full["std"][-1].cpu()
    mask = chunk["track_valid_mask"]
    assert torch.isfinite(q).all() and torch.isfinite(r).all() and torch.isfinite(s).all()
    assert r[~mask].abs().max().item() == 0 and s[~mask].abs().max().item() == 0
    assert (q - chunk["query_points"]).abs().max().item() <= 3.0 + 1e-4          # 7x7 reference window: +-3 px
    assert ((r - chunk["reference_points_coarse"])[mask]).abs().max().item() <= 7.0 + 1e-4   # 15x15 query window: +-7 px
    frozen = ~chunk["query_movable_mask"][0]
    assert torch.equal(q[0][frozen], chunk["query_points"][0][frozen])           # frozen reference points do not move
    # sub-chunk = a contiguous slice of tracks (still sorted by valid-view count)
    sl = slice(700, 900)
    sub = {k: v for k, v in chunk.items()}
    for k in ("query_points", "query_img_idxs", "query_movable_mask"):
        sub[k] = chunk[k][:, sl]
    for k in ("reference_points_coarse", "track_valid_mask", "reference_img_idxs", "scales_relative", "view_point_vector"):
        sub[k] = chunk[k][:, :, sl]
    subc = to_cuda(sub)
    m(subc)
    assert torch.equal(subc["query_points_refined"].cpu(), q[:, sl])
    assert torch.equal(subc["reference_points_refined"][-1].cpu(), r[:, :, sl])
    assert torch.equal(subc["std"][-1].cpu(), s[:, :, sl])
