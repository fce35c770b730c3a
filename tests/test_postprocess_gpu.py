"""SURVEY 8(f) row 1 on the GPU: csrc/postprocess.cu behind dfsfm_post_merge_keypoints vs the CPU oracle (bit-exact: integer ids,
truncated coordinates, float32 casts of float64 sums accumulated in the reference's order) and vs the committed reference
outputs; size-independent properties at a larger size."""
import itertools
import os

import numpy as np
import pytest
import torch

from tests import util  # noqa: E402

from oracle import postprocess_oracle as po

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _merger():
    from detectorfreesfm_b200 import KeypointMerger
    return KeypointMerger()


def _assert_equal(out, ref, names, keys):
    for name in names:
        assert out[0][name].shape == ref[0][name].shape and out[0][name].dtype == ref[0][name].dtype, name
        assert np.array_equal(out[0][name], ref[0][name]), name
        assert out[1][name].dtype == np.float32 and np.array_equal(out[1][name], ref[1][name]), name
    for k in keys:
        assert out[2][k].shape == ref[2][k].shape and out[2][k].dtype == ref[2][k].dtype, k
        assert np.array_equal(out[2][k], ref[2][k]), k


@pytest.mark.parametrize("case", [0, 1, 2, 3, 4, 5])
def test_merge_keypoints_matches_oracle(case):
    n, m, dup = [(4, 50, 0.3), (6, 300, 0.3), (5, [0, 10, 200], 0.5), (3, 1, 0.0), (9, 3000, 0.2), (2, 70000, 0.9)][case]
    pairs = list(itertools.combinations(range(n), 2))
    if case == 2:
        pairs = [p for p in pairs if 4 not in p]          # image 4 never matched; some pairs empty
    matches, names = util.synth_matches(n, pairs, m, seed=case, dup=dup)
    ref = po.merge_keypoints(matches, names, " ")
    mg = _merger()
    _assert_equal(mg(matches, names, " "), ref, names, matches.keys())
    # CUDA-tensor inputs (the matcher's own outputs) take the same path
    cuda_matches = {k: torch.from_numpy(v).cuda() for k, v in matches.items()}
    _assert_equal(mg(cuda_matches, names, " "), ref, names, matches.keys())


def test_reference_golden_vector():
    g = torch.load(os.path.join(GOLD, "postprocess_small.pt"), weights_only=False)
    out = _merger()(g["matches"], g["names"], " ")
    _assert_equal(out, (g["final_keypoints"], g["final_scores"], g["updated_matches"]), g["names"], g["matches"].keys())


def test_no_matches_at_all_and_separator_fallback():
    names = ["a", "b", "c"]
    mg = _merger()
    empty = {"a b": np.empty((0, 5), dtype=np.float32), "b c": np.empty((0, 5), dtype=np.float32)}
    fk, fs, upd = mg(empty, names, " ")
    ref = po.merge_keypoints(empty, names, " ")
    _assert_equal((fk, fs, upd), ref, names, empty.keys())
    assert all(v.shape == (0, 2) for v in fk.values()) and all(v.shape == (0, 2) for v in upd.values())
    # merge_kpts.py:27-30: keys that do not split on the configured separator split on '-'
    m = {"a-b": np.array([[1.5, 2.5, 3.5, 4.5, 0.5], [1.2, 2.9, 7.0, 8.0, 0.25]], dtype=np.float32)}
    _assert_equal(mg(m, names, " "), po.merge_keypoints(m, names, " "), names, m.keys())


def test_rejects_negative_coordinates():
    from detectorfreesfm_b200 import DfsfmError
    m = {"a b": np.array([[-1.0, 2.0, 3.0, 4.0, 0.5]], dtype=np.float32)}
    with pytest.raises(DfsfmError):
        _merger()(m, ["a", "b"], " ")


def test_large_properties():
    """64 images, all 2016 pairs, ~1000 matches each (4 M observations): size-independent properties instead of the oracle."""
    n = 64
    pairs = list(itertools.combinations(range(n), 2))
    matches, names = util.synth_matches(n, pairs, 1000, seed=11, dup=0.25)
    fk, fs, upd = _merger()(matches, names, " ")
    index = {nm: i for i, nm in enumerate(names)}
    total_conf = 0.0
    for k, v in matches.items():
        n0, n1 = k.split(" ")
        ids = upd[k]
        assert ids.shape == (v.shape[0], 2) and ids.min() >= 0
        # every match points at the key point carrying its truncated coordinates
        assert np.array_equal(fk[n0][ids[:, 0]], np.trunc(v[:, 0:2])) and np.array_equal(fk[n1][ids[:, 1]], np.trunc(v[:, 2:4]))
        total_conf += 2.0 * float(v[:, 4].astype(np.float64).sum())
    got = 0.0
    for nm in names:
        s = fs[nm]
        assert np.all(s[:-1] >= s[1:])                                   # ranked by score
        assert np.unique(fk[nm], axis=0).shape[0] == fk[nm].shape[0]     # key points are unique
        got += float(s.astype(np.float64).sum())
    assert abs(got - total_conf) / total_conf < 1e-6                     # confidence mass is conserved
    # one image against the oracle (full check of a slice of the big problem)
    sub = {k: v for k, v in matches.items() if index[k.split(" ")[0]] < 3 and index[k.split(" ")[1]] < 3}
    ref = po.merge_keypoints(sub, names[:3], " ")
    out = _merger()(sub, names[:3], " ")
    _assert_equal(out, ref, names[:3], sub.keys())
