"""HP-2 parity on the GPU: B200MultiviewMatcher (C ABI, sm_100a kernels) vs the CPU oracle (oracle/multiview_oracle.py).
Tolerance (north_star): refined keypoints within 0.1 px; we assert 1e-2 px and identical reference-point moves."""
import pytest
import torch

from oracle import build_native
from oracle import multiview_oracle as mo
from tests import weights
from tests import util

pytestmark = pytest.mark.gpu


multiview_config, to_cuda = util.multiview_config, util.to_cuda   # shared with bench.py / smoke (generators live in tests/util.py)


def test_crop_and_resize_matches_oracle(lib):
    """L0 op through the C ABI vs the C restatement of the reference's crop_and_resize (bit-exact expected)."""
    from detectorfreesfm_b200 import _lib
    g = torch.Generator().manual_seed(0)
    image = torch.rand(2, 3, 60, 80, generator=g)
    n = 50
    ctr = torch.rand(n, 2, generator=g) * torch.tensor([80., 60.]) * 1.2 - 8
    boxes_xyxy = torch.cat([ctr - 17, ctr + 17], 1)
    x1, y1, x2, y2 = boxes_xyxy.split(1, 1)
    nb = torch.cat([y1 / 59., x1 / 79., y2 / 59., x2 / 79.], 1).contiguous()
    bi = torch.randint(0, 2, (n,), generator=g, dtype=torch.int32)
    ref = build_native.roialign_forward(image, nb, bi, 35, 35)
    crops = torch.empty(n, 3, 35, 35, device="cuda")
    img_d, nb_d, bi_d = image.cuda(), nb.cuda(), bi.cuda()
    _lib.check(lib.dfsfm_crop_and_resize_forward(_lib.ptr(img_d), 2, 3, 60, 80, _lib.ptr(nb_d), _lib.ptr(bi_d), n, 0.0, 35, 35, _lib.ptr(crops), None))
    torch.cuda.synchronize()
    # same float expression tree with explicit round-to-nearest mul/add on the GPU -> bit-exact
    assert torch.equal(crops.cpu(), ref)


@pytest.mark.parametrize("W,LW,M,seed", [(15, 7, 48, 2), (11, 3, 40, 5), (15, 7, 130, 7), (7, 3, 33, 9), (15, 7, 1, 3)])
def test_refine_chunk(W, LW, M, seed):
    from detectorfreesfm_b200 import B200MultiviewMatcher
    sd = weights.multiview_state_dict(0)
    data = util.synth_chunk(M=M, n_img=6, max_views=5, seed=seed)
    ref = mo.multiview_forward(data, sd, W, LW)
    m = B200MultiviewMatcher(multiview_config(W, LW), test=True).cuda().eval()
    m.load_state_dict(sd)
    d = to_cuda(data)
    m(d)
    mask = data["track_valid_mask"]
    q = d["query_points_refined"].cpu()
    r = d["reference_points_refined"][-1].cpu()
    s = d["std"][-1].cpu()
    assert q.shape == ref["query_points_refined"].shape and r.shape == ref["reference_points_refined"].shape
    dq = (q - ref["query_points_refined"]).abs().max().item()
    dr = (r - ref["reference_points_refined"])[mask].abs().max().item()
    ds = (s - ref["std"])[mask].abs().max().item()
    assert dq < 1e-3, f"reference-point move differs by {dq} px"
    assert dr < 1e-2, f"refined keypoints differ by {dr} px"
    assert ds < 1e-3, ds
    if (~mask).any():
        assert r[~mask].abs().max().item() == 0 and s[~mask].abs().max().item() == 0
