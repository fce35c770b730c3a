"""__graft_entry__.smoke(): one small invocation of each hot path on cuda:0, checked against the CPU oracle."""
import torch


def run():
    assert torch.cuda.is_available(), "smoke() needs a CUDA device"
    from detectorfreesfm_b200 import B200LoFTR, B200MultiviewMatcher
    from oracle import loftr_oracle as lo
    from oracle import multiview_oracle as mo
    from tests import weights
    from tests import util
    from tests.util import multiview_config, to_cuda

    # HP-1: one 96x128 pair through the plugin interface
    sd = weights.loftr_state_dict(0)
    m = B200LoFTR(util.loftr_config(thr=0.2, temperature=0.01)).cuda(0).eval()
    m.load_state_dict(sd)
    im0, im1 = util.synth_pair(96, 128, seed=1)
    ref = lo.loftr_forward({"image0": im0, "image1": im1}, sd, {"thr": 0.2, "temperature": 0.01}, keep=True)
    data = {"image0": im0.cuda(), "image1": im1.cuda(), "_return_conf_matrix": True}
    m(data)
    err = (data["conf_matrix"].cpu() - ref["conf_matrix"]).abs().max().item()
    assert err < 1e-3, f"confidence parity {err}"
    assert torch.equal(data["i_ids"].cpu(), ref["i_ids"]) and torch.equal(data["j_ids"].cpu(), ref["j_ids"])
    print(f"[smoke] HP-1 coarse match: {len(ref['i_ids'])} matches, max |dconf| = {err:.2e}")

    # HP-2: one 16-track chunk
    sdm = weights.multiview_state_dict(0)
    chunk = util.synth_chunk(M=16, n_img=4, max_views=3, seed=4)
    refm = mo.multiview_forward(chunk, sdm, 15, 7)
    rm = B200MultiviewMatcher(multiview_config(15, 7), test=True).cuda(0).eval()
    rm.load_state_dict(sdm)
    d = to_cuda(chunk)
    rm(d)
    mask = chunk["track_valid_mask"]
    dq = (d["query_points_refined"].cpu() - refm["query_points_refined"]).abs().max().item()
    dr = (d["reference_points_refined"][-1].cpu() - refm["reference_points_refined"])[mask].abs().max().item()
    assert dq < 1e-3 and dr < 0.1, (dq, dr)
    print(f"[smoke] HP-2 refinement chunk: max |d query| = {dq:.2e} px, max |d refined| = {dr:.2e} px")

    # post-processing: 4 images, 6 pairs, bit-exact vs the numpy oracle
    import itertools
    import numpy as np
    from detectorfreesfm_b200 import merge_keypoints
    from oracle import postprocess_oracle as po
    matches, names = util.synth_matches(4, list(itertools.combinations(range(4), 2)), 120, seed=2)
    ref_k, ref_s, ref_m = po.merge_keypoints(matches, names, " ")
    out_k, out_s, out_m = merge_keypoints(matches, names, " ")
    for n in names:
        assert np.array_equal(out_k[n], ref_k[n]) and np.array_equal(out_s[n], ref_s[n])
    for k in matches:
        assert np.array_equal(out_m[k], ref_m[k])
    print(f"[smoke] post-processing: {sum(v.shape[0] for v in out_k.values())} key points from {sum(v.shape[0] for v in matches.values())} matches, bit-exact")

    # image pipeline: GPU Lanczos resize + /255 vs PIL (the reference's resize), bit-exact
    from PIL import Image
    from detectorfreesfm_b200 import GpuImageReader
    from oracle import image_oracle as imo
    photo = util.synth_photo(300, 400, 1)
    ref_img = np.asarray(Image.fromarray(photo).resize((160, 120), resample=Image.LANCZOS)).astype("float32") / 255.
    out_img = GpuImageReader().resize_gray(photo, (160, 120)).cpu().numpy()
    assert np.array_equal(out_img, ref_img)
    print("[smoke] image pipeline: 300x400 -> 120x160 PIL-LANCZOS parity, bit-exact")

    # chunk dataset: native bag assignment + chunk dicts feeding the refinement matcher
    from detectorfreesfm_b200 import B200MatchingMultiviewData
    from detectorfreesfm_b200 import refine_stage as rs
    ds = util.SynthColmapDataset(n_images=10, n_points=60, max_obs=6, seed=3, hw=(96, 128), dup_frac=0.0)
    chunks = B200MatchingMultiviewData(ds, {"max_track_length": 16, "chunk": 40})
    res = rs.match_worker(torch.utils.data.DataLoader(chunks, num_workers=0), rm, list(ds.colmap_images.keys()))
    nodes = sum(r.shape[0] for r in res)
    assert nodes >= sum(len(set(p.image_ids.tolist())) for p in ds.colmap_3ds.values()) and all(np.isfinite(r).all() for r in res)
    print(f"[smoke] chunk dataset -> refinement worker loop: {len(chunks)} chunks, {nodes} refined nodes")
