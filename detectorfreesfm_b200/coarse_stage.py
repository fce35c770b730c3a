"""The coarse-matching stage end to end on the GPU: rows a1-a2 of SURVEY.md section 8 around HP-1, plus 8(f) rows 1 and 3.

Mirrors ``match_worker`` (src/coarse_match/coarse_match_worker.py:103-145: dataset -> per-pair ``detector(data); matcher(data)``
-> ``(M,5)`` arrays keyed ``"path0<split>path1"``, optional grid rounding :133-136) and the part of
``detector_free_coarse_matching`` between the matcher and the h5 files (src/coarse_match/coarse_match.py:190-237), with

  * the image pipeline of image_pipeline.py (every image decoded and resized once, device resident),
  * any matcher honouring the HP-1 contract (``matcher(data)`` adds mkpts0_f / mkpts1_f / mconf / m_bids; B200LoFTR here),
  * the post-processing of postprocess.py on the match arrays while they are still on the device.

Nothing leaves the GPU between the decoded uint8 image and the final keypoint / match-index arrays except the dict
bookkeeping; the outputs have the reference's types, so ``save_h5`` (coarse_match.py:239-256) applies unchanged.
"""
import os
import threading

import torch

from . import _lib
from .image_pipeline import B200CoarseMatchingDataset
from .postprocess import KeypointMerger


@torch.no_grad()
def match_worker(subset_ids, image_lists, covis_pairs_out, cfgs, matcher, detector=None, keep_on_device=False, dataset=None):
    """coarse_match_worker.py:103-145 with a ready-built ``matcher`` (build_model is the caller's, see plugin.py).

    -> ``{pair_key: (M,5)}`` of [x0, y0, x1, y1, conf]: numpy float32 arrays like the reference, or CUDA tensors when
    ``keep_on_device`` (what ``merge_keypoints`` consumes without a host round trip)."""
    args = cfgs["matcher"]
    if dataset is None:
        dataset = B200CoarseMatchingDataset(cfgs["data"], image_lists, covis_pairs_out, subset_ids)
    loader = torch.utils.data.DataLoader(dataset, num_workers=0)      # items are CUDA tensors (no worker processes)
    # :134 tests ``args['model']['type'] is not 'coarse_only'`` -- an IDENTITY test against a literal, true for every run-time
    # string (hydra hands over a fresh str object), so the reference rounds whenever a ratio is configured, whatever the match
    # type.  The shipped coarse_only configs set the ratio to null; coarse points are multiples of 8 px anyway.
    rounding = args.get("round_matches_ratio") is not None
    matches = {}
    for data in loader:
        f_name0, f_name1 = data["pair_key"][0][0], data["pair_key"][1][0]
        data_c = {k: v.cuda() if isinstance(v, torch.Tensor) else v for k, v in data.items()}
        if detector is not None:
            detector(data_c)
        matcher(data_c)
        assert bool((data_c["m_bids"] == 0).all())                    # extract_preds: bs == 1
        mkpts0, mkpts1, mconf = data_c["mkpts0_f"], data_c["mkpts1_f"], data_c["mconf"]
        if rounding:                                                  # round to the grid for the later feature tracks
            r = args["round_matches_ratio"]
            s0, s1 = data_c["scale0"][:, [1, 0]], data_c["scale1"][:, [1, 0]]
            mkpts0 = torch.round((mkpts0 / s0) / r) * r * s0
            mkpts1 = torch.round((mkpts1 / s1) / r) * r * s1
        m = torch.cat([mkpts0, mkpts1, mconf[:, None]], -1).float()   # (N, 5)
        matches[args["pair_name_split"].join([f_name0, f_name1])] = m if keep_on_device else m.cpu().numpy()
    return matches


def pool_thread_begin(n_workers):
    """Called by a host thread before it drives one of ``n_workers`` pair workers of a GPU: programmatic dependent launch off for this
    thread's launches when n_workers > 1 (include/dfsfm_b200.h ``dfsfm_thread_set_pdl``; DFSFM_POOL_PDL=1 keeps it on, A/B)."""
    if n_workers > 1 and os.environ.get("DFSFM_POOL_PDL", "0") != "1":
        _lib.load_library().dfsfm_thread_set_pdl(0)


def pool_thread_end():
    _lib.load_library().dfsfm_thread_set_pdl(-1)


def match_workers(all_subset_ids, image_lists, covis_pairs_out, cfgs, matchers, detector=None, keep_on_device=False, datasets=None):
    """coarse_match.py:126-140 on ONE GPU: the reference deals the pair ids into ``n_workers`` subsets (``chunk_index``), runs one Ray
    actor -- its own model, ``n_gpus_per_worker`` of a GPU -- per subset and ChainMaps the result dicts.  Here every subset gets one
    host thread driving ``match_worker`` with its OWN matcher (own engine handle and workspaces) on its OWN CUDA stream of the current
    device; the threads spend their time inside the C ABI and CUDA calls (GIL released), so the pairs of different workers overlap on the
    GPU.  At 832x832 a single pair leaves ~40 % of the SMs idle in the transformer (43 / 86 tiles for 74 CTA pairs, DESIGN.md section
    6); two workers measured +15-20 % pairs/s, three +27 %.  (Ordering the workers' backbone phases with events -- one worker's backbone
    always beside another's transformer -- was tried and measured WORSE, 204-243 vs 229-255 pairs/s: profiles/r02_worker_pool.txt.)

    OPEN ISSUE: with three workers at 832x832 about one bench invocation in five stopped making progress on the device
    (profiles/r02_pool_hang.txt; never with one stream, never at small sizes, not reproduced with two workers) -- the pool is an opt-in
    until that is understood; bench.py defaults to one worker.

    ``matchers``: one per subset (e.g. ``[build_model(cfg) for _ in range(n)]``).  ``datasets``: optional, one per subset.
    Results are identical to one worker's (every pair is computed independently); key collisions resolve like ``dict(ChainMap(*results))``.
    """
    n = len(all_subset_ids)
    assert len(matchers) == n and (datasets is None or len(datasets) == n)
    assert len({id(m) for m in matchers}) == n, "every worker needs its own matcher (workspaces are per engine handle)"
    dev = torch.cuda.current_device()
    main = torch.cuda.current_stream(dev)
    streams = [main] + [torch.cuda.Stream(dev) for _ in range(n - 1)]
    start = torch.cuda.Event()
    start.record(main)
    results, errors = [None] * n, [None] * n

    def run(w):
        try:
            torch.cuda.set_device(dev)
            pool_thread_begin(n)
            streams[w].wait_event(start)
            with torch.cuda.stream(streams[w]):
                results[w] = match_worker(all_subset_ids[w], image_lists, covis_pairs_out, cfgs, matchers[w], detector=detector,
                                          keep_on_device=keep_on_device, dataset=None if datasets is None else datasets[w])
        except BaseException as e:  # re-raised in the calling thread
            errors[w] = e
        finally:
            pool_thread_end()

    threads = [threading.Thread(target=run, args=(w,)) for w in range(1, n)]
    for t in threads:
        t.start()
    run(0)
    for t in threads:
        t.join()
    for w in range(1, n):
        main.wait_stream(streams[w])
    for e in errors:
        if e is not None:
            raise e
    merged = {}
    for w in range(n - 1, -1, -1):           # ChainMap: the first mapping wins
        if keep_on_device and w > 0:
            for v in results[w].values():    # allocated on the worker's stream, consumed on the caller's
                v.record_stream(main)
        merged.update(results[w])
    return merged


def coarse_matching_stage(image_lists, covis_pairs, cfgs, matcher, detector=None, merger=None):
    """coarse_match.py:190-237 without Ray and without the h5 cache: all pairs -> (final_keypoints, final_scores,
    updated_matches, raw matches).  ``covis_pairs``: list of "path0 path1" strings or a file of them.  ``matcher``: one matcher, or a
    list of them = that many pair workers on this GPU (``match_workers``)."""
    if isinstance(covis_pairs, list):
        pair_list = covis_pairs
    else:
        with open(covis_pairs, "r") as f:
            pair_list = f.read().rstrip("\n").split("\n")
    if isinstance(matcher, (list, tuple)) and len(matcher) > 1:    # several workers on this GPU (ray n_workers / n_gpus_per_worker < 1)
        subsets = [list(range(w, len(pair_list), len(matcher))) for w in range(len(matcher))]
        matches = match_workers(subsets, image_lists, pair_list, cfgs, list(matcher), detector=detector, keep_on_device=True)
    else:
        one = matcher[0] if isinstance(matcher, (list, tuple)) else matcher
        matches = match_worker(list(range(len(pair_list))), image_lists, pair_list, cfgs, one, detector=detector, keep_on_device=True)
    merger = merger if merger is not None else KeypointMerger()
    final_keypoints, final_scores, updated_matches = merger(matches, image_lists, cfgs["matcher"]["pair_name_split"])
    return final_keypoints, final_scores, updated_matches, matches
