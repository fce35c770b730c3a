"""Post-processing when the pairs are sharded over ranks (SURVEY.md 8(e) for the 8(f)-1 row).

The match -> keypoint merge groups end points by IMAGE across all pairs, so with pairs spread over ranks it needs one real
exchange: every (pair, side) half -- whose rows all lie on one image -- is sent to the rank owning that image
(``all_to_all_single`` of [x, y, conf] rows; NCCL over NVLink, gloo in the CPU test), the owner merges its images with the
single-GPU kernels (postprocess.py), and the keypoint ids travel back the same way.  Received halves are ordered by their global
pair index before the merge, so the float64 score sums are accumulated in exactly the reference's order and the result is
bit-identical to the single-process one, whatever the sharding.

The owner feeds the halves to the pair-based device entry point as pseudo pairs (image, sink): the sink image collects the
unused second end point (0, 0) and is dropped.  Status: exchange logic covered by tests/test_host_cpu.py on gloo (world 2, the
numpy oracle standing in for the device merge); the NCCL run is a round-2 item.
"""
import numpy as np
import torch
import torch.distributed as dist

from .postprocess import _split_pair


def image_owner(n_images, world):
    """Contiguous blocks of images per rank -> int array [n_images]."""
    per = -(-n_images // world)
    return np.minimum(np.arange(n_images) // per, world - 1)


def merge_keypoints_sharded(local_matches, global_pair_ids, image_lists, pair_name_split=" ", local_merge=None, device=None):
    """local_matches: this rank's ordered {pair_key: (M,5)}; global_pair_ids: position of each of them in the global matches
    dict (the reference's accumulation order).  Every rank must call it.  Returns (final_keypoints, final_scores) for ALL images
    on every rank and updated_matches for the LOCAL pairs."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    names = list(image_lists)
    index = {n: i for i, n in enumerate(names)}
    owner = image_owner(len(names), world)
    if local_merge is None:
        from .postprocess import KeypointMerger
        local_merge = KeypointMerger(device).merge
    keys = list(local_matches.keys())
    assert len(keys) == len(global_pair_ids)
    tens = []
    for k in keys:
        v = local_matches[k]
        v = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
        tens.append(v.float() if device is None else v.to(device, torch.float32))
    dev = tens[0].device if tens else (torch.device("cpu") if device is None else torch.device(device))

    # ---- halves to send: (destination, global pair id, side, image, rows)
    halves = []
    for li, k in enumerate(keys):
        n0, n1 = _split_pair(k, pair_name_split)
        for side, n in ((0, n0), (1, n1)):
            img = index[n]
            halves.append((int(owner[img]), int(global_pair_ids[li]), side, img, li))
    order = sorted(range(len(halves)), key=lambda h: halves[h][0])            # stable: destination-major, local order inside
    send_meta = [[] for _ in range(world)]                                     # per destination: (gpid, side, image, count)
    send_rows = []
    for h in order:
        d, gpid, side, img, li = halves[h]
        m = tens[li]
        send_meta[d].append((gpid, side, img, int(m.shape[0])))
        send_rows.append(m[:, [2 * side, 2 * side + 1, 4]])
    send = torch.cat(send_rows, 0).contiguous() if send_rows else torch.empty((0, 3), dtype=torch.float32, device=dev)
    in_split = [sum(c for *_, c in send_meta[d]) for d in range(world)]

    # ---- exchange the metadata (small python lists), then the rows
    if world > 1:
        all_meta = [None] * world
        dist.all_gather_object(all_meta, send_meta)
    else:
        all_meta = [send_meta]
    recv_meta = [all_meta[src][rank] for src in range(world)]                  # what every source sends to this rank, in order
    out_split = [sum(c for *_, c in recv_meta[src]) for src in range(world)]
    recv = torch.empty((sum(out_split), 3), dtype=torch.float32, device=dev)
    if world > 1:
        dist.all_to_all_single(recv, send, out_split, in_split)
    else:
        recv.copy_(send)

    # ---- owner side: order the received halves by global pair id (the reference's concatenation order), merge
    segs = []
    off = 0
    for src in range(world):
        for (gpid, side, img, cnt) in recv_meta[src]:
            segs.append((gpid, side, img, cnt, off))
            off += cnt
    by_pair = sorted(range(len(segs)), key=lambda s: (segs[s][0], segs[s][1]))
    own_imgs = [i for i in range(len(names)) if owner[i] == rank]
    local_idx = {g: l for l, g in enumerate(own_imgs)}
    sink = len(own_imgs)
    perm = torch.cat([torch.arange(segs[s][4], segs[s][4] + segs[s][3], device=dev) for s in by_pair]) if segs else torch.empty(0, dtype=torch.long, device=dev)
    rows3 = recv[perm]
    T = int(rows3.shape[0])
    rows5 = torch.zeros((T, 5), dtype=torch.float32, device=dev)
    rows5[:, 0:2] = rows3[:, 0:2]
    rows5[:, 4] = rows3[:, 2]
    counts = np.array([segs[s][3] for s in by_pair], dtype=np.int64)
    pair_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    pair_img = np.array([[local_idx[segs[s][2]], sink] for s in by_pair], dtype=np.int32).reshape(-1, 2)
    kpt_xy, kpt_score, img_off, ids = local_merge(rows5, pair_off, pair_img, sink + 1)
    img_off = img_off.cpu().numpy()
    ids0 = ids[:, 0].to(torch.int32) if T else torch.empty(0, dtype=torch.int32, device=dev)

    # ---- ids back to the senders (inverse permutation, mirrored split sizes)
    back = torch.empty(T, dtype=torch.int32, device=dev)
    if T:
        back[perm] = ids0
    got = torch.empty(sum(in_split), dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_to_all_single(got, back, in_split, out_split)
    else:
        got.copy_(back)
    updated = {k: np.empty((int(t.shape[0]), 2), dtype=np.int64) for k, t in zip(keys, tens)}
    got_np = got.cpu().numpy()
    off = 0
    for h in order:
        _, _, side, _, li = halves[h]
        m = int(tens[li].shape[0])
        updated[keys[li]][:, side] = got_np[off:off + m]
        off += m

    # ---- key points of the owned images -> every rank
    kx, ks = kpt_xy.cpu().numpy(), kpt_score.cpu().numpy()
    mine = {}
    for l, g in enumerate(own_imgs):
        a, b = int(img_off[l]), int(img_off[l + 1])
        mine[names[g]] = (kx[a:b].copy() if b > a else np.empty((0, 2)), ks[a:b].copy())
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    else:
        parts = [mine]
    final_kpts, final_scores = {}, {}
    for n in names:
        for part in parts:
            if n in part:
                final_kpts[n], final_scores[n] = part[n]
    return final_kpts, final_scores, updated
