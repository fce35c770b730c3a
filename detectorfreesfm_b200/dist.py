"""One process per GPU: static partition of the independent work units + one variable-length gather to rank 0.

Replaces the reference's Ray fan-out (src/coarse_match/coarse_match.py:128-140: ``chunk_index`` + ``ray.get`` +
``ChainMap``; src/post_optimization/matcher_model/multiview_match.py:40-62).  Image pairs and track chunks never
interact, so there is no collective on the data path; the only exchange is the final gather of ``(M,5)`` match arrays
/ ``[K,4]`` refined-keypoint arrays (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard(n_items, rank, world):
    """Strided partition (cost-balanced when neighbouring items have similar cost): indices owned by ``rank``."""
    return list(range(rank, n_items, world))


def gather_varlen(local, dst=0):
    """Gather a list of 2-D fp32 tensors [m_k, C] (one per local work unit) to ``dst``.

    Returns on dst: list (rank order) of lists of tensors; elsewhere: None.  Two collectives: all_gather of the
    per-unit row counts, then all_gather of one packed, max-padded [rows, C] buffer per rank.
    """
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [list(local)]
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = local[0].device if len(local) else torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    C = local[0].shape[1] if len(local) else 0
    meta = torch.tensor([len(local), C], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    n_units = [int(m[0]) for m in metas]
    C = max(int(m[1]) for m in metas)
    max_units = max(n_units) if n_units else 0
    counts = torch.zeros(max(max_units, 1), dtype=torch.int64, device=dev)
    for i, t in enumerate(local):
        counts[i] = t.shape[0]
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    totals = [int(c[:n].sum()) for c, n in zip(all_counts, n_units)]
    max_rows = max(max(totals), 1)
    packed = torch.zeros(max_rows, max(C, 1), dtype=torch.float32, device=dev)
    if len(local) and totals[rank] > 0:
        packed[:totals[rank], :C] = torch.cat([t.to(torch.float32) for t in local], 0)
    bufs = [torch.zeros_like(packed) for _ in range(world)]
    dist.all_gather(bufs, packed)
    if rank != dst:
        return None
    out = []
    for r in range(world):
        rows = all_counts[r][:n_units[r]].tolist()
        out.append(list(torch.split(bufs[r][:totals[r], :C], rows)) if n_units[r] else [])
    return out


def gather_varlen_to(local, dst=0):
    """Like gather_varlen, but nothing travels to the ranks that do not need it: the per-unit row counts are all-gathered (a few
    hundred bytes), then every rank sends ONE packed [rows, C] buffer of its exact size to ``dst`` (point-to-point over NCCL / gloo).
    Returns on dst: list (rank order) of lists of tensors; elsewhere: None."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [list(local)]
    world, rank = dist.get_world_size(), dist.get_rank()
    nccl = dist.get_backend() == "nccl"
    dev = local[0].device if len(local) else (torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu"))
    C = local[0].shape[1] if len(local) else 0
    counts = [int(t.shape[0]) for t in local]
    metas = [None] * world
    dist.all_gather_object(metas, (C, counts))
    C = max(m[0] for m in metas)
    if rank != dst:
        if sum(counts) > 0:
            dist.send(torch.cat([t.to(torch.float32) for t in local], 0).contiguous(), dst)
        return None
    out = []
    for r in range(world):
        rows = metas[r][1]
        if r == rank:
            out.append([t.to(torch.float32) for t in local])
            continue
        total = sum(rows)
        buf = torch.empty((total, C), dtype=torch.float32, device=dev)
        if total > 0:
            dist.recv(buf, r)
        out.append(list(torch.split(buf, rows)) if rows else [])
    return out


def shard_pairs_by_image(pairs, rank, world):
    """Locality-aware partition of a pair list: pairs are grouped by their FIRST image and the groups are dealt to the ranks
    as contiguous image ranges of (nearly) equal pair count, so that a rank's per-image feature cache sees every first image once and
    the second images of its range once each.  -> indices into ``pairs`` owned by ``rank`` (ascending)."""
    first = sorted({p[0] for p in pairs})
    per_img = {i: 0 for i in first}
    for p in pairs:
        per_img[p[0]] += 1
    target = len(pairs) / float(world)
    owner, acc, r = {}, 0, 0
    for i in first:
        if r < world - 1 and acc >= target * (r + 1) - per_img[i] / 2.0:
            r += 1
        owner[i] = r
        acc += per_img[i]
    return [k for k, p in enumerate(pairs) if owner[p[0]] == rank]


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value, device):
    """max over ranks of a python float (timings are reported as the slowest rank's)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
