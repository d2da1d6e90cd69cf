"""Frame sharding across the GPUs of one node.

Frames are independent for detection and description; matching couples frame f only with f-1.  Each
rank therefore takes a contiguous range [lo, hi) of the sequence plus, when lo > 0, the one frame before
it, so that every consecutive-pair match is computed locally.  There is no collective on the data path:
the only communication is the gather of the small per-frame results (keypoints, descriptors, matches)
to rank 0, which torch.distributed does over NCCL on the GPUs or gloo in the CPU tests.
"""
import numpy as np


def shard_range(n_frames, rank, world):
    """Contiguous range of frames owned by `rank`: [lo, hi)."""
    lo = (n_frames * rank) // world
    hi = (n_frames * (rank + 1)) // world
    return lo, hi


def shard_plan(n_frames, world):
    """Per rank: (first frame to load, lo, hi).  Frames [first, hi) are processed; results for [lo, hi) and
    matches ending in [lo, hi) (pair (f-1, f) belongs to the owner of f) are reported."""
    plan = []
    for r in range(world):
        lo, hi = shard_range(n_frames, r, world)
        first = lo - 1 if (lo > 0 and hi > lo) else lo
        plan.append((first, lo, hi))
    return plan


def process_shard(ctx, load_frames, n_frames, rank, world, batch, do_match=True):
    """Runs the front end over this rank's shard in batches of `batch` frames through the C ABI.

    load_frames(a, b) -> uint8 array [b-a, H, W] (pinned or pageable host memory).
    Returns a dict of per-frame results for frames [lo, hi): n_kp, rows, cols, scores, desc and, for f > 0,
    match_idx / match_dist of the pair (f-1, f)."""
    if batch < 2:
        # each batch re-loads the previous batch's last frame (the seam): a 1-frame batch would never advance
        raise ValueError("process_shard needs batch >= 2 (got %d)" % batch)
    first, lo, hi = shard_plan(n_frames, world)[rank]
    K = ctx.max_kp
    n_own = hi - lo
    res = dict(frame=np.arange(lo, hi), n_kp=np.zeros(n_own, np.int32), rows=np.zeros((n_own, K), np.int32),
               cols=np.zeros((n_own, K), np.int32), scores=np.zeros((n_own, K), np.float32),
               desc=np.zeros((n_own, K, 32), np.uint8), match_idx=np.full((n_own, K), -1, np.int32),
               match_dist=np.full((n_own, K), -1, np.int32))
    if n_own == 0:
        return res
    a = first
    while a < hi:
        # each batch re-loads the last frame of the previous one so the pair across the seam is matched
        b = min(hi, a + batch)
        out = ctx.process_host_batch(np.ascontiguousarray(load_frames(a, b)), do_match)
        for i in range(b - a):
            f = a + i
            if f < lo:
                continue
            j = f - lo
            for k in ("n_kp", "rows", "cols", "scores", "desc"):
                res[k][j] = out[k][i]
            if i > 0:  # row i of a batch holds the match of (f-1, f)
                res["match_idx"][j] = out["match_idx"][i]
                res["match_dist"][j] = out["match_dist"][i]
        if b >= hi:
            break
        a = b - 1
    return res


RESULT_KEYS = ("n_kp", "rows", "cols", "scores", "desc", "match_idx", "match_dist")


def alloc_result_tensors(n_own, max_kp, device):
    """Fixed-shape torch tensors for the results of `n_own` owned frames (what yavo_fetch_batch fills)."""
    import torch
    K = max_kp
    return dict(n_kp=torch.zeros(n_own, dtype=torch.int32, device=device),
                rows=torch.zeros((n_own, K), dtype=torch.int32, device=device),
                cols=torch.zeros((n_own, K), dtype=torch.int32, device=device),
                scores=torch.zeros((n_own, K), dtype=torch.float32, device=device),
                desc=torch.zeros((n_own, K, 32), dtype=torch.uint8, device=device),
                match_idx=torch.full((n_own, K), -1, dtype=torch.int32, device=device),
                match_dist=torch.full((n_own, K), -1, dtype=torch.int32, device=device))


def gather_tensors_to_rank0(res, n_frames, rank, world, group=None):
    """Gathers per-rank result TENSORS (first dimension = the rank's owned frames [lo, hi)) on rank 0 with
    torch.distributed.gather of fixed-shape buffers, padded to the largest shard: NCCL over NVLink for CUDA tensors,
    gloo for CPU tensors.  Nothing is pickled (a seq-00-length run returns ~470 MB).  Returns the concatenated tensors
    on rank 0 (frame order), None elsewhere."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return res
    plan = shard_plan(n_frames, world)
    own = [hi - lo for _, lo, hi in plan]
    max_own = max(own)
    out = {}
    for k in RESULT_KEYS:
        t = res[k]
        assert t.shape[0] == own[rank], (k, t.shape, own[rank])
        pad = t if t.shape[0] == max_own else torch.cat(
            [t, torch.zeros((max_own - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)])
        parts = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad.contiguous(), parts, dst=0, group=group)
        if rank == 0:
            out[k] = torch.cat([parts[r][:own[r]] for r in range(world)])
    return out if rank == 0 else None


def run_resident_shard(ctx, n_frames, rank, world, do_match=True):
    """Front end over this rank's shard, already uploaded to slots [0, hi - first) (slot 0 = frame `first`, the seam frame
    when the shard does not start the sequence).  Returns (first, lo, hi)."""
    first, lo, hi = shard_plan(n_frames, world)[rank]
    if hi > first:
        ctx.frontend_batch(0, hi - first, do_match)
    return first, lo, hi


def fetch_owned(ctx, first, lo, hi, out):
    """Device results of the owned frames [lo, hi) into `out` (alloc_result_tensors, on the context's GPU or on the
    host): the seam frame's own results belong to the previous rank and are skipped; row 0 of the owned block holds the
    match of the seam pair (lo-1, lo)."""
    if hi > lo:
        ctx.fetch_batch_ptrs(lo - first, hi - lo, {k: out[k].data_ptr() for k in RESULT_KEYS})
    return out


def results_digest(res, n_frames):
    """sha256 over the defined part of a full result set (entries beyond n_kp are unspecified): what makes a sharded
    run comparable with a single-GPU run bit for bit."""
    import hashlib
    h = hashlib.sha256()
    g = {k: (v.cpu().numpy() if hasattr(v, "cpu") else np.asarray(v)) for k, v in res.items()}
    h.update(np.ascontiguousarray(g["n_kp"][:n_frames]).tobytes())
    for f in range(n_frames):
        k = int(g["n_kp"][f])
        for key in ("rows", "cols", "scores", "desc"):
            h.update(np.ascontiguousarray(g[key][f, :k]).tobytes())
        if f > 0:
            kq = int(g["n_kp"][f - 1])
            h.update(np.ascontiguousarray(g["match_idx"][f, :kq]).tobytes())
            h.update(np.ascontiguousarray(g["match_dist"][f, :kq]).tobytes())
    return h.hexdigest()


def gather_to_rank0(res, rank, world, group=None):
    """Gathers the per-rank result dicts on rank 0 by pickling them (small runs and tests; a full-length run uses
    gather_tensors_to_rank0)."""
    if world == 1:
        return res
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object(res, parts, dst=0, group=group)
    if rank != 0:
        return None
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in res}
