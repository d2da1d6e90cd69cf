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


def gather_to_rank0(res, rank, world, group=None):
    """Gathers the per-rank result dicts on rank 0 (no collective runs on the kernels' data path)."""
    if world == 1:
        return res
    import torch.distributed as dist
    parts = [None] * world if rank == 0 else None
    dist.gather_object(res, parts, dst=0, group=group)
    if rank != 0:
        return None
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in res}
