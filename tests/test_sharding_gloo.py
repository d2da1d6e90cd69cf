"""CPU, world_size 2 over gloo: the multi-GPU host logic (frame ranges with the one-frame overlap, batching
across seams, gather to rank 0).  The kernels are replaced by a stand-in context whose 'descriptor' of a frame is
a function of its pixels and whose 'match' of (f-1, f) is a function of both, so coverage is checkable."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from ya_vo_b200 import sharding


class FakeCtx:
    max_kp = 4

    def process_host_batch(self, frames, do_match=True):
        n = frames.shape[0]
        K = self.max_kp
        sig = frames.reshape(n, -1).astype(np.int64).sum(axis=1)
        out = dict(n_kp=(sig % 5).astype(np.int32), rows=np.zeros((n, K), np.int32), cols=np.zeros((n, K), np.int32),
                   scores=np.zeros((n, K), np.float32), desc=np.zeros((n, K, 32), np.uint8),
                   match_idx=np.full((n, K), -1, np.int32), match_dist=np.full((n, K), -1, np.int32))
        out["rows"][:, 0] = sig
        for i in range(1, n):
            out["match_dist"][i, 0] = (sig[i - 1] * 31 + sig[i]) % 100003
        return out


def make_frames(n):
    rng = np.random.default_rng(3)
    return rng.integers(0, 256, (n, 6, 7), dtype=np.uint8)


def expected(frames):
    sig = frames.reshape(frames.shape[0], -1).astype(np.int64).sum(axis=1)
    md = np.full(frames.shape[0], -1, np.int64)
    md[1:] = (sig[:-1] * 31 + sig[1:]) % 100003
    return sig, md


def test_plan_covers_every_frame_and_pair():
    for n in (1, 2, 7, 64, 4541):
        for world in (1, 2, 3, 4, 8):
            plan = sharding.shard_plan(n, world)
            owned = []
            for first, lo, hi in plan:
                assert first in (lo, lo - 1) and first >= 0
                if hi > lo and lo > 0:
                    assert first == lo - 1  # the pair (lo-1, lo) is local
                owned += list(range(lo, hi))
            assert owned == list(range(n))


def test_single_rank_batches_across_seams():
    frames = make_frames(23)
    res = sharding.process_shard(FakeCtx(), lambda a, b: frames[a:b], 23, 0, 1, batch=5)
    sig, md = expected(frames)
    assert np.array_equal(res["rows"][:, 0], sig)
    assert np.array_equal(res["match_dist"][:, 0], md)


def _worker(rank, world, port, n_frames, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = make_frames(n_frames)
    res = sharding.process_shard(FakeCtx(), lambda a, b: frames[a:b], n_frames, rank, world, batch)
    full = sharding.gather_to_rank0(res, rank, world)
    if rank == 0:
        sig, md = expected(frames)
        ok = (np.array_equal(full["frame"], np.arange(n_frames)) and np.array_equal(full["rows"][:, 0], sig)
              and np.array_equal(full["match_dist"][:, 0], md))
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames,batch", [(17, 4), (64, 64), (3, 8)])
def test_two_ranks_gloo(n_frames, batch):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


class FakeResidentCtx(FakeCtx):
    """Stand-in for the device-resident path: frontend_batch over uploaded slots, fetch_batch_ptrs into torch tensors."""

    def __init__(self, frames):
        self.frames = frames

    def frontend_batch(self, slot0, n, do_match=True):
        self.out = self.process_host_batch(self.frames[slot0:slot0 + n], do_match)

    def fetch_batch_ptrs(self, slot0, n, ptrs):
        import ctypes as C
        for k, p in ptrs.items():
            src = np.ascontiguousarray(self.out[k][slot0:slot0 + n])
            C.memmove(int(p), src.ctypes.data, src.nbytes)


def _tensor_worker(rank, world, port, n_frames, q):
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = make_frames(n_frames)
    first, lo, hi = sharding.shard_plan(n_frames, world)[rank]
    ctx = FakeResidentCtx(frames[first:hi])
    assert sharding.run_resident_shard(ctx, n_frames, rank, world) == (first, lo, hi)
    res = sharding.fetch_owned(ctx, first, lo, hi, sharding.alloc_result_tensors(hi - lo, ctx.max_kp, torch.device("cpu")))
    full = sharding.gather_tensors_to_rank0(res, n_frames, rank, world)
    if rank == 0:
        sig, md = expected(frames)
        ok = np.array_equal(full["rows"][:, 0].numpy(), sig) and np.array_equal(full["match_dist"][:, 0].numpy(), md)
        # the digest of the gathered result equals the digest of a single-rank run
        one = FakeResidentCtx(frames)
        one.frontend_batch(0, n_frames)
        ref = sharding.fetch_owned(one, 0, 0, n_frames, sharding.alloc_result_tensors(n_frames, one.max_kp, torch.device("cpu")))
        ok = ok and sharding.results_digest(full, n_frames) == sharding.results_digest(ref, n_frames)
        q.put(bool(ok))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [17, 64, 3])
def test_two_ranks_tensor_gather_gloo(n_frames):
    """The seq-00 path of bench.py (--config seq00): resident shard incl. the seam frame, results fetched into fixed-shape
    tensors and gathered with dist.gather (no pickling); rank 0's result is bit-identical to a single-rank run."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_tensor_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
