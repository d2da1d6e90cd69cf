"""GPU: BASELINE.json configs 3-5 at (or near) full size — exact against the oracle where the oracle finishes in
seconds, size-independent properties beyond that."""
import numpy as np
import pytest

from ya_vo_b200 import sharding, synth

pytestmark = pytest.mark.gpu


def hamming_rows(a, b):
    return np.unpackbits(np.bitwise_xor(a, b), axis=1).sum(axis=1)


def test_4k_frame_20k_keypoints(cuda_lib, oracle, offsets):
    """config 4: 3840x2160 uniform noise, cap raised to 20,000 (through the C ABI's max_kp)."""
    img = synth.synth_frame("U", 3, 2160, 3840)
    with cuda_lib.Context(device=0, n_slots=1, max_rows=2160, max_cols=3840, max_kp=20000) as ctx:
        ctx.set_brief_offsets(offsets)
        ctx.upload(0, img)
        assert np.array_equal(ctx.blurred(0), oracle.gaussian_blur(img))
        r, c, s = ctx.fast_candidates(0)
        er, ec, es = oracle.fast_candidates(img)
        assert r.size == er.size > 300000
        assert np.array_equal(r, er) and np.array_equal(c, ec) and np.array_equal(s.view(np.uint32), es.view(np.uint32))
        kr, kc, ks, nc = ctx.fast_detect(0, 20000)
        okr, okc, oks, onc = oracle.fast_detect(img, 20000)
        assert nc == onc and kr.size == 20000
        assert np.array_equal(kr, okr) and np.array_equal(kc, okc) and np.array_equal(ks.view(np.uint32), oks.view(np.uint32))
        d, v, oob = ctx.brief_describe(0, kr, kc)
        od, ov, ooob = oracle.brief(img, offsets, okr, okc)
        assert np.array_equal(v, ov) and np.array_equal(d, od) and oob == ooob


def test_candidate_overflow_is_reported(cuda_lib):
    img = synth.synth_frame("U", 0)
    with cuda_lib.Context(device=0, n_slots=1, max_rows=376, max_cols=1241, max_kp=2000, max_cand=1000) as ctx:
        ctx.upload(0, img)
        with pytest.raises(cuda_lib.YavoError):
            ctx.fast_detect(0)


@pytest.mark.parametrize("n1,n2", [(1024, 65536), (16384, 16384), (65536, 1000)])
def test_matcher_sweep_exact(cuda_lib, oracle, n1, n2):
    """config 5 corners that the oracle still finishes in seconds."""
    d1 = synth.synth_descriptors(n1, n1 * 131 + n2)
    d2 = synth.planted_descriptors(d1, n2, 9)
    with cuda_lib.Context(device=0, n_slots=1, max_rows=64, max_cols=128, max_kp=16) as ctx:
        idx, dist, sec, rev = ctx.match(d1, d2, extensions=True)
    eidx, edist, esec, erev = oracle.match(d1, d2, extensions=True)
    assert np.array_equal(idx, eidx) and np.array_equal(dist, edist)
    assert np.array_equal(sec, esec) and np.array_equal(rev, erev)


def test_matcher_64k_x_64k_properties(cuda_lib):
    """config 5 at 64k x 64k (4.3 G pairs): properties instead of the full oracle."""
    n = 65536
    d1 = synth.synth_descriptors(n, 1)
    rng = np.random.default_rng(2)
    # train set: every query planted once with exactly 6 bits flipped, at a permuted position
    d2 = d1.copy()
    for k in range(6):
        bits = rng.integers(0, 256, n)
        d2[np.arange(n), bits // 8] ^= (1 << (bits % 8)).astype(np.uint8)
    perm = rng.permutation(n)
    d2p = np.empty_like(d2)
    d2p[perm] = d2
    with cuda_lib.Context(device=0, n_slots=1, max_rows=64, max_cols=128, max_kp=16) as ctx:
        idx, dist, sec, rev = ctx.match(d1, d2p, extensions=True)
    assert idx.min() >= 0 and idx.max() < n
    # the reported distance is the distance to the reported index
    assert np.array_equal(dist, hamming_rows(d1, d2p[idx]))
    # the planted twin is at most 6 bits away, random others ~128: the twin must be found
    assert dist.max() <= 6 and np.array_equal(idx, perm)
    assert np.all(sec >= dist) and sec.min() > 60
    # cross-check direction agrees (each train descriptor's best query is its source)
    assert np.array_equal(rev[perm], np.arange(n))
    # minimality, brute force on a sample of queries
    for i in rng.integers(0, n, 24):
        full = hamming_rows(np.broadcast_to(d1[i], d2p.shape), d2p)
        assert full.min() == dist[i] and int(full.argmin()) == idx[i]


def test_sequence_batches_across_seams(cuda_lib, oracle, offsets):
    """config 3 in miniature: a 13-frame sequence through process_shard in batches of 5 (seams re-use one frame),
    as one rank and as two 'ranks' run one after the other; every frame and every consecutive pair once."""
    F = 13
    frames = synth.synth_batch(F, "G30", 1000)
    frames[5] = synth.shifted_pair(frames[4], 5)
    exp = oracle.pipeline(frames, offsets, 2000, True, nthreads=8)
    with cuda_lib.Context(device=0, n_slots=8, max_rows=376, max_cols=1241, max_kp=2000) as ctx:
        ctx.set_brief_offsets(offsets)
        parts = [sharding.process_shard(ctx, lambda a, b: frames[a:b], F, r, 2, batch=5) for r in range(2)]
        one = sharding.process_shard(ctx, lambda a, b: frames[a:b], F, 0, 1, batch=5)
    two = {k: np.concatenate([p[k] for p in parts], axis=0) for k in one}
    for res in (one, two):
        assert np.array_equal(res["frame"], np.arange(F)) and np.array_equal(res["n_kp"], exp["n_kp"])
        for f in range(F):
            k = exp["n_kp"][f]
            assert np.array_equal(res["rows"][f, :k], exp["rows"][f, :k])
            assert np.array_equal(res["desc"][f, :k], exp["desc"][f, :k])
            if f > 0:
                kq = exp["n_kp"][f - 1]
                assert np.array_equal(res["match_idx"][f, :kq], exp["match_idx"][f, :kq])
                assert np.array_equal(res["match_dist"][f, :kq], exp["match_dist"][f, :kq])


def test_pipelined_and_resident_paths_agree(cuda_lib, offsets):
    """pinned host batch (3-stream pipeline) == pageable host batch == device-resident batch."""
    import torch
    frames = synth.synth_batch(40, "G30", 2000)
    pinned = torch.empty(frames.shape, dtype=torch.uint8).pin_memory()
    pinned.numpy()[:] = frames
    with cuda_lib.Context(device=0, n_slots=40, max_rows=376, max_cols=1241, max_kp=2000) as ctx:
        ctx.set_brief_offsets(offsets)
        ctx.set_pipeline_chunk(7)
        a = ctx.process_host_batch(pinned.numpy(), True)
        a = {k: v.copy() for k, v in a.items()}
        b = ctx.process_host_batch(frames, True)
        b = {k: v.copy() for k, v in b.items()}
        ctx.upload_batch(0, frames)
        ctx.set_sub_batch(9)
        ctx.frontend_batch(0, 40, True)
        c = ctx.fetch_batch(0, 40)
    for k in ("n_kp", "rows", "cols", "scores", "desc"):
        assert np.array_equal(a[k], b[k]) and np.array_equal(a[k], c[k]), k
    for f in range(1, 40):
        kq = a["n_kp"][f - 1]
        for k in ("match_idx", "match_dist"):
            assert np.array_equal(a[k][f, :kq], b[k][f, :kq]) and np.array_equal(a[k][f, :kq], c[k][f, :kq]), (k, f)


def test_async_submit_two_batches_in_flight(cuda_lib, offsets):
    """yavo_submit_host_batch / yavo_wait_batch / yavo_wait: consecutive batches overlap and still give the
    results of the synchronous path."""
    import torch
    fa = synth.synth_batch(24, "G30", 3000)
    fb = synth.synth_batch(24, "B4", 4000)
    pa = torch.empty(fa.shape, dtype=torch.uint8).pin_memory()
    pb = torch.empty(fb.shape, dtype=torch.uint8).pin_memory()
    pa.numpy()[:] = fa
    pb.numpy()[:] = fb
    with cuda_lib.Context(device=0, n_slots=24, max_rows=376, max_cols=1241, max_kp=2000) as ctx:
        ctx.set_brief_offsets(offsets)
        ctx.set_pipeline_chunk(5)
        ra = {k: v.copy() for k, v in ctx.process_host_batch(pa.numpy(), True).items()}
        rb = {k: v.copy() for k, v in ctx.process_host_batch(pb.numpy(), True).items()}
        for rounds in range(2):
            oa, ta = ctx.submit_host_batch(pa.numpy(), True)
            ob, tb = ctx.submit_host_batch(pb.numpy(), True)
            oc, tc = ctx.submit_host_batch(pa.numpy(), True)
            ctx.wait_batch(ta)  # batch A is complete while B and C are still in flight
            assert np.array_equal(oa["n_kp"], ra["n_kp"])
            assert all(np.array_equal(oa["desc"][f, :ra["n_kp"][f]], ra["desc"][f, :ra["n_kp"][f]]) for f in range(24))
            ctx.wait()
            for got, exp in ((oa, ra), (ob, rb), (oc, ra)):
                assert np.array_equal(got["n_kp"], exp["n_kp"])
                for f in range(24):
                    k = exp["n_kp"][f]  # entries beyond n_kp are unspecified (stale slot contents)
                    for key in ("rows", "cols", "scores", "desc"):
                        assert np.array_equal(got[key][f, :k], exp[key][f, :k]), (key, f)
                    if f > 0:
                        kq = exp["n_kp"][f - 1]
                        assert np.array_equal(got["match_idx"][f, :kq], exp["match_idx"][f, :kq])
                        assert np.array_equal(got["match_dist"][f, :kq], exp["match_dist"][f, :kq])
        with pytest.raises(cuda_lib.YavoError):
            ctx.submit_host_batch(fa, True)  # pageable memory is refused by the asynchronous entry point


def test_seq00_length_batch(cuda_lib, oracle, offsets):
    """config 3 at full length: 4541 frames of 1241x376 resident on one GPU, one set of launches; frames spread over
    the batch are compared with the oracle bit for bit, the rest through batch-wide invariants."""
    F = 4541
    rng = np.random.default_rng(99)
    # uniform noise is the cheapest generator (and the densest candidate lists: ~17 k per frame)
    frames = rng.integers(0, 256, (F, 376, 1241), dtype=np.uint8)
    picks = [1, 777, 2270, 4540]
    with cuda_lib.Context(device=0, n_slots=F, max_rows=376, max_cols=1241, max_kp=2000, max_cand=32768) as ctx:
        ctx.set_brief_offsets(offsets)
        ctx.upload_batch(0, frames)
        ctx.frontend_batch(0, F, True)
        out = ctx.fetch_batch(0, F)
    assert out["n_kp"].min() > 1800 and out["n_kp"].max() <= 2000
    for f in picks:
        exp = oracle.pipeline(frames[f - 1:f + 1], offsets, 2000, True, nthreads=2)
        k, kq = exp["n_kp"][1], exp["n_kp"][0]
        assert out["n_kp"][f] == k and out["n_kp"][f - 1] == kq
        assert np.array_equal(out["rows"][f, :k], exp["rows"][1, :k]) and np.array_equal(out["cols"][f, :k], exp["cols"][1, :k])
        assert np.array_equal(out["scores"][f, :k].view(np.uint32), exp["scores"][1, :k].view(np.uint32))
        assert np.array_equal(out["desc"][f, :k], exp["desc"][1, :k])
        assert np.array_equal(out["match_idx"][f, :kq], exp["match_idx"][1, :kq])
        assert np.array_equal(out["match_dist"][f, :kq], exp["match_dist"][1, :kq])
    # invariants over every frame: scores non-increasing, match distances consistent with the descriptors
    for f in range(1, F, 97):
        k, kq = out["n_kp"][f], out["n_kp"][f - 1]
        assert np.all(np.diff(out["scores"][f, :k]) <= 0)
        idx, dist = out["match_idx"][f, :kq], out["match_dist"][f, :kq]
        assert idx.min() >= 0 and idx.max() < k
        assert np.array_equal(dist, hamming_rows(out["desc"][f - 1, :kq], out["desc"][f, idx]))


def test_pinned_arrays_through_the_c_abi(cuda_lib, offsets):
    """yavo_pinned_alloc / yavo_pinned_free: page-locked frame and result arrays for callers without a CUDA runtime
    binding; the asynchronous submit / wait pair on them gives the same results as the synchronous call."""
    frames = np.stack([synth.synth_frame("G30", 1000 + f, 120, 320) for f in range(6)])
    buf = cuda_lib.pinned_zeros(frames.shape, np.uint8)
    buf[:] = frames
    assert buf.flags["C_CONTIGUOUS"] and buf.dtype == np.uint8
    with cuda_lib.Context(device=0, n_slots=6, max_rows=120, max_cols=320, max_kp=500) as ctx:
        ctx.set_brief_offsets(offsets)
        ref = ctx.process_host_batch(frames, True)
        out = ctx.alloc_batch_outputs(6, pinned=True)
        _, ticket = ctx.submit_host_batch(buf, True, out)
        ctx.wait_batch(ticket)
        ctx.wait()
    for k in ("n_kp", "rows", "cols", "desc", "match_idx", "match_dist"):
        assert np.array_equal(out[k], ref[k]), k
    del buf, out  # the blocks are freed when the last view goes away


def test_set_matcher_rejects_unknown_kinds(cuda_lib):
    with cuda_lib.Context(device=0, n_slots=1, max_rows=64, max_cols=128, max_kp=16) as ctx:
        for kind in (0, 1, 2, "tc", "tc8", "popc"):
            ctx.set_matcher(kind)
        with pytest.raises(cuda_lib.YavoError):
            ctx.set_matcher(3)
        ctx.set_matcher("tc")
