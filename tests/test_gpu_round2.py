"""Round-2 regression and feature tests on the GPU (through the C ABI, checked against the oracle)."""
import numpy as np
import pytest

from ya_vo_b200 import synth

pytestmark = pytest.mark.gpu


def _pin(a):
    from ya_vo_b200 import capi
    p = capi.pinned_zeros(a.shape, a.dtype)
    p[...] = a
    return p


def _same_results(got, exp, n):
    assert np.array_equal(got["n_kp"][:n], exp["n_kp"][:n])
    for f in range(n):
        k = exp["n_kp"][f]
        for key in ("rows", "cols", "scores", "desc"):
            assert np.array_equal(got[key][f, :k], exp[key][f, :k]), (key, f)
        if f > 0:
            kq = exp["n_kp"][f - 1]
            assert np.array_equal(got["match_idx"][f, :kq], exp["match_idx"][f, :kq]), f
            assert np.array_equal(got["match_dist"][f, :kq], exp["match_dist"][f, :kq]), f


def test_alternating_batch_sizes_without_waiting(cuda_lib, offsets):
    """Stage size changes between un-waited submits (64 frames -> 16-frame stages, 512 -> 128): the staging halves move,
    and a new half may cover both halves of the previous submit (round-1 advisor finding)."""
    small = _pin(synth.synth_batch(64, "G30", 100, 120, 320))
    big = _pin(synth.synth_batch(512, "B4", 900, 120, 320))
    with cuda_lib.Context(device=0, n_slots=512, max_rows=120, max_cols=320, max_kp=300) as ctx:
        ctx.set_brief_offsets(offsets)
        es = {k: v.copy() for k, v in ctx.process_host_batch(small, True).items()}
        eb = {k: v.copy() for k, v in ctx.process_host_batch(big, True).items()}
        for _ in range(3):
            o1, t1 = ctx.submit_host_batch(small, True)
            o2, t2 = ctx.submit_host_batch(big, True)
            o3, t3 = ctx.submit_host_batch(small, True)
            ctx.wait_batch(t1)
            _same_results(o1, es, 64)
            ctx.wait_batch(t2)
            _same_results(o2, eb, 512)
            ctx.wait()
            _same_results(o3, es, 64)


def test_wait_batch_reports_the_overflow_of_its_own_batch(cuda_lib, offsets):
    """A candidate-list overflow is reported by yavo_wait_batch for the batch that had it, not for its neighbours."""
    ok = _pin(synth.synth_batch(8, "G30", 5, 120, 320))
    bad = _pin(synth.synth_batch(8, "U", 6, 120, 320))  # uniform noise: ~4x the candidates
    with cuda_lib.Context(device=0, n_slots=8, max_rows=120, max_cols=320, max_kp=300, max_cand=900) as ctx:
        ctx.set_brief_offsets(offsets)
        exp = {k: v.copy() for k, v in ctx.process_host_batch(ok, True).items()}
        _, t1 = ctx.submit_host_batch(ok, True)
        _, t2 = ctx.submit_host_batch(bad, True)
        o3, t3 = ctx.submit_host_batch(ok, True)
        ctx.wait_batch(t1)
        with pytest.raises(cuda_lib.YavoError):
            ctx.wait_batch(t2)
        ctx.wait_batch(t3)
        _same_results(o3, exp, 8)
        ctx.wait()  # nothing left to report
        _, t4 = ctx.submit_host_batch(bad, True)
        with pytest.raises(cuda_lib.YavoError):
            ctx.wait()  # a batch nobody waited for by ticket is reported here


def test_filter_pairs_empty_train_set(cuda_lib, oracle, offsets):
    """Train frame without keypoints: distances INT_MAX, nothing kept by the device filter or the host one."""
    a = synth.synth_frame("G30", 1, 120, 320)
    frames = np.stack([a, np.full_like(a, 128)])
    with cuda_lib.Context(device=0, n_slots=2, max_rows=120, max_cols=320, max_kp=300) as ctx:
        ctx.set_brief_offsets(offsets)
        out = ctx.process_host_batch(frames, True)
        assert out["n_kp"][1] == 0 and out["n_kp"][0] > 0
        k = out["n_kp"][0]
        assert (out["match_dist"][1, :k] == 2**31 - 1).all() and (out["match_idx"][1, :k] == -1).all()
        n_pairs, min_dist, _ = ctx.filter_pairs(0, 2, 20)
        assert n_pairs[1] == 0
        assert not cuda_lib.remove_outliers(out["match_dist"][1, :k], 20).any()
        assert not oracle.remove_outliers(out["match_dist"][1, :k], 20).any()


def test_two_shards_on_real_contexts_equal_one(cuda_lib, offsets):
    """The seq-00 sharding (bench.py --config seq00) with real contexts: two shards with the seam frame, results fetched
    into torch CUDA tensors (device pointers through the C ABI) — concatenated they equal the single-context run bit for
    bit, the seam pair included."""
    import torch
    from ya_vo_b200 import sharding
    F, world = 23, 2
    frames = synth.synth_batch(F, "G30", 1000, 120, 320)
    frames[12] = synth.shifted_pair(frames[11], 5)  # true matches across the seam (shard 1 starts at frame 11)
    dev = torch.device("cuda", 0)
    parts = []
    for rank in range(world):
        first, lo, hi = sharding.shard_plan(F, world)[rank]
        with cuda_lib.Context(device=0, n_slots=hi - first, max_rows=120, max_cols=320, max_kp=300) as ctx:
            ctx.set_brief_offsets(offsets)
            ctx.upload_batch(0, np.ascontiguousarray(frames[first:hi]))
            assert sharding.run_resident_shard(ctx, F, rank, world) == (first, lo, hi)
            res = sharding.fetch_owned(ctx, first, lo, hi, sharding.alloc_result_tensors(hi - lo, 300, dev))
            parts.append({k: v.cpu() for k, v in res.items()})
    full = {k: torch.cat([p[k] for p in parts]) for k in sharding.RESULT_KEYS}
    with cuda_lib.Context(device=0, n_slots=F, max_rows=120, max_cols=320, max_kp=300) as ctx:
        ctx.set_brief_offsets(offsets)
        ctx.upload_batch(0, frames)
        ctx.frontend_batch(0, F, True)
        one = ctx.fetch_batch(0, F)
    assert sharding.results_digest(full, F) == sharding.results_digest(one, F)
    kq = int(one["n_kp"][11])
    assert kq > 0 and np.array_equal(full["match_idx"][12, :kq].numpy(), one["match_idx"][12, :kq])


def test_overlapped_feature_pipeline_gives_identical_results(cuda_lib, offsets):
    """yavo_set_overlap: chunks rotating over internal streams change the schedule, not the results."""
    frames = synth.synth_batch(48, "B4", 77, 120, 320)
    outs = []
    for chunk, streams in ((0, 1), (8, 3), (16, 2), (5, 4)):
        with cuda_lib.Context(device=0, n_slots=48, max_rows=120, max_cols=320, max_kp=300) as ctx:
            ctx.set_brief_offsets(offsets)
            ctx.set_overlap(chunk, streams)
            ctx.upload_batch(0, frames)
            for _ in range(2):
                ctx.frontend_batch(0, 48, True)
            outs.append(ctx.fetch_batch(0, 48))
    for o in outs[1:]:
        _same_results(o, outs[0], 48)


@pytest.mark.parametrize("n1,n2", [(1, 1), (5, 1), (7, 3), (128, 224), (129, 225), (128, 448), (500, 449), (2000, 2000), (1950, 1949),
                                   (333, 4097), (4096, 129), (300, 0), (9000, 7000)])
def test_second_best_from_the_tensor_core_epilogue(cuda_lib, oracle, n1, n2):
    """Ratio-test extension (north_star; no reference counterpart, parity pinned to the oracle's definition only): the
    second smallest distance of every query comes out of the tcgen05 matcher's own epilogue (two smallest keys per
    row), the cross-check index from the same kernel with the roles swapped.  Duplicated train descriptors make
    second == best; single-column train sets leave it at INT_MAX."""
    d1 = synth.synth_descriptors(n1, n1 * 131 + n2)
    d2 = synth.planted_descriptors(d1, n2, 7) if (n1 and n2) else synth.synth_descriptors(n2, 1)
    if n2 > 300:
        d2[290] = d2[10]   # duplicates in different tiles
        d2[40] = d2[10]    # ... and inside one tile
    if n1 > 5 and n2 > 5:
        d1[3] = d2[5]
        d1[4] = ~d2[5]
    with cuda_lib.Context(device=0, n_slots=1, max_rows=64, max_cols=128, max_kp=16) as ctx:
        ctx.set_matcher("tc")
        idx, dist, sec, rev = ctx.match(d1, d2, extensions=True)
        kl = ctx.kernel_launches
    eidx, edist, esec, erev = oracle.match(d1, d2, extensions=True)
    assert np.array_equal(idx, eidx) and np.array_equal(dist, edist)
    assert np.array_equal(sec, esec)
    assert np.array_equal(rev, erev)
    assert kl <= 2  # forward (with second) + reverse: no integer-pipe fallback, no reduce kernels


@pytest.mark.parametrize("kind,seed,H,W,K,min_n", [("U", 3, 376, 1241, 2000, 2048), ("U", 4, 376, 1241, 2000, 8192),
                                                    ("B4", 5, 376, 1241, 2000, 1024), ("G30", 6, 376, 1241, 500, 512),
                                                    ("U", 7, 200, 333, 3000, 700)])
def test_cluster_prepartition_of_large_lists_is_exact(cuda_lib, oracle, kind, seed, H, W, K, min_n):
    """select_big_kernel: the top of the std::sort replay on a thread-block cluster (8 CTAs per frame), forced onto
    moderate frames by lowering the candidate threshold; order, ties and scores must still be the oracle's, for single
    frames and inside a batch."""
    frames = synth.synth_batch(3, kind, seed, H, W)
    with cuda_lib.Context(device=0, n_slots=3, max_rows=H, max_cols=W, max_kp=K) as ctx:
        ctx.set_big_select(min_n)
        for f in range(3):
            ctx.upload(f, frames[f])
            r, c, s, nc = ctx.fast_detect(f, K)
            er, ec, es, enc = oracle.fast_detect(frames[f], K)
            assert nc == enc and nc > min_n
            assert np.array_equal(r, er) and np.array_equal(c, ec) and np.array_equal(s.view(np.uint32), es.view(np.uint32))
        ctx.set_brief_offsets(synth.brief_offsets())
        ctx.upload_batch(0, frames)
        ctx.frontend_batch(0, 3, True)
        on = ctx.fetch_batch(0, 3)
        ctx.set_big_select(0)
        ctx.frontend_batch(0, 3, True)
        off = ctx.fetch_batch(0, 3)
    _same_results(on, off, 3)


@pytest.mark.parametrize("H,W,K", [(376, 1241, 2000), (120, 320, 300), (17, 33, 50), (9, 9, 10), (8, 40, 10), (200, 333, 3000)])
def test_frame_features_matches_oracle(cuda_lib, oracle, offsets, H, W, K):
    """yavo_frame_features (the reference's per-frame sequence as one CUDA graph): detector output, admitted list, ids and
    descriptors against the oracle; repeated calls replay the captured graph, alternating slots and frames."""
    frames = synth.synth_batch(3, "G30" if H > 50 else "U", 21, H, W)
    with cuda_lib.Context(device=0, n_slots=2, max_rows=H, max_cols=W, max_kp=K) as ctx:
        ctx.set_brief_offsets(offsets)
        for rep in range(3):
            for i, img in enumerate(frames):
                ff = ctx.frame_features((i + rep) % 2, img, K)
                er, ec, es, enc = oracle.fast_detect(img, K)
                assert ff["n_cand"] == enc
                assert np.array_equal(ff["rows"], er) and np.array_equal(ff["cols"], ec)
                assert np.array_equal(ff["scores"].view(np.uint32), es.view(np.uint32))
                d, v, _ = oracle.brief(img, offsets, er, ec)
                assert np.array_equal(ff["d_ids"], np.nonzero(v)[0])
                assert np.array_equal(ff["d_rows"], er[v]) and np.array_equal(ff["d_cols"], ec[v])
                assert np.array_equal(ff["desc"], d[v])


def test_frame_features_strided_rows_profiling_and_slot_reuse(cuda_lib, oracle, offsets):
    """Row pitch larger than the width, per-kernel profiling on (plain launches instead of the graph), a smaller cap on
    the same slot, and the other entry points on a slot that a graph filled."""
    H, W = 96, 200
    wide = synth.synth_frame("G30", 31, H, W + 56)
    img = wide[:, :W]  # a view: stride 256, width 200
    dense = np.ascontiguousarray(img)
    with cuda_lib.Context(device=0, n_slots=2, max_rows=H, max_cols=W, max_kp=400) as ctx:
        ctx.set_brief_offsets(offsets)
        L = cuda_lib.lib()
        import ctypes as C
        K = 400
        r, c = np.empty(K, np.int32), np.empty(K, np.int32)
        dr, dc, di = np.empty(K, np.int32), np.empty(K, np.int32), np.empty(K, np.int32)
        desc = np.empty((K, 32), np.uint8)
        n, nd, nc = C.c_int32(), C.c_int32(), C.c_int()
        ptr = C.c_void_p(img.ctypes.data)
        for prof in (False, True, False):
            ctx.set_profiling(prof)
            rc = L.yavo_frame_features(ctx._h, 0, ptr, H, W, img.strides[0], K, C.byref(n), C.c_void_p(r.ctypes.data),
                                       C.c_void_p(c.ctypes.data), None, C.byref(nd), C.c_void_p(dr.ctypes.data),
                                       C.c_void_p(dc.ctypes.data), C.c_void_p(di.ctypes.data), C.c_void_p(desc.ctypes.data), C.byref(nc))
            assert rc == 0
            er, ec, _, enc = oracle.fast_detect(dense, K)
            d, v, _ = oracle.brief(dense, offsets, er, ec)
            assert nc.value == enc and n.value == len(er) and nd.value == int(v.sum())
            assert np.array_equal(r[:n.value], er) and np.array_equal(desc[:nd.value], d[v])
        ctx.set_profiling(False)
        ff = ctx.frame_features(0, dense, 100)  # smaller cap: another graph for the same slot
        er, ec, _, _ = oracle.fast_detect(dense, 100)
        assert np.array_equal(ff["rows"], er)
        ctx._shape[0] = dense.shape
        assert np.array_equal(ctx.download(0), dense)          # the slot holds the frame
        assert np.array_equal(ctx.blurred(0), oracle.gaussian_blur(dense))
        assert L.yavo_slot_holds(ctx._h, 0, C.c_void_p(dense.ctypes.data), H, W, W) == 1
        other = dense.copy()
        other[5, 7] ^= 1
        assert L.yavo_slot_holds(ctx._h, 0, C.c_void_p(other.ctypes.data), H, W, W) == 0
        assert L.yavo_slot_holds(ctx._h, 1, C.c_void_p(dense.ctypes.data), H, W, W) == 0


DENSE_PATTERN = np.array([[0, 60, 0, 60], [0, 120, 255, 180], [60, 120, 60, 120], [180, 60, 60, 255]], np.uint8)


def test_very_dense_corners_take_several_list_rounds(cuda_lib, oracle, offsets):
    """A periodic texture on which every second pixel passes the segment test: 2048 corners per full 128x32 tile, twice
    the detect kernel's per-round corner list (K1_LIST = 1024), 17 k candidates on a 120x320 frame, all of them tied in a
    handful of Harris responses (the replay's worst case).  Candidates in scan order, responses, the ordered top-K and
    the descriptors must still be the oracle's; the default candidate capacity (a quarter of the pixels) reports the
    overflow instead."""
    H, W = 120, 320
    dense = np.tile(DENSE_PATTERN, (H // 4, W // 4))
    mixed = dense.copy()
    mixed[:, W // 2:] = synth.synth_frame("G30", 9, H, W)[:, W // 2:]
    frames = np.stack([dense, mixed])
    with cuda_lib.Context(device=0, n_slots=2, max_rows=H, max_cols=W, max_kp=2000, max_cand=20000) as ctx:
        ctx.set_brief_offsets(offsets)
        for f in range(2):
            ctx.upload(f, frames[f])
            r, c, s = ctx.fast_candidates(f)
            er, ec, es = oracle.fast_candidates(frames[f])
            assert er.size > (8000 if f else 17000)
            assert np.array_equal(r, er) and np.array_equal(c, ec) and np.array_equal(s.view(np.uint32), es.view(np.uint32))
            kr, kc, ks, nc = ctx.fast_detect(f, 2000)
            okr, okc, oks, onc = oracle.fast_detect(frames[f], 2000)
            assert nc == onc and np.array_equal(kr, okr) and np.array_equal(kc, okc)
        out = ctx.process_host_batch(frames, True)
        exp = oracle.pipeline(frames, offsets, 2000, True, nthreads=2)
        _same_results(out, exp, 2)
    with cuda_lib.Context(device=0, n_slots=1, max_rows=H, max_cols=W, max_kp=2000) as ctx:  # default capacity: pixels / 4
        ctx.upload(0, dense)
        with pytest.raises(cuda_lib.YavoError):
            ctx.fast_detect(0)


_UMMA_SNIPPET = r"""
import sys, numpy as np
sys.path.insert(0, %(root)r)
from oracle import pyoracle as po
from ya_vo_b200 import capi, synth
off = synth.brief_offsets()
bad = 0
for kind, seed, H, W in (("G30", 5, 376, 1241), ("U", 6, 97, 211), ("B4", 7, 33, 130), ("G30", 8, 200, 1280)):
    frames = synth.synth_batch(3, kind, seed, H, W)
    if kind == "B4":
        frames[1, 5:30, 20:100] = 255  # the largest 16-bit row sums
    with capi.Context(device=0, n_slots=3, max_rows=H, max_cols=W, max_kp=700) as ctx:
        ctx.set_brief_offsets(off)
        ctx.upload_batch(0, frames)
        blur = [ctx.blurred(f) for f in range(3)]
        out = ctx.process_host_batch(frames, True)
    exp = po.pipeline(frames, off, 700, True, nthreads=2)
    for f in range(3):
        bad += int(np.count_nonzero(blur[f] != po.gaussian_blur(frames[f])))
        k = exp["n_kp"][f]
        bad += int(out["n_kp"][f] != k) + int(np.count_nonzero(out["desc"][f, :k] != exp["desc"][f, :k]))
        bad += int(np.count_nonzero(out["scores"][f, :k].view(np.uint32) != exp["scores"][f, :k].view(np.uint32)))
print("UMMA_BLUR_MISMATCHES", bad)
"""


@pytest.mark.parametrize("variant", ["umma", "hyb"])
def test_blur_on_the_tensor_cores_variant_is_bit_exact(cuda_lib, variant):
    """build/libyavo_umma.so (-DYAVO_BLUR_UMMA=1: the 9x9 Gaussian as two banded tcgen05.mma kind::i8 products) and
    build/libyavo_hyb.so (-DYAVO_BLUR_UMMA=2: horizontal pass on the tensor cores, the operand arriving by a 4-D tensor
    copy, vertical pass from the accumulator registers), both built by __graft_entry__.build() from blur_umma.cuh.  Not
    the default (measured slower), but kept exact: blurred planes, descriptors and score bits against the oracle on
    interior, edge, tiny and saturated frames."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "build", "libyavo_%s.so" % variant)
    if not os.path.exists(lib):
        pytest.skip("variant library not built")
    env = dict(os.environ, YAVO_LIB_PATH=lib)
    r = subprocess.run([sys.executable, "-c", _UMMA_SNIPPET % {"root": root}], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "UMMA_BLUR_MISMATCHES 0" in r.stdout, r.stdout[-500:]


def test_frame_features_frames_of_different_widths_on_one_slot(cuda_lib, oracle, offsets):
    """The single-frame path stages a frame at the device pitch and copies it straight into the slot: after a wide frame a
    narrower one must not see the wide frame's pixels in its row padding (blur reflects at the new width, the detector's
    staged tiles read beyond it), and the slot must download as the narrow frame."""
    wide = synth.synth_frame("G30", 41, 120, 500)
    narrow = synth.synth_frame("G30", 42, 90, 333)
    with cuda_lib.Context(device=0, n_slots=1, max_rows=120, max_cols=500, max_kp=600) as ctx:
        ctx.set_brief_offsets(offsets)
        for img in (wide, narrow, wide, narrow):
            ff = ctx.frame_features(0, img, 600)
            er, ec, es, enc = oracle.fast_detect(img, 600)
            d, v, _ = oracle.brief(img, offsets, er, ec)
            assert ff["n_cand"] == enc and np.array_equal(ff["rows"], er) and np.array_equal(ff["cols"], ec)
            assert np.array_equal(ff["scores"].view(np.uint32), es.view(np.uint32))
            assert np.array_equal(ff["desc"], d[v])
            ctx._shape[0] = img.shape
            assert np.array_equal(ctx.download(0), img)
            assert np.array_equal(ctx.blurred(0), oracle.gaussian_blur(img))
