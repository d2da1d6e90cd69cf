"""The streaming caller: decode-ahead FrameStream on the CPU; the whole loop against the oracle on the GPU."""
import os

import numpy as np
import pytest

from ya_vo_b200 import stream, synth


def test_frame_stream_batches_and_seams():
    frames = synth.synth_batch(11, "U", 1, 12, 20)
    s = stream.FrameStream(11, lambda i: frames[i], batch=4, shape=(12, 20), prefetch=2)
    got = []
    for a, b, buf, bi in s:
        assert np.array_equal(buf, frames[a:b])
        got.append((a, b))
        s.release(bi)
    assert got == [(0, 4), (3, 7), (6, 10), (9, 11)]


def test_frame_stream_reports_bad_frames():
    frames = synth.synth_batch(3, "U", 1, 12, 20)
    s = stream.FrameStream(3, lambda i: frames[i][:, :5] if i == 2 else frames[i], batch=2, shape=(12, 20))
    with pytest.raises(ValueError):
        for a, b, buf, bi in s:
            s.release(bi)


def test_from_directory_sorts_and_reads(tmp_path):
    import cv2
    frames = synth.synth_batch(5, "G30", 7, 40, 64)
    for i in (3, 0, 4, 1, 2):
        cv2.imwrite(str(tmp_path / ("%06d.png" % i)), frames[i])
    s = stream.FrameStream.from_directory(str(tmp_path), batch=3, pinned=False)
    seen = {}
    for a, b, buf, bi in s:
        for k in range(b - a):
            seen[a + k] = buf[k].copy()
        s.release(bi)
    assert sorted(seen) == [0, 1, 2, 3, 4] and all(np.array_equal(seen[i], frames[i]) for i in range(5))


@pytest.mark.gpu
def test_run_sequence_matches_oracle(cuda_lib, oracle, offsets, tmp_path):
    """PNG files on disk -> decode-ahead -> pinned batches -> submit/wait pipeline -> per-frame callbacks."""
    import cv2
    F = 14
    frames = synth.synth_batch(F, "G30", 1000)
    frames[6] = synth.shifted_pair(frames[5], 3)
    for i in range(F):
        cv2.imwrite(str(tmp_path / ("%06d.png" % i)), frames[i])
    exp = oracle.pipeline(frames, offsets, 2000, True, nthreads=8)
    got = {}

    def on_frame(f, n_kp, rows, cols, scores, desc, mi, md):
        got[f] = (n_kp, rows.copy(), cols.copy(), desc.copy(), None if mi is None else mi.copy(),
                  None if md is None else md.copy())

    with cuda_lib.Context(device=0, n_slots=5, max_rows=376, max_cols=1241, max_kp=2000) as ctx:
        ctx.set_brief_offsets(offsets)
        s = stream.FrameStream.from_directory(str(tmp_path), batch=5, prefetch=2, pinned=True)
        n = stream.run_sequence(ctx, s, True, on_frame)
    assert n == F and sorted(got) == list(range(F))
    for f in range(F):
        k = exp["n_kp"][f]
        assert got[f][0] == k
        assert np.array_equal(got[f][1], exp["rows"][f, :k]) and np.array_equal(got[f][3], exp["desc"][f, :k])
        if f > 0:
            kq = exp["n_kp"][f - 1]
            assert np.array_equal(got[f][4], exp["match_idx"][f, :kq]) and np.array_equal(got[f][5], exp["match_dist"][f, :kq])
        else:
            assert got[f][4] is None


@pytest.mark.gpu
def test_run_sequence_with_tracking_matches_oracle(cuda_lib, oracle, offsets):
    """The steady-state loop of the reference (FAST + BRIEF on the new frame, then calcOpticalFlowPyrLK from the last one)
    in one pipeline: tracks delivered per frame equal the oracle's, across stage and batch seams."""
    F, H, W = 11, 200, 320
    frames = np.empty((F, H, W), np.uint8)
    frames[0] = synth.synth_frame("B4", 60, H, W)
    for f in range(1, F):
        frames[f] = synth.shifted_pair(frames[f - 1], 60 + f)
    tracks = {}

    def on_tracks(f, xy, st, er):
        tracks[f] = (xy.copy(), st.copy(), er.copy())

    import torch
    pinned = [torch.empty((4, H, W), dtype=torch.uint8).pin_memory() for _ in range(4)]
    with cuda_lib.Context(device=0, n_slots=4, max_rows=H, max_cols=W, max_kp=2000) as ctx:
        ctx.set_brief_offsets(offsets)
        ctx.set_pipeline_chunk(3)  # several copy/compute stages per batch: pairs across stage seams
        s = stream.FrameStream(F, lambda i: frames[i], batch=4, shape=(H, W), prefetch=2, buffers=[t.numpy() for t in pinned])
        n = stream.run_sequence(ctx, s, False, None, on_tracks)
    assert n == F and sorted(tracks) == list(range(1, F))
    for f in range(1, F):
        r, c, sc, nc = oracle.fast_detect(frames[f - 1], 2000)
        pts = np.stack([c, r], 1).astype(np.float32)
        o_next, o_st, o_err = oracle.klt_track(frames[f - 1], frames[f], pts)
        k = r.size
        xy, st, er = tracks[f]
        assert np.array_equal(st[:k], o_st), f
        assert np.array_equal(xy[:k].view(np.uint32), o_next.view(np.uint32)), f
        ok = o_st == 1
        assert np.array_equal(er[:k][ok].view(np.uint32), o_err[ok].view(np.uint32)), f


def test_batch_of_one_is_rejected():
    """A 1-frame batch would re-deliver its seam frame forever (round-1 advisor finding)."""
    from ya_vo_b200 import sharding
    frames = synth.synth_batch(3, "U", 1, 12, 20)
    with pytest.raises(ValueError):
        stream.FrameStream(3, lambda i: frames[i], batch=1, shape=(12, 20))
    with pytest.raises(ValueError):
        sharding.process_shard(None, lambda a, b: frames[a:b], 3, 0, 1, batch=1)
