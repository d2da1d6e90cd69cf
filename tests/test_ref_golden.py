"""Outputs of the reference's OWN sources (src/FastDetector.cc, src/BriefDescriptor.cc, src/Image.cc compiled
unmodified over the stub OpenCV tree, oracle/ref_shim) stored in tests/golden/ref_golden.npz:
  * CPU: the oracle reproduces them (this is what pins the oracle to the reference's control flow);
  * CPU, build container only: the live shimmed reference agrees with the stored fixture;
  * GPU: the CUDA path reproduces them through the C ABI.
"""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN)


def frames(g):
    return [str(n) for n in g["names"]]


def test_fixture_covers_the_cut_and_ties(gold):
    ns = {n: gold[n + "_fast_rows"].size for n in frames(gold)}
    assert max(ns.values()) == 2000  # a frame with more than fastCornerNumThreshold candidates
    assert len(ns) >= 8


def test_oracle_reproduces_reference_fast_and_brief(oracle, gold):
    off = gold["offsets"]
    for n in frames(gold):
        img = gold[n + "_img"]
        r, c, s, nc = oracle.fast_detect(img, 2000)
        assert np.array_equal(r, gold[n + "_fast_rows"]) and np.array_equal(c, gold[n + "_fast_cols"]), n
        assert np.array_equal(s[:64].view(np.uint32), gold[n + "_harris64"].view(np.uint32)[:s[:64].size]), n
        d, v, _ = oracle.brief(img, off, r, c)
        assert np.array_equal(np.nonzero(v)[0], gold[n + "_kp_id"]), n
        assert np.array_equal(r[v], gold[n + "_kp_x"]) and np.array_equal(c[v], gold[n + "_kp_y"]), n
        assert np.array_equal(d[v], gold[n + "_desc"]), n


def test_oracle_reproduces_reference_match(oracle, gold):
    names = frames(gold)
    idx, dist = oracle.match(gold[names[2] + "_desc"], gold[names[3] + "_desc"])
    assert np.array_equal(idx, gold["match_a_idx"]) and np.array_equal(dist, gold["match_a_dist"])
    assert np.array_equal(oracle.remove_outliers(dist, 20), gold["match_a_keep"])
    idx, dist = oracle.match(gold["match_b_d1"], gold["match_b_d2"])
    assert np.array_equal(idx, gold["match_b_idx"]) and np.array_equal(dist, gold["match_b_dist"])
    assert np.array_equal(oracle.remove_outliers(dist, 20), gold["match_b_keep"])


def test_live_shimmed_reference_matches_fixture(gold):
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref not built (needs the reference checkout)")
    for n in frames(gold)[:5]:
        img = gold[n + "_img"]
        r, c = pyref.fast(img)
        assert np.array_equal(r, gold[n + "_fast_rows"]) and np.array_equal(c, gold[n + "_fast_cols"])
        x, y, ids, desc = pyref.brief(img, gold["offsets"], r, c)
        assert np.array_equal(desc, gold[n + "_desc"]) and np.array_equal(ids, gold[n + "_kp_id"])
    cnt, pts = pyref.ring(25, 25)
    assert cnt == 16
    # the reference's known-answer tests (tests/FastDetectorTest.cc:38-80) through its own code
    img = np.zeros((50, 50), np.uint8)
    for x, y in pts:
        img[y, x] = 255
    assert pyref.check_contiguous(img, 25, 25) is True
    img[25, 25] = 255
    assert pyref.check_contiguous(img, 25, 25) is False


def test_live_reference_stream_agrees_with_oracle(oracle, gold):
    """The steady-state timing form bench.py --impl reference uses (oracle/_ref ref_stream_push) produces the
    oracle's keypoint counts and kept-match counts on a short frame sequence."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref not built (needs the reference checkout)")
    import subprocess, sys, json, textwrap
    names = frames(gold)[:3]
    # the reference prints on every call: run it in a child whose stdout is discarded
    code = textwrap.dedent("""
        import os, sys, json, numpy as np
        keep = os.dup(1); os.dup2(os.open(os.devnull, os.O_WRONLY), 1)
        sys.path.insert(0, %r)
        from oracle import pyref
        g = np.load(%r)
        st = pyref.Stream(g["offsets"])
        out = [st.push(g[n + "_img"]) for n in %r]
        os.dup2(keep, 1)
        print(json.dumps(out))
    """ % (os.path.dirname(os.path.dirname(os.path.dirname(GOLDEN))), GOLDEN, names))
    got = json.loads(subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1])
    off = gold["offsets"]
    prev = None
    for n, (nk, kept) in zip(names, got):
        img = gold[n + "_img"]
        r, c, s, nc = oracle.fast_detect(img, 2000)
        d, v, _ = oracle.brief(img, off, r, c)
        assert nk == int(v.sum()), n
        if prev is not None and prev.shape[0] and int(v.sum()):
            idx, dist = oracle.match(prev, d[v])
            assert kept == int(np.count_nonzero(oracle.remove_outliers(dist, 20))), n
        else:
            assert kept == 0
        prev = d[v]


@pytest.mark.gpu
def test_cuda_reproduces_reference_outputs(cuda_lib, gold):
    off = gold["offsets"]
    with cuda_lib.Context(device=0, n_slots=2, max_rows=256, max_cols=512, max_kp=2000) as ctx:
        ctx.set_brief_offsets(off)
        for n in frames(gold):
            img = gold[n + "_img"]
            ctx.upload(0, img)
            r, c, s, nc = ctx.fast_detect(0)
            assert np.array_equal(r, gold[n + "_fast_rows"]) and np.array_equal(c, gold[n + "_fast_cols"]), n
            assert np.array_equal(s[:64].view(np.uint32), gold[n + "_harris64"].view(np.uint32)[:s[:64].size]), n
            d, v, _ = ctx.brief_describe(0, r, c)
            assert np.array_equal(np.nonzero(v)[0], gold[n + "_kp_id"]), n
            assert np.array_equal(d[v], gold[n + "_desc"]), n
        names = frames(gold)
        idx, dist = ctx.match(gold[names[2] + "_desc"], gold[names[3] + "_desc"])
        assert np.array_equal(idx, gold["match_a_idx"]) and np.array_equal(dist, gold["match_a_dist"])
        assert np.array_equal(cuda_lib.remove_outliers(dist, 20), gold["match_a_keep"])
        idx, dist = ctx.match(gold["match_b_d1"], gold["match_b_d2"])
        assert np.array_equal(idx, gold["match_b_idx"]) and np.array_equal(dist, gold["match_b_dist"])
