"""CPU: the arithmetic of the tensor-core matchers (ya_vo_b200/csrc/match_tc4.cuh, match_tc.cuh) restated in numpy and
checked against the oracle's Brief::matchFeatures (reference src/BriefDescriptor.cc:139-183).

What the kernels rely on, and what is checked here without a GPU:
  * the expansion of descriptor bits to +-1 operands (e2m1 nibbles / e4m3 bytes) in the kernels' permuted K order,
    decoded back to numbers, gives  sum_k a_k * b_k * scale = 256 * hamming - 32768;
  * the constant index slice adds exactly the train column j (base-4 digits x weights in FP4, j & 15 / j >> 4 in FP8);
  * every partial sum is an integer below 2^24, so float32 accumulation in ANY order is exact;
  * min over a tile row of the accumulator, strict '<' on the distance across tiles, and the lexicographic combine of the
    two epilogue warp groups reproduce "lowest index among equal minima".
"""
import numpy as np
import pytest

from ya_vo_b200 import synth

E2M1 = {0x0: 0.0, 0x1: 0.5, 0x2: 1.0, 0x3: 1.5, 0x4: 2.0, 0x5: 3.0, 0x6: 4.0, 0x7: 6.0}


def e2m1(n):
    n = np.asarray(n, np.uint32)
    mag = np.vectorize(lambda v: E2M1[int(v) & 7])(n).astype(np.float32)
    return np.where(n & 8, -mag, mag).astype(np.float32)


def e4m3(b):
    b = np.asarray(b, np.uint32)
    e, m = (b >> 3) & 15, b & 7
    mag = np.where(e == 0, m / 8.0 * 2.0 ** -6, (1 + m / 8.0) * 2.0 ** (e.astype(np.float64) - 7))
    return np.where(b & 0x80, -mag, mag).astype(np.float32)


def words(desc):
    return np.ascontiguousarray(desc, np.uint8).view("<u4").reshape(desc.shape[0], 8)


def expand_fp4(desc, train):
    """expand_row4: output word s of source word w = ((w << (3 - s)) & 0x88888888) | 0x22222222 (queries) or ^ 0xAAAAAAAA
    (train); eight nibbles per output word, four output words per source word -> 256 numbers per descriptor."""
    w = words(desc)
    out = []
    for s in range(4):
        x = (w << np.uint32(3 - s)) & np.uint32(0x88888888)
        x = (x ^ np.uint32(0xAAAAAAAA)) if train else (x | np.uint32(0x22222222))
        out.append(np.stack([(x >> np.uint32(4 * q)) & np.uint32(15) for q in range(8)], -1))  # [n, 8 words, 8 nibbles]
    nib = np.stack(out, 2)  # [n, word, s, q]
    return e2m1(nib.reshape(desc.shape[0], 256))


def expand_fp8(desc, train):
    """expand_half: output word s of source word w = ((w << (7 - s)) & 0x80808080) | 0x38383838 or ^ 0xF0F0F0F0."""
    w = words(desc)
    out = []
    for s in range(8):
        x = (w << np.uint32(7 - s)) & np.uint32(0x80808080)
        x = (x ^ np.uint32(0xF0F0F0F0)) if train else (x | np.uint32(0x38383838))
        out.append(np.stack([(x >> np.uint32(8 * q)) & np.uint32(255) for q in range(4)], -1))
    byt = np.stack(out, 2)
    return e4m3(byt.reshape(desc.shape[0], 256))


def index_slice_fp4(query, j):
    """index_slice(): 22 nibbles, weights {1, 4, 4 x4, 4 x16} against the base-4 digits of j."""
    enc = {0: 0, 1: 2, 2: 4, 3: 5, 4: 6}
    vals = []
    for e in range(22):
        digit = (j & 3) if e == 0 else ((j >> 2) & 3) if e == 1 else ((j >> 4) & 3) if e < 6 else ((j >> 6) & 3)
        vals.append(enc[(1 if e == 0 else 4) if query else digit])
    return e2m1(np.array(vals))


def hamming(d1, d2):
    x = d1[:, None, :] ^ d2[None, :, :]
    return np.unpackbits(x, axis=2).sum(2).astype(np.int64)


@pytest.mark.parametrize("fmt", ["fp4", "fp8"])
def test_accumulator_is_the_key(fmt):
    rng = np.random.default_rng(5)
    d1 = synth.synth_descriptors(96, 1)
    d2 = synth.planted_descriptors(d1, 224, 7)
    d1[3] = d2[5]
    d1[4] = ~d2[5]
    h = hamming(d1, d2)
    if fmt == "fp4":
        a, b = expand_fp4(d1, False), expand_fp4(d2, True) * np.float32(128.0)  # train scale factors: 2^7
        assert set(np.unique(a)) <= {-1.0, 1.0} and set(np.unique(b)) <= {-128.0, 128.0}
        ax = index_slice_fp4(True, 0)
        bx = np.stack([index_slice_fp4(False, j) for j in range(224)])
        col = bx @ ax
    else:
        a, b = expand_fp8(d1, False), expand_fp8(d2, True)
        assert set(np.unique(a)) <= {-1.0, 1.0} and set(np.unique(b)) <= {-128.0, 128.0}
        j = np.arange(224)
        # query bytes {1.0 = 0x38, 16.0 = 0x58}, train bytes {j & 15, j >> 4} as e4m3 small integers
        assert e4m3(np.array([0x38, 0x58])).tolist() == [1.0, 16.0]
        col = (j & 15) * 1.0 + (j >> 4) * 16.0
    assert np.array_equal(col, np.arange(224, dtype=np.float32))
    # float32 accumulation, in a scrambled order: exact all the same
    perm = rng.permutation(256)
    acc = (a[:, perm].astype(np.float32) @ b[:, perm].T.astype(np.float32)) + col[None, :].astype(np.float32)
    assert acc.dtype == np.float32
    assert np.array_equal(acc, (256 * h + np.arange(224)[None, :] - 32768).astype(np.float32))
    assert np.abs(acc).max() < 2 ** 24


@pytest.mark.parametrize("n1,n2,tile", [(200, 1950, 224), (130, 700, 256), (64, 225, 224), (50, 224, 224), (40, 3, 224)])
def test_first_minimum_through_tiles_and_warp_groups(oracle, n1, n2, tile):
    d1 = synth.synth_descriptors(n1, n1 * 131 + n2)
    d2 = synth.planted_descriptors(d1, n2, 7)
    if n2 > 300:
        d2[290] = d2[10]
        d2[40] = d2[10]
    h = hamming(d1, d2)
    # per epilogue warp group: (distance, index) per row
    best = [np.full((n1, 2), [0x7FFFFFFF, -1], np.int64) for _ in range(2)]
    half = tile // 2
    for t in range((n2 + tile - 1) // tile):
        nvalid = min(tile, n2 - t * tile)
        key = 256 * h[:, t * tile:t * tile + nvalid] + np.arange(nvalid)[None, :]
        for g in range(2):  # group g drains columns [g*half, (g+1)*half) of every tile
            lo, hi = g * half, min((g + 1) * half, nvalid)
            if lo >= hi:
                continue  # this group's columns all lie beyond the train set
            m = key[:, lo:hi].min(1)
            upd = (m >> 8) < best[g][:, 0]
            best[g][upd, 0] = (m >> 8)[upd]
            best[g][upd, 1] = t * tile + (m & 255)[upd]
    a, b = best
    take_b = (b[:, 0] < a[:, 0]) | ((b[:, 0] == a[:, 0]) & (b[:, 1] < a[:, 1]))
    out = np.where(take_b[:, None], b, a)
    eidx, edist = oracle.match(d1, d2)
    assert np.array_equal(out[:, 1], eidx) and np.array_equal(out[:, 0], edist)
