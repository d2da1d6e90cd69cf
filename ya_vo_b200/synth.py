"""Seeded synthetic inputs.  Pure numpy; shared by tests/, bench.py and the
golden-fixture generator so that every leg sees identical pixels."""
import os

import numpy as np

KITTI_H, KITTI_W = 376, 1241
_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def synth_frame(kind, seed, H=KITTI_H, W=KITTI_W):
    """kind: 'U' uniform noise; 'G30' clip(normal(128,30)); 'B4' uniform noise at 1/4 resolution,
    nearest-upsampled x4 (blocky: many tied Harris scores)."""
    rng = np.random.default_rng(seed)
    if kind == "U":
        return rng.integers(0, 256, (H, W), dtype=np.uint8)
    if kind == "G30":
        return np.clip(np.rint(rng.normal(128.0, 30.0, (H, W))), 0, 255).astype(np.uint8)
    if kind == "B4":
        h4, w4 = (H + 3) // 4, (W + 3) // 4
        small = rng.integers(0, 256, (h4, w4), dtype=np.uint8)
        return np.ascontiguousarray(np.repeat(np.repeat(small, 4, axis=0), 4, axis=1)[:H, :W])
    raise ValueError(kind)


def shifted_pair(frame, seed, drow=1, dcol=3):
    """Config 2: frame B = A shifted by (+drow, +dcol) with edge replication plus integers(-2,3) noise."""
    H, W = frame.shape
    rr = np.clip(np.arange(H) - drow, 0, H - 1)
    cc = np.clip(np.arange(W) - dcol, 0, W - 1)
    b = frame[rr][:, cc].astype(np.int16)
    noise = np.random.default_rng(seed).integers(-2, 3, (H, W)).astype(np.int16)
    return np.clip(b + noise, 0, 255).astype(np.uint8)


def synth_batch(n_frames, kind="G30", seed0=1000, H=KITTI_H, W=KITTI_W):
    """Config 3: frame f drawn from seed seed0+f."""
    out = np.empty((n_frames, H, W), np.uint8)
    for f in range(n_frames):
        out[f] = synth_frame(kind, seed0 + f, H, W)
    return out


def brief_offsets():
    """The fixed 256x4 table {drow1,dcol1,drow2,dcol2} in [-8,8] (tests/golden/brief_offsets.npy);
    identical to default_rng(7).integers(-8, 9, (256, 4))."""
    p = os.path.join(_GOLDEN, "brief_offsets.npy")
    if os.path.exists(p):
        return np.load(p).astype(np.int32)
    return np.random.default_rng(7).integers(-8, 9, (256, 4)).astype(np.int32)


def synth_descriptors(n, seed):
    return np.random.default_rng(seed).integers(0, 256, (n, 32), dtype=np.uint8)


def planted_descriptors(d1, n2, seed, frac=0.5, flips=10):
    """Config 5 'planted' variant: frac of set 2 are set-1 rows with `flips` random bits flipped."""
    rng = np.random.default_rng(seed)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    n_pl = int(n2 * frac)
    src = rng.integers(0, d1.shape[0], n_pl)
    pos = rng.permutation(n2)[:n_pl]
    planted = d1[src].copy()
    for k in range(n_pl):
        bits = rng.choice(256, flips, replace=False)
        for b in bits:
            planted[k, b // 8] ^= np.uint8(1 << (b % 8))
    d2[pos] = planted
    return d2
