"""CPU: the arithmetic of the tensor-core blur variants (ya_vo_b200/csrc/blur_umma.cuh, build variants -DYAVO_BLUR_UMMA=1 / 2)
restated in numpy and checked against the oracle's cv::GaussianBlur restatement (reference src/BriefDescriptor.cc:90).

What the kernels rely on, and what is checked here without a GPU:
  * pass 1 as five K = 32 products against row-shifted views of ONE 256 x 32 band matrix F (F[r][j] = g[j + 116 - r]):
    sum_k F[x + 128 - 32 k][j] * P[n][32 k + j] is the horizontal 9-tap sum of output column x from the staged row n
    (pixels start at staged byte 16), for every x of the 128-column tile;
  * the shared-memory images of F and T2 built by bu_fill_constants: K-major core matrices (8 rows x 16 bytes, 128 bytes
    apart; 16-byte K chunks LBO bytes apart);
  * the 16-bit row sums split into low / high bytes, two products against T2[r][j] = g[j - r] over a 32-row window, and
    out = (lo + 256 * hi + 32768) >> 16 give the blurred pixel, for both 16-row halves of a tile (all integer, exact);
  * the hybrid's vertical pass on pairs of row sums (IDP.2A form) is the same number.
"""
import numpy as np

from oracle import pyoracle as po
from ya_vo_b200 import synth

G = np.array([12, 22, 31, 41, 44, 41, 31, 22, 12], np.int64)
F_ROWS, F_LBO, T2_LBO = 256, 256 * 16, 16 * 16


def fill_constants():
    """bu_fill_constants of blur_umma.cuh, byte for byte."""
    buf = np.zeros(2 * F_LBO + 2 * T2_LBO, np.uint8)
    for r in range(F_ROWS):
        for j in range(32):
            t = j + 116 - r
            if 0 <= t <= 8:
                buf[(j // 16) * F_LBO + (r // 8) * 128 + (r % 8) * 16 + j % 16] = G[t]
    for r in range(16):
        for j in range(32):
            t = j - r
            if 0 <= t <= 8:
                buf[2 * F_LBO + (j // 16) * T2_LBO + (r // 8) * 128 + (r % 8) * 16 + j % 16] = G[t]
    return buf


def read_kmajor(buf, base, lbo, rows, row0=0):
    """rows x 32 operand read the way a K-major no-swizzle descriptor (start = base + row0 * 16 for row0 % 8 == 0) does."""
    out = np.zeros((rows, 32), np.int64)
    for r in range(rows):
        rr = row0 + r
        for j in range(32):
            out[r, j] = buf[base + (j // 16) * lbo + (rr // 8) * 128 + (rr % 8) * 16 + j % 16]
    return out


def staged_tile(img, x0, y0):
    """40 rows x 160 bytes: image rows y0-4 .. y0+35 and columns x0-16 .. x0+143, BORDER_REFLECT_101 where outputs read."""
    H, W = img.shape

    def refl(p, n):
        while p < 0 or p >= n:
            p = -p if p < 0 else 2 * (n - 1) - p
        return p
    t = np.zeros((40, 160), np.int64)
    for n in range(40):
        gr = refl(y0 - 4 + n, H)
        for c in range(160):
            gc = x0 - 16 + c
            if -4 <= gc < W + 4:
                t[n, c] = img[gr, refl(gc, W)]
    return t


def test_band_matrix_views_give_the_horizontal_pass_and_byte_split_gives_the_blur():
    img = synth.synth_frame("G30", 9, 70, 300)
    img[10:40, 100:250] = 255  # the largest row sums
    exp = po.gaussian_blur(img)
    cst = fill_constants()
    for x0, y0 in ((0, 0), (128, 32), (256, 64)):
        P = staged_tile(img, x0, y0)
        D1 = np.zeros((128, 40), np.int64)  # [x][staged row]
        for k in range(5):
            A = read_kmajor(cst, 0, F_LBO, 128, row0=128 - 32 * k)       # start address advanced by (16 - 4k) * 128 bytes
            D1 += A @ P[:, 32 * k:32 * k + 32].T
        ref_h = np.stack([sum(G[t] * P[:, 12 + x + t] for t in range(9)) for x in range(128)])
        assert np.array_equal(D1, ref_h) and D1.max() < 65536
        T2 = read_kmajor(cst, 2 * F_LBO, T2_LBO, 16)
        lo, hi = D1 & 255, D1 >> 8
        for half in range(2):
            w = slice(16 * half, 16 * half + 32)
            lo_w = np.zeros((128, 32), np.int64)
            hi_w = np.zeros((128, 32), np.int64)
            n_have = min(32, 40 - 16 * half)
            lo_w[:, :n_have], hi_w[:, :n_have] = lo[:, w][:, :n_have], hi[:, w][:, :n_have]
            lo_w[:, n_have:] = 173  # pad rows hold garbage in TMEM: their taps are zero
            out = (lo_w @ T2.T + 256 * (hi_w @ T2.T) + 32768) >> 16
            for r in range(16):
                gr = y0 + 16 * half + r
                if gr >= img.shape[0]:
                    continue
                for x in range(128):
                    if x0 + x < img.shape[1]:
                        assert out[x, r] == exp[gr, x0 + x], (x0, y0, half, r, x)
        # hybrid: vertical pass on pairs of row sums, two outputs per five pairs (yavo_blur_v2_raw)
        for t in range(0, 32, 2):
            a = 32768 + sum(G[k] * D1[:, t + k] for k in range(9))
            b = 32768 + sum(G[k] * D1[:, t + 1 + k] for k in range(9))
            for x in range(0, 128, 17):
                if y0 + t < img.shape[0] and x0 + x < img.shape[1]:
                    assert (a[x] >> 16) == exp[y0 + t, x0 + x]
                if y0 + t + 1 < img.shape[0] and x0 + x < img.shape[1]:
                    assert (b[x] >> 16) == exp[y0 + t + 1, x0 + x]
