// fast_core.h — per-thread arithmetic of the detect kernel (segment test, Harris response,
// fixed-point Gaussian taps).  Host/device: the device build uses the native byte-SIMD and
// IEEE round-to-nearest intrinsics, the host build (tests/emul) plain C with the same rounding,
// so the bit logic can be checked on the CPU before it ever reaches a GPU.
#ifndef YAVO_FAST_CORE_H
#define YAVO_FAST_CORE_H

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define YAVO_HD __host__ __device__ __forceinline__
#else
#ifndef YAVO_HD
#define YAVO_HD inline
#endif
#endif

// ---- byte-SIMD helpers -----------------------------------------------------------------------
YAVO_HD uint32_t yavo_absdiff4(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __vabsdiffu4(a, b);  // VABSDIFF4.U8, native on sm_100a
#else
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        int x = (a >> (8 * i)) & 255, y = (b >> (8 * i)) & 255;
        r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i);
    }
    return r;
#endif
}
YAVO_HD uint32_t yavo_perm(uint32_t lo, uint32_t hi, uint32_t sel) {
#ifdef __CUDA_ARCH__
    return __byte_perm(lo, hi, sel);
#else
    uint64_t v = ((uint64_t)hi << 32) | lo;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 255) << (8 * i);
    return r;
#endif
}
YAVO_HD uint32_t yavo_dp4a(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    return __dp4a(a, b, c);
#else
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 255) * ((b >> (8 * i)) & 255);
    return c;
#endif
}
// a = two u16, b = four u8: lo uses bytes 0,1 of b, hi uses bytes 2,3
YAVO_HD uint32_t yavo_dp2a_lo(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    return __dp2a_lo(a, b, c);
#else
    return c + (a & 0xffff) * (b & 255) + (a >> 16) * ((b >> 8) & 255);
#endif
}
YAVO_HD uint32_t yavo_dp2a_hi(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    return __dp2a_hi(a, b, c);
#else
    return c + (a & 0xffff) * ((b >> 16) & 255) + (a >> 16) * ((b >> 24) & 255);
#endif
}

// bit 7 of every byte of the result is set iff |c - p| >= 40 for that byte
// (a ring pixel "differs", reference src/FastDetector.cc:155-161 with intensityThreshold 40)
YAVO_HD uint32_t yavo_differs4(uint32_t c4, uint32_t p4) {
    uint32_t d = yavo_absdiff4(c4, p4);
    return (((d & 0x7f7f7f7fu) + 0x58585858u) | d) & 0x80808080u;  // 0x58 = 128 - 40
}

// Bytes x+dy .. x+dy+3 of a row, given the three aligned words covering x-4 .. x+7 (x % 4 == 0).
template <int DY>
YAVO_HD uint32_t yavo_shift_bytes(uint32_t wm, uint32_t w0, uint32_t wp) {
    if (DY == 0) return w0;
    if (DY == -3) return yavo_perm(wm, w0, 0x4321);
    if (DY == -2) return yavo_perm(wm, w0, 0x5432);
    if (DY == -1) return yavo_perm(wm, w0, 0x6543);
    if (DY == 1) return yavo_perm(w0, wp, 0x4321);
    if (DY == 2) return yavo_perm(w0, wp, 0x5432);
    return yavo_perm(w0, wp, 0x6543);  // DY == 3
}

// Segment test for 4 horizontally adjacent pixels.  rows[d+3] points at the aligned word holding
// pixel x of image row r+d (d = -3..3), so rows[.][-1], [0], [1] cover bytes x-4 .. x+7.
// Ring order (row offset, col offset), reference src/FastDetector.cc:50-112:
//  0:(0,-3) 1:(1,-3) 2:(2,-2) 3:(3,-1) 4:(3,0) 5:(3,1) 6:(2,2) 7:(1,3)
//  8:(0,3) 9:(-1,3) 10:(-2,2) 11:(-3,1) 12:(-3,0) 13:(-3,-1) 14:(-2,-2) 15:(-1,-3)
// A pixel is a corner iff ring 0 and 7 differ, ring 4 or 12 differs, and 12 consecutive ring
// indices (no wrap-around) differ (:304-320, :135-153).  Every 12-window inside 0..15 contains
// indices 4..11, so the condition equals D0 & (D4&..&D11) & OR_{s=0..4} (window s), and the
// pre-tests on 7 and 4 are implied by the common core.
// Returns a 4-bit mask (bit b = pixel x+b).
template <typename WP>
YAVO_HD uint32_t yavo_fast4(const WP rm3, const WP rm2, const WP rm1, const WP r0, const WP rp1,
                            const WP rp2, const WP rp3, bool *any_pre) {
    const uint32_t c4 = r0[0];
    const uint32_t a0 = r0[-1], a1 = r0[1];
    const uint32_t d0 = yavo_differs4(c4, yavo_shift_bytes<-3>(a0, c4, a1));
    const uint32_t d8 = yavo_differs4(c4, yavo_shift_bytes<3>(a0, c4, a1));
    const uint32_t b0 = rp1[-1], b1 = rp1[0], b2 = rp1[1];
    const uint32_t d7 = yavo_differs4(c4, yavo_shift_bytes<3>(b0, b1, b2));
    const uint32_t d4 = yavo_differs4(c4, rp3[0]);
    uint32_t core = d0 & d7 & d4 & d8;
    *any_pre = core != 0;
    if (core == 0) return 0;
    const uint32_t d1 = yavo_differs4(c4, yavo_shift_bytes<-3>(b0, b1, b2));
    const uint32_t e0 = rp2[-1], e1 = rp2[0], e2 = rp2[1];
    const uint32_t d2 = yavo_differs4(c4, yavo_shift_bytes<-2>(e0, e1, e2));
    const uint32_t d6 = yavo_differs4(c4, yavo_shift_bytes<2>(e0, e1, e2));
    const uint32_t f0 = rp3[-1], f1 = rp3[0], f2 = rp3[1];
    const uint32_t d3 = yavo_differs4(c4, yavo_shift_bytes<-1>(f0, f1, f2));
    const uint32_t d5 = yavo_differs4(c4, yavo_shift_bytes<1>(f0, f1, f2));
    const uint32_t g0 = rm1[-1], g1 = rm1[0], g2 = rm1[1];
    const uint32_t d9 = yavo_differs4(c4, yavo_shift_bytes<3>(g0, g1, g2));
    const uint32_t d15 = yavo_differs4(c4, yavo_shift_bytes<-3>(g0, g1, g2));
    const uint32_t h0 = rm2[-1], h1 = rm2[0], h2 = rm2[1];
    const uint32_t d10 = yavo_differs4(c4, yavo_shift_bytes<2>(h0, h1, h2));
    const uint32_t d14 = yavo_differs4(c4, yavo_shift_bytes<-2>(h0, h1, h2));
    const uint32_t i0 = rm3[-1], i1 = rm3[0], i2 = rm3[1];
    const uint32_t d11 = yavo_differs4(c4, yavo_shift_bytes<1>(i0, i1, i2));
    const uint32_t d12 = yavo_differs4(c4, i1);
    const uint32_t d13 = yavo_differs4(c4, yavo_shift_bytes<-1>(i0, i1, i2));
    core &= d5 & d6 & d9 & d10 & d11;           // d0 & d4..d11
    const uint32_t lo = d1 & d2 & d3;           // window 0 (d0 already in core)
    const uint32_t w1 = lo & d12;               // 1..12
    const uint32_t w2 = d2 & d3 & d12 & d13;    // 2..13
    const uint32_t w3 = d3 & d12 & d13 & d14;   // 3..14
    const uint32_t w4 = d12 & d13 & d14 & d15;  // 4..15
    const uint32_t r = core & (lo | w1 | w2 | w3 | w4);
    // gather bit 7 of each byte into a nibble
    return (((r >> 7) & 0x01010101u) * 0x01020408u) >> 24;
}

// The same test for a quad whose necessary condition `core` = d0 & d7 & d4 & d8 (yavo_fast4_core, non-zero) is already
// known: the remaining twelve ring positions only.
template <typename WP>
YAVO_HD uint32_t yavo_fast4_rest(const WP rm3, const WP rm2, const WP rm1, const WP r0, const WP rp1, const WP rp2, const WP rp3,
                                 uint32_t core) {
    const uint32_t c4 = r0[0];
    const uint32_t b0 = rp1[-1], b1 = rp1[0];
    const uint32_t d1 = yavo_differs4(c4, yavo_shift_bytes<-3>(b0, b1, 0u));
    const uint32_t e0 = rp2[-1], e1 = rp2[0], e2 = rp2[1];
    const uint32_t d2 = yavo_differs4(c4, yavo_shift_bytes<-2>(e0, e1, e2));
    const uint32_t d6 = yavo_differs4(c4, yavo_shift_bytes<2>(e0, e1, e2));
    const uint32_t f0 = rp3[-1], f1 = rp3[0], f2 = rp3[1];
    const uint32_t d3 = yavo_differs4(c4, yavo_shift_bytes<-1>(f0, f1, f2));
    const uint32_t d5 = yavo_differs4(c4, yavo_shift_bytes<1>(f0, f1, f2));
    const uint32_t g0 = rm1[-1], g1 = rm1[0], g2 = rm1[1];
    const uint32_t d9 = yavo_differs4(c4, yavo_shift_bytes<3>(g0, g1, g2));
    const uint32_t d15 = yavo_differs4(c4, yavo_shift_bytes<-3>(g0, g1, g2));
    const uint32_t h0 = rm2[-1], h1 = rm2[0], h2 = rm2[1];
    const uint32_t d10 = yavo_differs4(c4, yavo_shift_bytes<2>(h0, h1, h2));
    const uint32_t d14 = yavo_differs4(c4, yavo_shift_bytes<-2>(h0, h1, h2));
    const uint32_t i0 = rm3[-1], i1 = rm3[0], i2 = rm3[1];
    const uint32_t d11 = yavo_differs4(c4, yavo_shift_bytes<1>(i0, i1, i2));
    const uint32_t d12 = yavo_differs4(c4, i1);
    const uint32_t d13 = yavo_differs4(c4, yavo_shift_bytes<-1>(i0, i1, i2));
    core &= d5 & d6 & d9 & d10 & d11;           // d0 & d4..d11
    const uint32_t lo = d1 & d2 & d3;           // window 0 (d0 already in core)
    const uint32_t w1 = lo & d12;               // 1..12
    const uint32_t w2 = d2 & d3 & d12 & d13;    // 2..13
    const uint32_t w3 = d3 & d12 & d13 & d14;   // 3..14
    const uint32_t w4 = d12 & d13 & d14 & d15;  // 4..15
    const uint32_t r = core & (lo | w1 | w2 | w3 | w4);
    return (((r >> 7) & 0x01010101u) * 0x01020408u) >> 24;
}

// The four ring positions every 12-window needs besides D0 (0, 4, 7, 8): a cheap necessary condition.  The
// detect kernel runs it for every quad and the full test only for the quads that survive it.
template <typename WP>
YAVO_HD uint32_t yavo_fast4_core(const WP r0, const WP rp1, const WP rp3) {
    const uint32_t c4 = r0[0];
    const uint32_t a0 = r0[-1], a1 = r0[1];
    const uint32_t d0 = yavo_differs4(c4, yavo_shift_bytes<-3>(a0, c4, a1));
    const uint32_t d8 = yavo_differs4(c4, yavo_shift_bytes<3>(a0, c4, a1));
    const uint32_t d7 = yavo_differs4(c4, yavo_shift_bytes<3>(rp1[-1], rp1[0], rp1[1]));
    const uint32_t d4 = yavo_differs4(c4, rp3[0]);
    return d0 & d7 & d4 & d8;
}

// scalar form of the same test (one pixel, ring values in ring order) — used for checks
YAVO_HD bool yavo_fast1(int c, const int ring[16]) {
    uint32_t d = 0;
    for (int k = 0; k < 16; k++) {
        int diff = c - ring[k];
        if (diff < 0) diff = -diff;
        if (diff >= 40) d |= 1u << k;
    }
    if (!((d & 1u) && (d & (1u << 7)))) return false;
    if (!((d & (1u << 4)) || (d & (1u << 12)))) return false;
    uint32_t m2 = d & (d >> 1);
    uint32_t m4 = m2 & (m2 >> 2);
    uint32_t m8 = m4 & (m4 >> 4);
    return (m8 & (m4 >> 8)) != 0;  // 12 consecutive set bits, no wrap
}

// ---- Harris response ---------------------------------------------------------------------------
// IEEE round-to-nearest single operations that the compiler may not contract or reassociate.
#ifdef __CUDA_ARCH__
#define YAVO_FMUL(a, b) __fmul_rn((a), (b))
#define YAVO_FADD(a, b) __fadd_rn((a), (b))
#define YAVO_FSUB(a, b) __fsub_rn((a), (b))
#define YAVO_FDIV(a, b) __fdiv_rn((a), (b))
#define YAVO_FSQRT(a) __fsqrt_rn((a))
#define YAVO_DMUL(a, b) __dmul_rn((a), (b))
#define YAVO_DSUB(a, b) __dsub_rn((a), (b))
#else  // host: build with -ffp-contract=off
#define YAVO_FMUL(a, b) ((float)((a) * (b)))
#define YAVO_FADD(a, b) ((float)((a) + (b)))
#define YAVO_FSUB(a, b) ((float)((a) - (b)))
#define YAVO_FDIV(a, b) ((float)((a) / (b)))
#define YAVO_FSQRT(a) (sqrtf((a)))
#define YAVO_DMUL(a, b) ((double)((a) * (b)))
#define YAVO_DSUB(a, b) ((double)((a) - (b)))
#endif

// OpenCV lapack.cpp hypot<float>
YAVO_HD float yavo_cv_hypot(float a, float b) {
    a = fabsf(a);
    b = fabsf(b);
    if (a > b) {
        b = YAVO_FDIV(b, a);
        return YAVO_FMUL(a, YAVO_FSQRT(YAVO_FADD(1.0f, YAVO_FMUL(b, b))));
    }
    if (b > 0.0f) {
        a = YAVO_FDIV(a, b);
        return YAVO_FMUL(b, YAVO_FSQRT(YAVO_FADD(1.0f, YAVO_FMUL(a, a))));
    }
    return 0.0f;
}

// Response of reference src/FastDetector.cc:244-273 from the integer structure tensor
// [a b; b c] (3x3 box sums of Sobel products; exact in float32 because < 2^24):
// cv::eigen (OpenCV JacobiImpl_<float>, n = 2) then l1*l2 - 0.04*(l1+l2)^2 in the
// reference's float/double mix.
YAVO_HD float yavo_harris_from_tensor(int ia, int ib, int ic) {
    float W0 = (float)ia, W1 = (float)ic;
    const float p = (float)ib;
    if (!(fabsf(p) <= 1.1920928955078125e-07f)) {
        float y = YAVO_FMUL(YAVO_FSUB(W1, W0), 0.5f);
        float t = YAVO_FADD(fabsf(y), yavo_cv_hypot(p, y));
        t = YAVO_FMUL(YAVO_FDIV(p, t), p);
        if (y < 0.0f) t = -t;
        W0 = YAVO_FSUB(W0, t);
        W1 = YAVO_FADD(W1, t);
    }
    const float l1 = W0 < W1 ? W1 : W0, l2 = W0 < W1 ? W0 : W1;
    const float prod = YAVO_FMUL(l1, l2);
    const float sum = YAVO_FADD(l2, l1);
    const double sq = YAVO_DMUL((double)sum, (double)sum);
    const double r = YAVO_DSUB((double)prod, YAVO_DMUL(0.04, sq));
    return (float)r;
}

// Structure tensor at (row, col) from raw pixels: Sobel 3x3 correlation (:164-214) at the nine
// neighbours, products summed (:255-262).  `px(r, c)` must be valid for a 5x5 window.
template <typename F>
YAVO_HD void yavo_structure_tensor(F px, int row, int col, int *oa, int *ob, int *oc) {
    int v[5][5];
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) v[i][j] = px(row - 2 + i, col - 2 + j);
    int a = 0, b = 0, c = 0;
    for (int i = 1; i <= 3; i++)
        for (int j = 1; j <= 3; j++) {
            int gx = (v[i - 1][j + 1] - v[i - 1][j - 1]) + 2 * (v[i][j + 1] - v[i][j - 1]) +
                     (v[i + 1][j + 1] - v[i + 1][j - 1]);
            int gy = (v[i + 1][j - 1] + 2 * v[i + 1][j] + v[i + 1][j + 1]) -
                     (v[i - 1][j - 1] + 2 * v[i - 1][j] + v[i - 1][j + 1]);
            a += gx * gx;
            b += gx * gy;
            c += gy * gy;
        }
    *oa = a;
    *ob = b;
    *oc = c;
}

// ---- fixed-point Gaussian (cv::GaussianBlur 9x9 sigma 2.5 on CV_8U, OpenCV 4.x) ------------------
// taps g = {12,22,31,41,44,41,31,22,12} (8 fractional bits, sum 256)
#define YAVO_G0 12u
#define YAVO_G1 22u
#define YAVO_G2 31u
#define YAVO_G3 41u
#define YAVO_G4 44u
#define YAVO_PK(a, b, c, d) ((a) | ((b) << 8) | ((c) << 16) | ((d) << 24))

// horizontal pass for 4 adjacent outputs x..x+3 from the three words covering x-4..x+7:
// out[i] = sum_k g[k] * p[x+i+k-4]; three dp4a per output with pre-shifted weight vectors
#ifdef __CUDACC__
// the twelve weight words of the horizontal pass in constant memory: IDP.4A takes them as constant-bank operands, so no
// instruction is spent on materialising them (as immediates they cost a UMOV each, re-issued in every loop iteration)
__constant__ uint32_t yavo_hw[12] = {
    YAVO_PK(YAVO_G0, YAVO_G1, YAVO_G2, YAVO_G3), YAVO_PK(YAVO_G4, YAVO_G3, YAVO_G2, YAVO_G1), YAVO_PK(YAVO_G0, 0u, 0u, 0u),
    YAVO_PK(0u, YAVO_G0, YAVO_G1, YAVO_G2),      YAVO_PK(YAVO_G3, YAVO_G4, YAVO_G3, YAVO_G2), YAVO_PK(YAVO_G1, YAVO_G0, 0u, 0u),
    YAVO_PK(0u, 0u, YAVO_G0, YAVO_G1),           YAVO_PK(YAVO_G2, YAVO_G3, YAVO_G4, YAVO_G3), YAVO_PK(YAVO_G2, YAVO_G1, YAVO_G0, 0u),
    YAVO_PK(0u, 0u, 0u, YAVO_G0),                YAVO_PK(YAVO_G1, YAVO_G2, YAVO_G3, YAVO_G4), YAVO_PK(YAVO_G3, YAVO_G2, YAVO_G1, YAVO_G0)};
#endif
YAVO_HD void yavo_blur_h4(uint32_t wm, uint32_t w0, uint32_t wp, uint32_t out[4]) {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int i = 0; i < 4; i++)
        out[i] = yavo_dp4a(wm, yavo_hw[3 * i], yavo_dp4a(w0, yavo_hw[3 * i + 1], yavo_dp4a(wp, yavo_hw[3 * i + 2], 0u)));
#else
    out[0] = yavo_dp4a(wm, YAVO_PK(YAVO_G0, YAVO_G1, YAVO_G2, YAVO_G3),
             yavo_dp4a(w0, YAVO_PK(YAVO_G4, YAVO_G3, YAVO_G2, YAVO_G1),
             yavo_dp4a(wp, YAVO_PK(YAVO_G0, 0u, 0u, 0u), 0u)));
    out[1] = yavo_dp4a(wm, YAVO_PK(0u, YAVO_G0, YAVO_G1, YAVO_G2),
             yavo_dp4a(w0, YAVO_PK(YAVO_G3, YAVO_G4, YAVO_G3, YAVO_G2),
             yavo_dp4a(wp, YAVO_PK(YAVO_G1, YAVO_G0, 0u, 0u), 0u)));
    out[2] = yavo_dp4a(wm, YAVO_PK(0u, 0u, YAVO_G0, YAVO_G1),
             yavo_dp4a(w0, YAVO_PK(YAVO_G2, YAVO_G3, YAVO_G4, YAVO_G3),
             yavo_dp4a(wp, YAVO_PK(YAVO_G2, YAVO_G1, YAVO_G0, 0u), 0u)));
    out[3] = yavo_dp4a(wm, YAVO_PK(0u, 0u, 0u, YAVO_G0),
             yavo_dp4a(w0, YAVO_PK(YAVO_G1, YAVO_G2, YAVO_G3, YAVO_G4),
             yavo_dp4a(wp, YAVO_PK(YAVO_G3, YAVO_G2, YAVO_G1, YAVO_G0), 0u)));
#endif
}

// vertical pass for two vertically adjacent outputs (rows r, r+1; r even in tile coordinates)
// from five vertical pairs P[i] = (h[r-4+2i] | h[r-3+2i] << 16):
//   out(r)   = g0 h[r-4] + g1 h[r-3] + ... + g8 h[r+4]
//   out(r+1) = g0 h[r-3] + ... + g8 h[r+5]
// raw form: the two fixed-point sums with the rounding constant already added; the blurred pixels are
// bits 16..23 of each (sum < 2^24), which the kernel extracts with byte permutes while packing four outputs
#ifdef __CUDACC__
__constant__ uint32_t yavo_vw[5] = {YAVO_PK(YAVO_G0, YAVO_G1, YAVO_G2, YAVO_G3), YAVO_PK(YAVO_G4, YAVO_G3, YAVO_G2, YAVO_G1),
                                    YAVO_PK(YAVO_G0, 0u, 0u, YAVO_G0), YAVO_PK(YAVO_G1, YAVO_G2, YAVO_G3, YAVO_G4),
                                    YAVO_PK(YAVO_G3, YAVO_G2, YAVO_G1, YAVO_G0)};  // constant-bank operands, as yavo_hw
#endif
YAVO_HD void yavo_blur_v2_raw(const uint32_t P[5], uint32_t *ra, uint32_t *rb) {
#ifdef __CUDA_ARCH__
    const uint32_t wA = yavo_vw[0], wB = yavo_vw[1], wC = yavo_vw[2], wD = yavo_vw[3], wE = yavo_vw[4];
#else
    const uint32_t wA = YAVO_PK(YAVO_G0, YAVO_G1, YAVO_G2, YAVO_G3);  // lo: g0,g1  hi: g2,g3
    const uint32_t wB = YAVO_PK(YAVO_G4, YAVO_G3, YAVO_G2, YAVO_G1);  // lo: g4,g5  hi: g6,g7
    const uint32_t wC = YAVO_PK(YAVO_G0, 0u, 0u, YAVO_G0);            // lo: g8,0   hi: 0,g0
    const uint32_t wD = YAVO_PK(YAVO_G1, YAVO_G2, YAVO_G3, YAVO_G4);  // lo: g1,g2  hi: g3,g4
    const uint32_t wE = YAVO_PK(YAVO_G3, YAVO_G2, YAVO_G1, YAVO_G0);  // lo: g5,g6  hi: g7,g8
#endif
    uint32_t a = yavo_dp2a_lo(P[0], wA, 32768u);
    a = yavo_dp2a_hi(P[1], wA, a);
    a = yavo_dp2a_lo(P[2], wB, a);
    a = yavo_dp2a_hi(P[3], wB, a);
    a = yavo_dp2a_lo(P[4], wC, a);
    uint32_t b = yavo_dp2a_hi(P[0], wC, 32768u);
    b = yavo_dp2a_lo(P[1], wD, b);
    b = yavo_dp2a_hi(P[2], wD, b);
    b = yavo_dp2a_lo(P[3], wE, b);
    b = yavo_dp2a_hi(P[4], wE, b);
    *ra = a;
    *rb = b;
}

YAVO_HD void yavo_blur_v2(const uint32_t P[5], uint32_t *o0, uint32_t *o1) {
    uint32_t a, b;
    yavo_blur_v2_raw(P, &a, &b);
    *o0 = a >> 16;
    *o1 = b >> 16;
}

#endif  // YAVO_FAST_CORE_H
