// klt_kernels.cuh — sm_100a kernels for the tracking step that follows the feature front end in the
// reference's VO loop:
//     cv::calcOpticalFlowPyrLK(last, cur, lastKpt, curKpt, status, error, Size(11,11), 3,
//                              TermCriteria(COUNT+EPS, 30, 0.01), 0, 0.001)          src/LoopHandler.cc:372-375
// K7 pyr_down_kernel   one pyramid level from the one below (cv::pyrDown: [1 4 6 4 1]^2, REFLECT_101,
//                      (sum + 128) >> 8), a batch of frame slots per launch — the only image pyramid the
//                      reference builds (inside OpenCV).
// K8 klt_track_kernel  one warp per point, all pyramid levels inside the kernel.  The Scharr derivatives are
//                      computed on the fly from a staged raw patch of the previous frame (they are needed
//                      only under tracked points, so no derivative planes are written to HBM); bilinear
//                      weights, descaling and the 2x2 solve follow OpenCV's fixed-point / float32 sequence,
//                      and the float sums are accumulated in the ORDER of OpenCV's 128-bit SIMD loops (lane
//                      accumulators over groups of 8 columns + a scalar tail), which makes positions, status
//                      and err bit-identical to cv2 4.13 (tests/golden/klt_golden.npz).
//                      klt_track_fixed_kernel<11,11> is the same algorithm with the reference's window size
//                      known at compile time; klt_track_kernel takes any window up to 31 x 31.
// K9 epipolar_inliers_kernel  inlier count of the fundamental-matrix RANSAC (src/3DHandler.cc:163-190).
// Float arithmetic uses the _rn intrinsics: nothing may be contracted to FMA or reassociated.
#pragma once
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

namespace yavo {

constexpr int KLT_MAX_LEVELS = 7;   // pyramid levels above level 0 a context can hold
constexpr int KLT_MAX_WIN = 31;     // largest window side
constexpr int KLT_WARPS = 4;        // points per CTA

struct KltLevels {
    const uint8_t *img[KLT_MAX_LEVELS + 1];   // level l of slot 0 (level 0 = the frame slots themselves)
    unsigned long long slot_stride[KLT_MAX_LEVELS + 1];
    int pitch[KLT_MAX_LEVELS + 1], H[KLT_MAX_LEVELS + 1], W[KLT_MAX_LEVELS + 1];
    int top;                                   // highest level used
};

struct KltParams {
    int ww, wh, max_count, flags;
    double eps2;        // criteria.epsilon squared (double, as OpenCV compares delta.ddot(delta))
    float min_eig;
};

__device__ __forceinline__ int klt_refl(int p, int n) {  // BORDER_REFLECT_101
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
}

// ------------------------------------------------------------------------------------------------
// K7  cv::pyrDown on 8-bit frames: separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8.
// One CTA = a 64 x 16 output tile.  The 131 x 35 input region is staged in shared memory (aligned 128-bit loads for
// tiles whose columns lie inside the frame, reflected byte loads at the left / right edge; rows are reflected per
// row either way); the horizontal pass turns four staged words into four sums with two PRMT and four IDP.4A
// (taps 1 4 6 4 in the dot product, the fifth tap added), stored transposed as u16 so that the vertical pass reads
// row pairs as words: two IDP.2A (taps 1 4 | 6 4, rounding constant in the accumulator) + the fifth row; four
// outputs are packed into one 32-bit store.  HBM traffic is the algorithmic one: every input byte read once
// (tile halos overlap by 3 of 131 columns / 35 rows), every output byte written once.
// grid (ceil(oW / 64), ceil(oH / 16), slots)
// ------------------------------------------------------------------------------------------------
constexpr int PD_TW = 64, PD_TH = 16;                 // output tile
constexpr int PD_IW = 160, PD_IH = 2 * PD_TH + 3;     // staged bytes per row (cols 2x0-16 .. 2x0+143, 16-byte aligned), rows (35)
constexpr int PD_HS = 38;                             // u16 per transposed column of horizontal sums (19 words: conflict-free)

__global__ void __launch_bounds__(256)
pyr_down_kernel(const uint8_t *__restrict__ src, size_t src_stride, int src_pitch, int sH, int sW,
                uint8_t *__restrict__ dst, size_t dst_stride, int dst_pitch, int oH, int oW) {
    __shared__ __align__(16) uint32_t tile[PD_IH][PD_IW / 4];
    __shared__ __align__(4) uint16_t hT[PD_TW][PD_HS];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * PD_TW, y0 = blockIdx.y * PD_TH;
    const uint8_t *s = src + (size_t)blockIdx.z * src_stride;
    const int cbase = 2 * x0 - 16;  // frame column of staged byte 0 (a multiple of 16)
    for (int i = tid; i < PD_IH * (PD_IW / 16); i += 256) {   // 350 x 16 bytes, rows reflected
        const int r = i / (PD_IW / 16), q = i - r * (PD_IW / 16);
        const int gr = klt_refl(2 * y0 - 2 + r, sH), c = cbase + 16 * q;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (c >= 0 && c + 16 <= src_pitch) v = __ldg(reinterpret_cast<const uint4 *>(s + (size_t)gr * src_pitch + c));
        reinterpret_cast<uint4 *>(&tile[r][0])[q] = v;  // columns >= sW hold row padding here; the ones a kept output
    }                                                   // can touch are replaced below
    if (cbase < 0 || cbase + PD_IW > sW) {
        // BORDER_REFLECT_101 columns: -2, -1 on the left; sW, sW+1, sW+2 on the right (outputs x < oW reach no further)
        __syncthreads();
        uint8_t *tb = reinterpret_cast<uint8_t *>(&tile[0][0]);
        for (int i = tid; i < PD_IH * 5; i += 256) {
            const int r = i / 5, j = i - r * 5;
            const int c = j < 2 ? j - 2 : sW + j - 2, sc = c - cbase;
            if (sc < 0 || sc >= PD_IW) continue;
            const int gr = klt_refl(2 * y0 - 2 + r, sH);
            tb[r * PD_IW + sc] = __ldg(s + (size_t)gr * src_pitch + klt_refl(c, sW));
        }
    }
    __syncthreads();
    // horizontal: item = (staged row r, four outputs 4g .. 4g+3); output x's window starts at staged byte 2x + 14
    for (int i = tid; i < PD_IH * (PD_TW / 4); i += 256) {
        const int r = i >> 4, g = i & 15;
        const uint32_t w3 = tile[r][2 * g + 3], w6 = tile[r][2 * g + 6];
        const uint2 w45 = *reinterpret_cast<const uint2 *>(&tile[r][2 * g + 4]);
        const uint32_t T = 0x04060401u;  // taps 1 4 6 4 on four consecutive bytes; the fifth tap (1) is added
        const uint32_t h0 = __dp4a(__byte_perm(w3, w45.x, 0x5432), T, (w45.x >> 16) & 0xffu);
        const uint32_t h1 = __dp4a(w45.x, T, w45.y & 0xffu);
        const uint32_t h2 = __dp4a(__byte_perm(w45.x, w45.y, 0x5432), T, (w45.y >> 16) & 0xffu);
        const uint32_t h3 = __dp4a(w45.y, T, w6 & 0xffu);
        hT[4 * g][r] = (uint16_t)h0;
        hT[4 * g + 1][r] = (uint16_t)h1;
        hT[4 * g + 2][r] = (uint16_t)h2;
        hT[4 * g + 3][r] = (uint16_t)h3;
    }
    __syncthreads();
    // vertical: thread = (output row y, four consecutive output columns)
    {
        const int y = tid >> 4, xq = (tid & 15) * 4;
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const uint32_t *col = reinterpret_cast<const uint32_t *>(&hT[xq + q][0]) + y;   // u16 index 2y
            uint32_t v = __dp2a_lo(col[0], 0x0401u, 128u);  // h[2y] + 4 h[2y+1] + rounding
            v = __dp2a_lo(col[1], 0x0406u, v);              // 6 h[2y+2] + 4 h[2y+3]
            v += col[2] & 0xffffu;                          // h[2y+4]
            o[q] = v >> 8;
        }
        const int gy = y0 + y, gx = x0 + xq;
        if (gy < oH && gx < oW)  // dst_pitch is a multiple of 16: the word may run past oW into the row padding
            *reinterpret_cast<uint32_t *>(dst + (size_t)blockIdx.z * dst_stride + (size_t)gy * dst_pitch + gx) =
                o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
    }
}

// ------------------------------------------------------------------------------------------------
// K8
// ------------------------------------------------------------------------------------------------
struct KltW {
    int w00, w01, w10, w11;
};
__device__ __forceinline__ KltW klt_weights(float a, float b) {  // 14-bit bilinear weights, cvRound = half to even
    const float oa = __fsub_rn(1.f, a), ob = __fsub_rn(1.f, b);
    KltW w;
    w.w00 = __float2int_rn(__fmul_rn(__fmul_rn(oa, ob), 16384.f));
    w.w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, ob), 16384.f));
    w.w10 = __float2int_rn(__fmul_rn(__fmul_rn(oa, b), 16384.f));
    w.w11 = 16384 - w.w00 - w.w01 - w.w10;
    return w;
}

// Ordered float accumulation as OpenCV's 128-bit SIMD loops do it.  Columns are taken in groups of 8: accumulator
// k (0..3) owns columns k and k+4 of every group; the columns past the last full group go to a scalar "tail"
// accumulator in scan order.  The result is tail + ((q0 + q2) + (q1 + q3)) (v_reduce_sum).
// Gradient matrix: the terms are float products (exact: |gradient| <= 4080), added one by one.
__device__ __forceinline__ float klt_chain_f(const float *__restrict__ T, int ww, int wh, int role) {
    const int vec = ww & ~7;
    float acc = 0.f;
    if (role < 4) {
#pragma unroll
        for (int y = 0; y < wh; y++) {
#pragma unroll
            for (int g = 0; g < vec; g += 8) {
                acc = __fadd_rn(T[y * ww + g + role], acc);
                acc = __fadd_rn(T[y * ww + g + role + 4], acc);
            }
        }
    } else {
#pragma unroll
        for (int y = 0; y < wh; y++) {
#pragma unroll
            for (int x = vec; x < ww; x++) acc = __fadd_rn(acc, T[y * ww + x]);
        }
    }
    return acc;
}
// Mismatch vector: the terms arrive as one contiguous float list per accumulator, already in accumulation order
// (pair sums of columns (x, x+4) added as int32 and converted, tail products converted one by one), so a chain is
// 128-bit loads plus dependent adds.  `cnt` terms, list 16-byte aligned, padding never added.
__device__ __forceinline__ float klt_chain_list(const float *__restrict__ F, int cnt) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < cnt; i += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(F + i);
        acc = __fadd_rn(acc, v.x);
        if (i + 1 < cnt) acc = __fadd_rn(acc, v.y);
        if (i + 2 < cnt) acc = __fadd_rn(acc, v.z);
        if (i + 3 < cnt) acc = __fadd_rn(acc, v.w);
    }
    return acc;
}

// tail + ((q0 + q2) + (q1 + q3)) with q_k in lanes base..base+3 and the tail in lane tail_lane
__device__ __forceinline__ float klt_combine(float v, int base, int tail_lane) {
    const float q0 = __shfl_sync(0xffffffffu, v, base), q1 = __shfl_sync(0xffffffffu, v, base + 1);
    const float q2 = __shfl_sync(0xffffffffu, v, base + 2), q3 = __shfl_sync(0xffffffffu, v, base + 3);
    const float t = __shfl_sync(0xffffffffu, v, tail_lane);
    return __fadd_rn(t, __fadd_rn(__fadd_rn(q0, q2), __fadd_rn(q1, q3)));
}

// per-warp shared memory: I / Ix / Iy window samples (int16); a region holding the three float term planes of the
// gradient matrix and later the term lists of the mismatch vector; the staged raw patch (u8, reused for the patches
// of the next frame); the derivative patch (short2)
// term lists of the mismatch vector: 8 lane accumulators (b1: 0-3, b2: 4-7) of LV floats, then two tails of LT
__host__ __device__ __forceinline__ int klt_list_v(int ww, int wh) { return (wh * (ww >> 3) + 3) & ~3; }
__host__ __device__ __forceinline__ int klt_list_t(int ww, int wh) { return (wh * (ww & 7) + 3) & ~3; }

__host__ __device__ __forceinline__ size_t klt_smem_per_warp(int ww, int wh) {
    const size_t area = (size_t)ww * wh;
    size_t b = (3 * ((area * 2 + 3) & ~size_t(3)) + 15) & ~size_t(15);
    const size_t lists = 4 * (size_t)(8 * klt_list_v(ww, wh) + 2 * klt_list_t(ww, wh));
    b += ((12 * area > lists ? 12 * area : lists) + 15) & ~size_t(15);
    b += ((size_t)(ww + 3) * (wh + 3) + 3) & ~size_t(3);
    b += (size_t)(ww + 1) * (wh + 1) * 4;
    return (b + 15) & ~size_t(15);
}

// Staged ph x pw patch of image `im` with origin (oy, ox), REFLECT_101 outside the image.
__device__ __forceinline__ void klt_stage_patch(uint8_t *dst, const uint8_t *__restrict__ im, int pitch, int H, int W, int oy,
                                                int ox, int ph, int pw, int lane) {
    if (oy >= 0 && ox >= 0 && oy + ph <= H && ox + pw <= W) {
        const uint8_t *src = im + (size_t)oy * pitch + ox;
#pragma unroll
        for (int p = lane; p < ph * pw; p += 32) {
            const int r = p / pw, c = p - r * pw;
            dst[p] = __ldg(src + r * pitch + c);
        }
    } else {
#pragma unroll
        for (int p = lane; p < ph * pw; p += 32) {
            const int r = p / pw, c = p - r * pw;
            dst[p] = __ldg(im + (size_t)klt_refl(oy + r, H) * pitch + klt_refl(ox + c, W));
        }
    }
}

// bilinear sample of the staged next-frame patch minus the stored previous-frame sample (5 fractional bits)
__device__ __forceinline__ int klt_diff(const uint8_t *q, int gw, const KltW &w, int iv) {
    return ((q[0] * w.w00 + q[1] * w.w01 + q[gw] * w.w10 + q[gw + 1] * w.w11 + (1 << 8)) >> 9) - iv;
}

// K8, any window up to KLT_MAX_WIN x KLT_MAX_WIN (window size taken from P at run time).  The reference's 11 x 11
// window runs klt_track_fixed_kernel below instead.
#ifndef YAVO_KLT_MIN_CTAS
#define YAVO_KLT_MIN_CTAS 5
#endif
__global__ void __launch_bounds__(KLT_WARPS * 32, YAVO_KLT_MIN_CTAS)
klt_track_kernel(KltLevels L, KltParams P, int prev_slot0, int next_slot0,
                 const float2 *__restrict__ prev_xy,                                  // explicit points, or
                 const int32_t *__restrict__ kp_row, const int32_t *__restrict__ kp_col,  // keypoints (row, col) per slot
                 const int *__restrict__ n_all, int n_fixed, int pts_stride,
                 const float2 *__restrict__ init_xy, float2 *__restrict__ next_xy, uint8_t *__restrict__ status,
                 float *__restrict__ err) {
    extern __shared__ __align__(16) uint8_t klt_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = blockIdx.y;
    const int pt = blockIdx.x * KLT_WARPS + warp;
    const int n = n_all ? n_all[prev_slot0 + pair] : n_fixed;
    if (pt >= n) return;
    const int ww = P.ww, wh = P.wh, area = ww * wh;
    const int rw = ww + 3, gw = ww + 1;
    const size_t plane = ((size_t)area * 2 + 3) & ~size_t(3);
    uint8_t *base = klt_smem + (size_t)warp * klt_smem_per_warp(ww, wh);
    int16_t *sI = reinterpret_cast<int16_t *>(base), *sIx = reinterpret_cast<int16_t *>(base + plane),
            *sIy = reinterpret_cast<int16_t *>(base + 2 * plane);
    const size_t planes = (3 * plane + 15) & ~size_t(15);
    float *sT = reinterpret_cast<float *>(base + planes);      // T11 | T12 | T22, later the mismatch term lists (16-byte aligned)
    const int G = ww >> 3, vec = G * 8, tw = ww - vec;          // column groups of 8, tail columns
    const int LV = klt_list_v(ww, wh), LT = klt_list_t(ww, wh);
    const int NVI = wh * G * 4, NTI = wh * tw;                  // pair items, tail items of one iteration
    const size_t lists = 4 * (size_t)(8 * LV + 2 * LT);
    uint8_t *sR = base + planes + (((12 * (size_t)area > lists ? 12 * (size_t)area : lists) + 15) & ~size_t(15));
    short2 *sG = reinterpret_cast<short2 *>(sR + (((size_t)rw * (wh + 3) + 3) & ~size_t(3)));

    const size_t o = (size_t)pair * pts_stride + pt;
    float2 p0;
    if (prev_xy) p0 = prev_xy[o];
    else p0 = make_float2((float)kp_col[(size_t)(prev_slot0 + pair) * pts_stride + pt],
                          (float)kp_row[(size_t)(prev_slot0 + pair) * pts_stride + pt]);
    const float halfx = __fmul_rn((float)(ww - 1), 0.5f), halfy = __fmul_rn((float)(wh - 1), 0.5f);
    const bool use_initial = (P.flags & 4) != 0, get_min_eig = (P.flags & 8) != 0;
    const float FLT_SCALE = 1.f / (1 << 20);
    float outx = 0.f, outy = 0.f, ev = 0.f;
    bool st = true;

    for (int level = L.top; level >= 0; level--) {
        const int H = L.H[level], W = L.W[level], pitch = L.pitch[level];
        const uint8_t *I = L.img[level] + (size_t)(prev_slot0 + pair) * L.slot_stride[level];
        const uint8_t *J = L.img[level] + (size_t)(next_slot0 + pair) * L.slot_stride[level];
        const float sc = __int_as_float((127 - level) << 23);  // 1 / (1 << level)
        float px = __fmul_rn(p0.x, sc), py = __fmul_rn(p0.y, sc);
        float nx, ny;
        if (level == L.top) {
            if (use_initial) {
                const float2 q = init_xy[o];
                nx = __fmul_rn(q.x, sc);
                ny = __fmul_rn(q.y, sc);
            } else {
                nx = px;
                ny = py;
            }
        } else {
            nx = __fmul_rn(outx, 2.f);
            ny = __fmul_rn(outy, 2.f);
        }
        outx = nx;
        outy = ny;
        px = __fsub_rn(px, halfx);
        py = __fsub_rn(py, halfy);
        const int ix = __float2int_rd(px), iy = __float2int_rd(py);
        if (ix < -ww || ix >= W || iy < -wh || iy >= H) {
            if (level == 0) {
                st = false;
                ev = 0.f;
            }
            continue;
        }
        KltW w = klt_weights(__fsub_rn(px, (float)ix), __fsub_rn(py, (float)iy));

        // ---- raw patch rows iy-1 .. iy+wh+1, cols ix-1 .. ix+ww+1 of the previous frame ----------------
        __syncwarp();
        klt_stage_patch(sR, I, pitch, H, W, iy - 1, ix - 1, wh + 3, rw, lane);
        __syncwarp();
        // Scharr (dx, dy) at window positions (iy + r, ix + c), r <= wh, c <= ww; zero outside the image
        const bool all_in = iy >= 0 && ix >= 0 && iy + wh < H && ix + ww < W;
#pragma unroll
        for (int p = lane; p < (wh + 1) * gw; p += 32) {
            const int r = p / gw, c = p - r * gw;
            short2 g = make_short2(0, 0);
            if (all_in || (iy + r >= 0 && iy + r < H && ix + c >= 0 && ix + c < W)) {
                const uint8_t *q0 = sR + r * rw + c, *q1 = q0 + rw, *q2 = q1 + rw;
                const int t0a = (q0[0] + q2[0]) * 3 + q1[0] * 10, t0c = (q0[2] + q2[2]) * 3 + q1[2] * 10;
                const int t1a = q2[0] - q0[0], t1b = q2[1] - q0[1], t1c = q2[2] - q0[2];
                g.x = (short)(t0c - t0a);
                g.y = (short)((t1c + t1a) * 3 + t1b * 10);
            }
            sG[p] = g;
        }
        __syncwarp();
        // bilinear samples of the patch and its derivatives; float terms of the gradient matrix
#pragma unroll
        for (int p = lane; p < area; p += 32) {
            const int y = p / ww, x = p - y * ww;
            const uint8_t *q = sR + (y + 1) * rw + (x + 1);
            const int iv = (q[0] * w.w00 + q[1] * w.w01 + q[rw] * w.w10 + q[rw + 1] * w.w11 + (1 << 8)) >> 9;
            const short2 g00 = sG[y * gw + x], g01 = sG[y * gw + x + 1], g10 = sG[(y + 1) * gw + x], g11 = sG[(y + 1) * gw + x + 1];
            const int gx = (g00.x * w.w00 + g01.x * w.w01 + g10.x * w.w10 + g11.x * w.w11 + (1 << 13)) >> 14;
            const int gy = (g00.y * w.w00 + g01.y * w.w01 + g10.y * w.w10 + g11.y * w.w11 + (1 << 13)) >> 14;
            sI[p] = (int16_t)iv;
            sIx[p] = (int16_t)gx;
            sIy[p] = (int16_t)gy;
            const float fx = (float)gx, fy = (float)gy;
            sT[p] = __fmul_rn(fx, fx);
            sT[area + p] = __fmul_rn(fx, fy);
            sT[2 * area + p] = __fmul_rn(fy, fy);
        }
        __syncwarp();
        // gradient matrix: lanes 0-3 A11, 4-7 A12, 8-11 A22 lane accumulators; 12-14 their tails
        float v = 0.f;
        if (lane < 12) v = klt_chain_f(sT + (lane >> 2) * area, ww, wh, lane & 3);
        else if (lane < 15) v = klt_chain_f(sT + (lane - 12) * area, ww, wh, 4);
        const float A11 = __fmul_rn(klt_combine(v, 0, 12), FLT_SCALE), A12 = __fmul_rn(klt_combine(v, 4, 13), FLT_SCALE),
                    A22 = __fmul_rn(klt_combine(v, 8, 14), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dd = __fsub_rn(A11, A22);
        const float rad = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * ww * wh));
        if (get_min_eig) ev = minEig;
        if (minEig < P.min_eig || D < 1.1920928955078125e-07f) {
            if (level == 0) st = false;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, halfx);
        ny = __fsub_rn(ny, halfy);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < P.max_count; j++) {
            const int jx = __float2int_rd(nx), jy = __float2int_rd(ny);
            if (jx < -ww || jx >= W || jy < -wh || jy >= H) {
                if (level == 0) st = false;
                break;
            }
            w = klt_weights(__fsub_rn(nx, (float)jx), __fsub_rn(ny, (float)jy));
            __syncwarp();
            klt_stage_patch(sR, J, pitch, H, W, jy, jx, wh + 1, gw, lane);
            __syncwarp();
            // work items in accumulation order: pair items (row y, group g, k < 4) = columns 8g+k and 8g+k+4 of one
            // row, then the tail pixels; each writes its finished float term into its accumulator's list
#pragma unroll
            for (int i = lane; i < NVI + NTI; i += 32) {
                if (i < NVI) {
                    const int y = i / (4 * G), rem = i - y * 4 * G, g = rem >> 2, k = rem & 3;
                    const int pa = y * ww + 8 * g + k, pb = pa + 4;
                    const int da = klt_diff(sR + y * gw + 8 * g + k, gw, w, sI[pa]);
                    const int db = klt_diff(sR + y * gw + 8 * g + k + 4, gw, w, sI[pb]);
                    sT[k * LV + y * G + g] = __int2float_rn(da * sIx[pa] + db * sIx[pb]);
                    sT[(4 + k) * LV + y * G + g] = __int2float_rn(da * sIy[pa] + db * sIy[pb]);
                } else {
                    const int t = i - NVI, y = t / tw, x = vec + t - y * tw, pa = y * ww + x;
                    const int da = klt_diff(sR + y * gw + x, gw, w, sI[pa]);
                    sT[8 * LV + t] = __int2float_rn(da * sIx[pa]);
                    sT[8 * LV + LT + t] = __int2float_rn(da * sIy[pa]);
                }
            }
            __syncwarp();
            float u = 0.f;
            if (lane < 8) u = klt_chain_list(sT + lane * LV, wh * G);
            else if (lane < 10) u = klt_chain_list(sT + 8 * LV + (lane - 8) * LT, NTI);
            const float b1 = __fmul_rn(klt_combine(u, 0, 8), FLT_SCALE), b2 = __fmul_rn(klt_combine(u, 4, 9), FLT_SCALE);
            const float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, ddx);
            ny = __fadd_rn(ny, ddy);
            outx = __fadd_rn(nx, halfx);
            outy = __fadd_rn(ny, halfy);
            if (__dadd_rn(__dmul_rn((double)ddx, (double)ddx), __dmul_rn((double)ddy, (double)ddy)) <= P.eps2) break;
            if (j > 0 && fabs((double)__fadd_rn(ddx, pdx)) < 0.01 && fabs((double)__fadd_rn(ddy, pdy)) < 0.01) {
                outx = __fsub_rn(outx, __fmul_rn(ddx, 0.5f));
                outy = __fsub_rn(outy, __fmul_rn(ddy, 0.5f));
                break;
            }
            pdx = ddx;
            pdy = ddy;
        }
        if (st && level == 0 && !get_min_eig) {
            // err = mean absolute difference of the patch at the final position (5 fractional bits removed)
            const float qx = __fsub_rn(outx, halfx), qy = __fsub_rn(outy, halfy);
            const int jx = __float2int_rd(qx), jy = __float2int_rd(qy);
            if (jx < -ww || jx >= W || jy < -wh || jy >= H) {
                st = false;
                continue;
            }
            w = klt_weights(__fsub_rn(qx, (float)jx), __fsub_rn(qy, (float)jy));
            __syncwarp();
            klt_stage_patch(sR, J, pitch, H, W, jy, jx, wh + 1, gw, lane);
            __syncwarp();
            int e = 0;
#pragma unroll
            for (int p = lane; p < area; p += 32) {
                const int y = p / ww, x = p - y * ww;
                e += abs(klt_diff(sR + y * gw + x, gw, w, sI[p]));
            }
#pragma unroll
            for (int s = 16; s; s >>= 1) e += __shfl_xor_sync(0xffffffffu, e, s);
            ev = __fdiv_rn(__fmul_rn((float)e, 1.f), (float)(32 * ww * wh));  // integer partial sums: order-free
        }
    }
    if (lane == 0) {
        next_xy[o] = make_float2(outx, outy);
        status[o] = st ? 1 : 0;
        err[o] = ev;
    }
}

// ------------------------------------------------------------------------------------------------
// K8, compile-time window (the reference's 11 x 11).  Same arithmetic as klt_track_kernel; what changes is where the
// per-step work lives:
//   * patches are staged as aligned 32-bit words (2-3 loads per lane instead of 5-7 byte loads); a byte shift
//     (origin column & 3) locates the pixels inside the staged rows;
//   * every lane owns up to NSLOT fixed work items of a Newton step — a pair item (columns x and x+4 of one row, whose
//     int32 products OpenCV adds before converting to float) or a tail pixel — with their staged-patch offsets,
//     term-list positions and the previous frame's I / Ix / Iy samples held in registers for the whole level;
//   * the ten accumulation chains read their terms as 128-bit loads from contiguous lists; the first CNTV terms are
//     added by all chain lanes in one instruction stream, only the two tails continue.
// ------------------------------------------------------------------------------------------------
template <int WW, int WH>
struct KltFixed {
    static constexpr int AREA = WW * WH, G = WW / 8, VEC = G * 8, TW = WW - VEC;
    static constexpr int CNTV = WH * G, CNTT = WH * TW;                    // terms per lane accumulator / per tail
    static constexpr int LV = (CNTV + 3) & ~3, LT = (CNTT + 3) & ~3;       // padded list lengths (floats)
    static constexpr int NVI = CNTV * 4, NTI = CNTT, NIT = NVI + NTI, NSLOT = (NIT + 31) / 32;
    static constexpr int JW = (WW + 1 + 3 + 3) / 4, RS = 4 * JW;           // next-frame patch: words / bytes per staged row
    static constexpr int IW = (WW + 3 + 3 + 3) / 4, RS1 = 4 * IW;          // raw previous-frame patch
    static constexpr int PLANE = (AREA * 2 + 3) & ~3, PLANES = (3 * PLANE + 15) & ~15;
    static constexpr int LISTS = 4 * (8 * LV + 2 * LT);
    static constexpr int TERMS = ((12 * AREA > LISTS + 8 ? 12 * AREA : LISTS + 8) + 15) & ~15;   // + 2 scratch floats
    static constexpr int PATCH = (((WH + 3) * RS1 > (WH + 1) * RS ? (WH + 3) * RS1 : (WH + 1) * RS) + 15) & ~15;
    static constexpr int DERIV = (WW + 1) * (WH + 1) * 4;
    static constexpr int SMEM = (PLANES + TERMS + PATCH + DERIV + 15) & ~15;
    static_assert(CNTT >= CNTV && G >= 1, "window shape not covered by the fixed kernel");
};

// ROWS x WORDS aligned words starting at row oy, byte column c0 (multiple of 4); all rows inside the image
template <int ROWS, int WORDS>
__device__ __forceinline__ void klt_stage_words(uint32_t *dst, const uint8_t *__restrict__ im, int pitch, int oy, int c0, int lane) {
    const uint8_t *src = im + (size_t)oy * pitch + c0;
#pragma unroll
    for (int p = lane; p < ROWS * WORDS; p += 32) {
        const int r = p / WORDS, wd = p - r * WORDS;
        dst[p] = (c0 + 4 * wd < pitch) ? __ldg(reinterpret_cast<const uint32_t *>(src + r * pitch + 4 * wd)) : 0u;
    }
}
// the same patch through REFLECT_101 (window reaches outside the image), byte by byte, shift 0
template <int ROWS, int COLS, int RSB>
__device__ __forceinline__ void klt_stage_reflect(uint8_t *dst, const uint8_t *__restrict__ im, int pitch, int H, int W, int oy, int ox,
                                                  int lane) {
#pragma unroll
    for (int p = lane; p < ROWS * COLS; p += 32) {
        const int r = p / COLS, c = p - r * COLS;
        dst[r * RSB + c] = __ldg(im + (size_t)klt_refl(oy + r, H) * pitch + klt_refl(ox + c, W));
    }
}

template <int RSB>
__device__ __forceinline__ int klt_bilin(const uint8_t *q, const KltW &w) {
    return (q[0] * w.w00 + q[1] * w.w01 + q[RSB] * w.w10 + q[RSB + 1] * w.w11 + (1 << 8)) >> 9;
}

template <int WW, int WH>
__global__ void __launch_bounds__(KLT_WARPS * 32, YAVO_KLT_MIN_CTAS)
klt_track_fixed_kernel(KltLevels L, KltParams P, int prev_slot0, int next_slot0, const float2 *__restrict__ prev_xy,
                       const int32_t *__restrict__ kp_row, const int32_t *__restrict__ kp_col, const int *__restrict__ n_all,
                       int n_fixed, int pts_stride, const float2 *__restrict__ init_xy, float2 *__restrict__ next_xy,
                       uint8_t *__restrict__ status, float *__restrict__ err) {
    using F = KltFixed<WW, WH>;
    extern __shared__ __align__(16) uint8_t klt_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = blockIdx.y;
    const int pt = blockIdx.x * KLT_WARPS + warp;
    const int n = n_all ? n_all[prev_slot0 + pair] : n_fixed;
    if (pt >= n) return;
    constexpr int ww = WW, wh = WH, area = F::AREA, gw = WW + 1;
    uint8_t *base = klt_smem + (size_t)warp * F::SMEM;
    int16_t *sI = reinterpret_cast<int16_t *>(base), *sIx = reinterpret_cast<int16_t *>(base + F::PLANE),
            *sIy = reinterpret_cast<int16_t *>(base + 2 * F::PLANE);
    float *sT = reinterpret_cast<float *>(base + F::PLANES);
    uint8_t *sR = base + F::PLANES + F::TERMS;
    short2 *sG = reinterpret_cast<short2 *>(sR + F::PATCH);

    // ---- this lane's work items of a Newton step (fixed for the whole kernel) ------------------------------
    int offA[F::NSLOT], offB[F::NSLOT], lpos[F::NSLOT], ldel[F::NSLOT], pixA[F::NSLOT], pixB[F::NSLOT];
    bool valid[F::NSLOT], ispair[F::NSLOT];
#pragma unroll
    for (int m = 0; m < F::NSLOT; m++) {
        const int i = lane + 32 * m;
        valid[m] = i < F::NIT;
        ispair[m] = i < F::NVI;
        int y, xa, xb;
        if (ispair[m]) {
            y = i / (4 * F::G);
            const int rem = i - y * 4 * F::G, g = rem >> 2, k = rem & 3;
            xa = 8 * g + k;
            xb = xa + 4;
            lpos[m] = k * F::LV + y * F::G + g;
            ldel[m] = 4 * F::LV;
        } else {
            const int t = valid[m] ? i - F::NVI : 0;
            y = t / F::TW;
            xa = F::VEC + t - y * F::TW;
            xb = xa;
            lpos[m] = valid[m] ? 8 * F::LV + t : F::LISTS / 4;   // an idle lane of the last slot writes to scratch
            ldel[m] = valid[m] ? F::LT : 1;
        }
        offA[m] = y * F::RS + xa;
        offB[m] = y * F::RS + xb;
        pixA[m] = y * ww + xa;
        pixB[m] = y * ww + xb;
    }
    const float *chain = sT + (lane < 8 ? lane * F::LV : 8 * F::LV + (lane & 1) * F::LT);  // lanes 8, 9: the tails

    const size_t o = (size_t)pair * pts_stride + pt;
    float2 p0;
    if (prev_xy) p0 = prev_xy[o];
    else p0 = make_float2((float)kp_col[(size_t)(prev_slot0 + pair) * pts_stride + pt],
                          (float)kp_row[(size_t)(prev_slot0 + pair) * pts_stride + pt]);
    const float halfx = __fmul_rn((float)(ww - 1), 0.5f), halfy = __fmul_rn((float)(wh - 1), 0.5f);
    const bool use_initial = (P.flags & 4) != 0, get_min_eig = (P.flags & 8) != 0;
    const float FLT_SCALE = 1.f / (1 << 20);
    float outx = 0.f, outy = 0.f, ev = 0.f;
    bool st = true;

    for (int level = L.top; level >= 0; level--) {
        const int H = L.H[level], W = L.W[level], pitch = L.pitch[level];
        const uint8_t *I = L.img[level] + (size_t)(prev_slot0 + pair) * L.slot_stride[level];
        const uint8_t *J = L.img[level] + (size_t)(next_slot0 + pair) * L.slot_stride[level];
        const float sc = __int_as_float((127 - level) << 23);  // 1 / (1 << level)
        float px = __fmul_rn(p0.x, sc), py = __fmul_rn(p0.y, sc);
        float nx, ny;
        if (level == L.top) {
            if (use_initial) {
                const float2 q = init_xy[o];
                nx = __fmul_rn(q.x, sc);
                ny = __fmul_rn(q.y, sc);
            } else {
                nx = px;
                ny = py;
            }
        } else {
            nx = __fmul_rn(outx, 2.f);
            ny = __fmul_rn(outy, 2.f);
        }
        outx = nx;
        outy = ny;
        px = __fsub_rn(px, halfx);
        py = __fsub_rn(py, halfy);
        const int ix = __float2int_rd(px), iy = __float2int_rd(py);
        if (ix < -ww || ix >= W || iy < -wh || iy >= H) {
            if (level == 0) {
                st = false;
                ev = 0.f;
            }
            continue;
        }
        KltW w = klt_weights(__fsub_rn(px, (float)ix), __fsub_rn(py, (float)iy));

        // ---- raw patch rows iy-1 .. iy+wh+1, cols ix-1 .. ix+ww+1 of the previous frame ----------------
        __syncwarp();
        const bool raw_in = iy >= 1 && ix >= 1 && iy + wh + 2 <= H && ix + ww + 2 <= W;
        int sh = 0;
        if (raw_in) {
            sh = (ix - 1) & 3;
            klt_stage_words<WH + 3, F::IW>(reinterpret_cast<uint32_t *>(sR), I, pitch, iy - 1, (ix - 1) & ~3, lane);
        } else {
            klt_stage_reflect<WH + 3, WW + 3, F::RS1>(sR, I, pitch, H, W, iy - 1, ix - 1, lane);
        }
        __syncwarp();
        const uint8_t *R = sR + sh;
        // Scharr (dx, dy) at window positions (iy + r, ix + c), r <= wh, c <= ww; zero outside the image
#pragma unroll
        for (int p = lane; p < (wh + 1) * gw; p += 32) {
            const int r = p / gw, c = p - r * gw;
            short2 g = make_short2(0, 0);
            if (raw_in || (iy + r >= 0 && iy + r < H && ix + c >= 0 && ix + c < W)) {
                const uint8_t *q0 = R + r * F::RS1 + c, *q1 = q0 + F::RS1, *q2 = q1 + F::RS1;
                const int t0a = (q0[0] + q2[0]) * 3 + q1[0] * 10, t0c = (q0[2] + q2[2]) * 3 + q1[2] * 10;
                const int t1a = q2[0] - q0[0], t1b = q2[1] - q0[1], t1c = q2[2] - q0[2];
                g.x = (short)(t0c - t0a);
                g.y = (short)((t1c + t1a) * 3 + t1b * 10);
            }
            sG[p] = g;
        }
        __syncwarp();
        // bilinear samples of the patch and its derivatives; float terms of the gradient matrix
#pragma unroll
        for (int p = lane; p < area; p += 32) {
            const int y = p / ww, x = p - y * ww;
            const int iv = klt_bilin<F::RS1>(R + (y + 1) * F::RS1 + (x + 1), w);
            const short2 g00 = sG[y * gw + x], g01 = sG[y * gw + x + 1], g10 = sG[(y + 1) * gw + x], g11 = sG[(y + 1) * gw + x + 1];
            const int gx = (g00.x * w.w00 + g01.x * w.w01 + g10.x * w.w10 + g11.x * w.w11 + (1 << 13)) >> 14;
            const int gy = (g00.y * w.w00 + g01.y * w.w01 + g10.y * w.w10 + g11.y * w.w11 + (1 << 13)) >> 14;
            sI[p] = (int16_t)iv;
            sIx[p] = (int16_t)gx;
            sIy[p] = (int16_t)gy;
            const float fx = (float)gx, fy = (float)gy;
            sT[p] = __fmul_rn(fx, fx);
            sT[area + p] = __fmul_rn(fx, fy);
            sT[2 * area + p] = __fmul_rn(fy, fy);
        }
        __syncwarp();
        // gradient matrix: lanes 0-3 A11, 4-7 A12, 8-11 A22 lane accumulators; 12-14 their tails
        float v = 0.f;
        if (lane < 12) v = klt_chain_f(sT + (lane >> 2) * area, ww, wh, lane & 3);
        else if (lane < 15) v = klt_chain_f(sT + (lane - 12) * area, ww, wh, 4);
        // this lane's samples of the previous frame, kept in registers for the Newton steps of this level
        int IvA[F::NSLOT], IxA[F::NSLOT], IyA[F::NSLOT], IvB[F::NSLOT], IxB[F::NSLOT], IyB[F::NSLOT];
#pragma unroll
        for (int m = 0; m < F::NSLOT; m++) {
            IvA[m] = sI[pixA[m]];
            IxA[m] = sIx[pixA[m]];
            IyA[m] = sIy[pixA[m]];
            IvB[m] = sI[pixB[m]];
            IxB[m] = ispair[m] ? sIx[pixB[m]] : 0;   // a tail item's second pixel contributes nothing
            IyB[m] = ispair[m] ? sIy[pixB[m]] : 0;
        }
        const float A11 = __fmul_rn(klt_combine(v, 0, 12), FLT_SCALE), A12 = __fmul_rn(klt_combine(v, 4, 13), FLT_SCALE),
                    A22 = __fmul_rn(klt_combine(v, 8, 14), FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dd = __fsub_rn(A11, A22);
        const float rad = __fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12));
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * ww * wh));
        if (get_min_eig) ev = minEig;
        if (minEig < P.min_eig || D < 1.1920928955078125e-07f) {
            if (level == 0) st = false;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, halfx);
        ny = __fsub_rn(ny, halfy);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < P.max_count; j++) {
            const int jx = __float2int_rd(nx), jy = __float2int_rd(ny);
            if (jx < -ww || jx >= W || jy < -wh || jy >= H) {
                if (level == 0) st = false;
                break;
            }
            w = klt_weights(__fsub_rn(nx, (float)jx), __fsub_rn(ny, (float)jy));
            __syncwarp();
            int js = 0;
            if (jy >= 0 && jx >= 0 && jy + wh + 1 <= H && jx + ww + 1 <= W) {
                js = jx & 3;
                klt_stage_words<WH + 1, F::JW>(reinterpret_cast<uint32_t *>(sR), J, pitch, jy, jx & ~3, lane);
            } else {
                klt_stage_reflect<WH + 1, WW + 1, F::RS>(sR, J, pitch, H, W, jy, jx, lane);
            }
            __syncwarp();
            const uint8_t *Q = sR + js;
#pragma unroll
            for (int m = 0; m < F::NSLOT; m++) {  // no branches: idle lanes compute pixel (0, VEC) again into scratch
                const int da = klt_bilin<F::RS>(Q + offA[m], w) - IvA[m];
                int t1 = da * IxA[m], t2 = da * IyA[m];
                if (32 * m < F::NVI) {  // this slot holds pair items (for its tail lanes IxB = IyB = 0)
                    const int db = klt_bilin<F::RS>(Q + offB[m], w) - IvB[m];
                    t1 += db * IxB[m];
                    t2 += db * IyB[m];
                }
                sT[lpos[m]] = __int2float_rn(t1);
                sT[lpos[m] + ldel[m]] = __int2float_rn(t2);
            }
            __syncwarp();
            // chains: lanes 0-3 b1 accumulators, 4-7 b2 accumulators, 8 / 9 the b1 / b2 tails
            float u = 0.f;
#pragma unroll
            for (int i = 0; i < F::CNTV; i += 4) {
                const float4 t = *reinterpret_cast<const float4 *>(chain + i);
                u = __fadd_rn(u, t.x);
                if (i + 1 < F::CNTV) u = __fadd_rn(u, t.y);
                if (i + 2 < F::CNTV) u = __fadd_rn(u, t.z);
                if (i + 3 < F::CNTV) u = __fadd_rn(u, t.w);
            }
            if ((lane >> 1) == 4) {
#pragma unroll
                for (int i = F::CNTV & ~3; i < F::CNTT; i += 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(chain + i);
                    if (i >= F::CNTV) u = __fadd_rn(u, t.x);
                    if (i + 1 >= F::CNTV && i + 1 < F::CNTT) u = __fadd_rn(u, t.y);
                    if (i + 2 >= F::CNTV && i + 2 < F::CNTT) u = __fadd_rn(u, t.z);
                    if (i + 3 >= F::CNTV && i + 3 < F::CNTT) u = __fadd_rn(u, t.w);
                }
            }
            // tail + ((q0 + q2) + (q1 + q3)) for both sums at once: lanes k and k^2 add, then k and k^1, lanes 0 / 4 add
            // their tail (lanes 8 / 9) and the two results are broadcast: 5 shuffles instead of 10
            const float tl = __shfl_sync(0xffffffffu, u, 8 + ((lane >> 2) & 1));  // lanes 0-3: b1 tail, 4-7: b2 tail
            float s1 = __fadd_rn(u, __shfl_xor_sync(0xffffffffu, u, 2));          // lanes 0,1 (4,5): q0+q2, q1+q3
            s1 = __fadd_rn(s1, __shfl_xor_sync(0xffffffffu, s1, 1));              // lane 0 (4): (q0+q2)+(q1+q3)
            const float sb = __fadd_rn(tl, s1);                                   // valid in lanes 0 and 4
            const float b1 = __fmul_rn(__shfl_sync(0xffffffffu, sb, 0), FLT_SCALE), b2 = __fmul_rn(__shfl_sync(0xffffffffu, sb, 4), FLT_SCALE);
            const float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, ddx);
            ny = __fadd_rn(ny, ddy);
            outx = __fadd_rn(nx, halfx);
            outy = __fadd_rn(ny, halfy);
            if (__dadd_rn(__dmul_rn((double)ddx, (double)ddx), __dmul_rn((double)ddy, (double)ddy)) <= P.eps2) break;
            // |float| < 0.01 (double)  <=>  |float| <= 0.01f: 0.01f is the largest float below the double 0.01
            if (j > 0 && fabsf(__fadd_rn(ddx, pdx)) <= 0.01f && fabsf(__fadd_rn(ddy, pdy)) <= 0.01f) {
                outx = __fsub_rn(outx, __fmul_rn(ddx, 0.5f));
                outy = __fsub_rn(outy, __fmul_rn(ddy, 0.5f));
                break;
            }
            pdx = ddx;
            pdy = ddy;
        }
        if (st && level == 0 && !get_min_eig) {
            // err = mean absolute difference of the patch at the final position (5 fractional bits removed)
            const float qx = __fsub_rn(outx, halfx), qy = __fsub_rn(outy, halfy);
            const int jx = __float2int_rd(qx), jy = __float2int_rd(qy);
            if (jx < -ww || jx >= W || jy < -wh || jy >= H) {
                st = false;
                continue;
            }
            w = klt_weights(__fsub_rn(qx, (float)jx), __fsub_rn(qy, (float)jy));
            __syncwarp();
            int js = 0;
            if (jy >= 0 && jx >= 0 && jy + wh + 1 <= H && jx + ww + 1 <= W) {
                js = jx & 3;
                klt_stage_words<WH + 1, F::JW>(reinterpret_cast<uint32_t *>(sR), J, pitch, jy, jx & ~3, lane);
            } else {
                klt_stage_reflect<WH + 1, WW + 1, F::RS>(sR, J, pitch, H, W, jy, jx, lane);
            }
            __syncwarp();
            const uint8_t *Q = sR + js;
            int e = 0;
#pragma unroll
            for (int m = 0; m < F::NSLOT; m++) {
                if (valid[m]) {
                    e += abs(klt_bilin<F::RS>(Q + offA[m], w) - IvA[m]);
                    if (ispair[m]) e += abs(klt_bilin<F::RS>(Q + offB[m], w) - IvB[m]);
                }
            }
#pragma unroll
            for (int s = 16; s; s >>= 1) e += __shfl_xor_sync(0xffffffffu, e, s);
            ev = __fdiv_rn(__fmul_rn((float)e, 1.f), (float)(32 * ww * wh));  // integer partial sums: order-free
        }
    }
    if (lane == 0) {
        next_xy[o] = make_float2(outx, outy);
        status[o] = st ? 1 : 0;
        err[o] = ev;
    }
}

// ------------------------------------------------------------------------------------------------
// K9  inlier count of the reference's fundamental-matrix RANSAC (src/3DHandler.cc:163-188).
// One CTA per candidate matrix F (the 8-point fits with cv::SVD stay on the host), threads stride over the
// matches: residual = p2.t() * F * p1 in double with OpenCV's left-to-right sums, inlier iff |residual| < threshold.
// A second tiny kernel picks the first maximum, the reference's `inlierCount > maxInliers` rule.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double epi_residual(const double *__restrict__ F, int x1, int y1, int x2, int y2) {
    const double a0 = (double)x2, a1 = (double)y2, b0 = (double)x1, b1 = (double)y1;
    const double r0 = __dadd_rn(__dadd_rn(__dmul_rn(a0, F[0]), __dmul_rn(a1, F[3])), F[6]);   // a2 = 1: 1 * F = F exactly
    const double r1 = __dadd_rn(__dadd_rn(__dmul_rn(a0, F[1]), __dmul_rn(a1, F[4])), F[7]);
    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(a0, F[2]), __dmul_rn(a1, F[5])), F[8]);
    return __dadd_rn(__dadd_rn(__dmul_rn(r0, b0), __dmul_rn(r1, b1)), r2);
}

__global__ void __launch_bounds__(256)
epipolar_inliers_kernel(const double *__restrict__ Fs, const int32_t *__restrict__ x1, const int32_t *__restrict__ y1,
                        const int32_t *__restrict__ x2, const int32_t *__restrict__ y2, int n, double threshold,
                        int32_t *__restrict__ counts, double *__restrict__ residuals) {
    __shared__ double F[9];
    __shared__ int red[8];
    const int i = blockIdx.x, tid = threadIdx.x;
    if (tid < 9) F[tid] = Fs[9 * (size_t)i + tid];
    __syncthreads();
    int c = 0;
    for (int k = tid; k < n; k += 256) {
        const double e = epi_residual(F, x1[k], y1[k], x2[k], y2[k]);
        if (residuals) residuals[(size_t)i * n + k] = e;
        c += fabs(e) < threshold;
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
    if ((tid & 31) == 0) red[tid >> 5] = c;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < 8; w++) t += red[w];
        counts[i] = t;
    }
}

__global__ void first_max_kernel(const int32_t *__restrict__ counts, int m, int32_t *__restrict__ best) {
    // one warp: lane-strided scan keeps the lowest index among equal maxima, then a shuffle reduction does the same
    const int lane = threadIdx.x;
    int bc = INT_MIN, bi = -1;
    for (int i = lane; i < m; i += 32)
        if (counts[i] > bc) {
            bc = counts[i];
            bi = i;
        }
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        const int oc = __shfl_xor_sync(0xffffffffu, bc, s), oi = __shfl_xor_sync(0xffffffffu, bi, s);
        if (oi >= 0 && (oc > bc || (oc == bc && (bi < 0 || oi < bi)))) {
            bc = oc;
            bi = oi;
        }
    }
    if (lane == 0) {
        best[0] = bi;
        best[1] = bc;
    }
}

}  // namespace yavo
