// blur_umma.cuh — the 9x9 sigma-2.5 fixed-point Gaussian of a 128 x 32 pixel tile (reference src/BriefDescriptor.cc:90:
// cv::GaussianBlur on CV_8U = taps {12,22,31,41,44,41,31,22,12}/256 per pass, one rounding (v + 32768) >> 16 at the
// end) as two banded Toeplitz products on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in
// TMEM).  Exact integer arithmetic: u8 x u8 products summed in int32.
//
//   pass 1 (horizontal)   D1[x, n] = sum_c T1[x, c] * P[n, c]       M = 128 output columns x, N = 48 staged rows
//                                                                   (40 used), K = 160 staged bytes, five instructions
//       A = T1 (constants): row x holds the nine taps at staged bytes x+12 .. x+20.  K slice k of T1 is the same
//           256 x 32 band matrix F read from row 128 - 32k on (A_k[x][j] = F[x + 128 - 32k][j], F[r][j] = g[j + 116 - r]):
//           8 KB of constants serve all five instructions; in the K-major no-swizzle core-matrix layout the row
//           shift is a shift of the descriptor's start address.
//       B = the staged pixel rows in core-matrix order [16-byte chunk c][row n][16 bytes] (LBO 640, SBO 128).
//   The 16-bit sums D1 are split into low and high bytes by the CTA's warps and written back to TMEM as the A
//   operands of
//   pass 2 (vertical)     D2lo/hi[x, r] = sum_n Hlo/hi[x, n] * T2[r, n]   M = 128, N = 16 output rows per half,
//                                                                          K = 32 staged rows (window 16*half ..)
//       T2[r][j] = g[j - r] is the same 16 x 32 matrix for both halves; out(r, x) = (D2lo + 256 D2hi + 32768) >> 16
//       (D2lo accumulates onto 32768 stored in its TMEM columns beforehand).
//   TMEM columns of a CTA (64 allocated, eight CTAs per SM): D1 at [0,48), A_lo at [40,52) (over D1's unused pad rows
//   only), A_hi at [52,64), D2 lo / hi of a half at [0,16) / [16,32).
#ifndef YAVO_BLUR_UMMA_CUH
#define YAVO_BLUR_UMMA_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "fast_core.h"  // yavo_blur_v2_raw (hybrid variant)

namespace yavo {
namespace bu {

constexpr int BU_SROW = 160, BU_SH = 40;                 // the detect kernel's staged tile: 40 rows of 160 bytes
constexpr int BU_UB_LBO = BU_SH * 16;                    // 640: distance between 16-byte K chunks of the pixel operand
constexpr int BU_UB_BYTES = 10 * BU_UB_LBO + 128;        // + the pad rows 40..47 of the last chunk
constexpr int BU_F_ROWS = 256, BU_F_LBO = BU_F_ROWS * 16;
constexpr int BU_F_BYTES = 2 * BU_F_LBO;                 // 8192
constexpr int BU_T2_LBO = 16 * 16, BU_T2_BYTES = 2 * BU_T2_LBO;  // 512
constexpr int BU_CONST_BYTES = BU_F_BYTES + BU_T2_BYTES;
constexpr uint32_t BU_TMEM_COLS = 64;
constexpr uint32_t BU_COL_ALO = 40, BU_COL_AHI = 52;

// host: the constant operands in their shared-memory layout (copied to the device once per context)
inline void bu_fill_constants(uint8_t *dst) {
    static const int g[9] = {12, 22, 31, 41, 44, 41, 31, 22, 12};
    for (int i = 0; i < BU_CONST_BYTES; i++) dst[i] = 0;
    for (int r = 0; r < BU_F_ROWS; r++)
        for (int j = 0; j < 32; j++) {
            const int t = j + 116 - r;
            if (t >= 0 && t <= 8) dst[(j / 16) * BU_F_LBO + (r / 8) * 128 + (r % 8) * 16 + j % 16] = (uint8_t)g[t];
        }
    for (int r = 0; r < 16; r++)
        for (int j = 0; j < 32; j++) {
            const int t = j - r;
            if (t >= 0 && t <= 8) dst[BU_F_BYTES + (j / 16) * BU_T2_LBO + (r / 8) * 128 + (r % 8) * 16 + j % 16] = (uint8_t)g[t];
        }
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t bu_saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bu_bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bu_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bu_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// K-major, no swizzle, sm_100 descriptor version (bit 46)
__device__ __forceinline__ uint64_t bu_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::i8 instruction descriptor: D = S32 (bits [4,6) = 2), A = B = U8 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t bu_idesc(int m, int n) {
    return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void bu_mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void bu_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void bu_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Kernel start: warp 0 allocates the CTA's 64 TMEM columns, thread 0 arms the barriers and starts the bulk copy of the
// constant operands (L2-resident).  bars[0]: constants, bars[1]: tcgen05.commit.  The CTA barrier that publishes the
// staged tile also publishes tmem_base_s (the caller issues tcgen05.fence::after_thread_sync via bu_relayout).
__device__ __forceinline__ void bu_prologue(uint8_t *cst, const uint8_t *__restrict__ consts, uint64_t *bars, uint32_t *tmem_base_s) {
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bu_saddr(tmem_base_s)), "r"(BU_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        if (threadIdx.x == 0) {
            const uint32_t b0 = bu_saddr(&bars[0]), b1 = bu_saddr(&bars[1]);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b1) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b0), "r"((uint32_t)BU_CONST_BYTES) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(bu_saddr(cst)),
                         "l"(consts), "r"((uint32_t)BU_CONST_BYTES), "r"(b0)
                         : "memory");
        }
        bu_fence_before();
    }
}

// The staged tile (row-major, 160-byte rows) -> core-matrix order for the tensor cores: 400 16-byte chunks, by 256
// threads.  Thread t takes row t & 63 (40 used) of chunk columns t >> 6, + 4, + 8: the stores of a warp are contiguous.
__device__ __forceinline__ void bu_relayout(const uint8_t *tile, uint8_t *ub) {
    const int n = threadIdx.x & 63, c0 = threadIdx.x >> 6;
    if (n < BU_SH) {
        const uint8_t *src = tile + n * BU_SROW + c0 * 16;
        uint8_t *dst = ub + c0 * BU_UB_LBO + n * 16;
        const uint4 v0 = *reinterpret_cast<const uint4 *>(src), v1 = *reinterpret_cast<const uint4 *>(src + 64);
        *reinterpret_cast<uint4 *>(dst) = v0;
        *reinterpret_cast<uint4 *>(dst + 4 * BU_UB_LBO) = v1;
        if (c0 < 2) *reinterpret_cast<uint4 *>(dst + 8 * BU_UB_LBO) = *reinterpret_cast<const uint4 *>(src + 128);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to tcgen05.mma's operand reads
}

__device__ __forceinline__ void bu_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void bu_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void bu_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void bu_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void bu_st2(uint32_t taddr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void bu_st1(uint32_t taddr, uint32_t a) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(a) : "memory");
}

// four 16-bit sums -> their low bytes and their high bytes as two packed words (K order = row order)
__device__ __forceinline__ void bu_split4(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t *lo, uint32_t *hi) {
    const uint32_t t1 = __byte_perm(a, b, 0x5140), t2 = __byte_perm(c, d, 0x5140);  // [a0 b0 a1 b1], [c0 d0 c1 d1]
    *lo = __byte_perm(t1, t2, 0x5410);
    *hi = __byte_perm(t1, t2, 0x7632);
}

// The blur of one tile in three steps, each called by all 256 threads of the CTA; the detect kernel puts its segment
// test between the first two and its corner scoring between the last two, so that the tensor pipe works underneath.
// Lane x of TMEM = output column x0 + x; warp w drains lane quarter w % 4, column half w / 4.
//
// Step 1 — after a CTA barrier that follows bu_relayout (or a tensor copy into `ub`): thread 0 issues pass 1.
__device__ __forceinline__ void bu_pass1_issue(const uint8_t *ub, const uint8_t *cst, uint64_t *bars, uint32_t tmem_base) {
    bu_fence_after();
    if (threadIdx.x == 0) {
        bu_bar_wait(bu_saddr(&bars[0]), 0);  // the constant operands have landed
        const uint32_t sF = bu_saddr(cst), sU = bu_saddr(ub);
#pragma unroll
        for (int k = 0; k < 5; k++)
            bu_mma_ss(tmem_base, bu_desc(sF + (16 - 4 * k) * 128, BU_F_LBO, 128), bu_desc(sU + 2 * k * BU_UB_LBO, BU_UB_LBO, 128),
                      bu_idesc(128, 48), k > 0);
        bu_commit(bu_saddr(&bars[1]));
    }
}

__device__ __forceinline__ void bu_pass2_issue(const uint8_t *cst, uint64_t *bars, uint32_t tmem_base, int half) {
    if (threadIdx.x == 0) {
        const uint64_t t2desc = bu_desc(bu_saddr(cst + BU_F_BYTES), BU_T2_LBO, 128);
        bu_fence_after();
        bu_mma_ts(tmem_base, tmem_base + BU_COL_ALO + 4 * half, t2desc, bu_idesc(128, 16), 1);  // onto the rounding constant
        bu_mma_ts(tmem_base + 16, tmem_base + BU_COL_AHI + 4 * half, t2desc, bu_idesc(128, 16), 0);
        bu_commit(bu_saddr(&bars[1]));
    }
}

__device__ __forceinline__ void bu_st8_const(uint32_t taddr, uint32_t c) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(c) : "memory");
}

// Step 2 — pass 1 drained: the 16-bit row sums of this thread's column become the byte operands of pass 2, whose first
// half (output rows 0..15 of the tile) is issued.  The low-byte accumulator starts at the rounding constant 32768.
__device__ __forceinline__ void bu_pass1_drain(const uint8_t *cst, uint64_t *bars, uint32_t tmem_base) {
    const int warp = threadIdx.x >> 5, q = warp & 3, h = warp >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
    if (threadIdx.x == 0) bu_bar_wait(bu_saddr(&bars[1]), 0);
    __syncthreads();
    bu_fence_after();
    {   // D1 rows 20h .. 20h+19 -> five low-byte and five high-byte words (eight rows at a time: the detect kernel
        // runs at 32 registers per thread)
#pragma unroll
        for (int g8 = 0; g8 < 2; g8++) {
            uint32_t v[8], lo[2], hi[2];
            bu_ld8(lane_base + 20 * h + 8 * g8, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            bu_split4(v[0], v[1], v[2], v[3], &lo[0], &hi[0]);
            bu_split4(v[4], v[5], v[6], v[7], &lo[1], &hi[1]);
            bu_st2(lane_base + BU_COL_ALO + 5 * h + 2 * g8, lo[0], lo[1]);
            bu_st2(lane_base + BU_COL_AHI + 5 * h + 2 * g8, hi[0], hi[1]);
        }
        uint32_t w[4], lo4, hi4;
        bu_ld4(lane_base + 20 * h + 16, w);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        bu_split4(w[0], w[1], w[2], w[3], &lo4, &hi4);
        bu_st1(lane_base + BU_COL_ALO + 5 * h + 4, lo4);
        bu_st1(lane_base + BU_COL_AHI + 5 * h + 4, hi4);
        // columns 0..15 (this warp's own rows 0..15 of D1, now in registers / converted) become the low-byte accumulator
        if (h == 0) {
            bu_st8_const(lane_base, 32768u);
            bu_st8_const(lane_base + 8, 32768u);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    bu_fence_before();
    __syncthreads();
    bu_pass2_issue(cst, bars, tmem_base, 0);
}

// Step 3 — one half of pass 2 drained: out = (lo + 256 hi) >> 16 (the rounding constant is in lo already).  The bytes go
// to shared memory (`ob`, 32 rows of 128 bytes: the pixel operand's buffer, free since pass 1 completed).  After the
// first half's accumulators are in registers the second half is issued.
__device__ __forceinline__ void bu_pass2_drain(const uint8_t *cst, uint64_t *bars, uint32_t tmem_base, uint8_t *ob, int half) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = warp & 3, h = warp >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
    if (tid == 0) bu_bar_wait(bu_saddr(&bars[1]), (half + 1) & 1);
    __syncthreads();
    bu_fence_after();
    uint32_t lo[8], hi[8];
    bu_ld8(lane_base + 8 * h, lo);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (half == 0) bu_st8_const(lane_base + 8 * h, 32768u);  // this warp's own low-byte columns, for the second half
    bu_ld8(lane_base + 16 + 8 * h, hi);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (half == 0) {  // the accumulators are in registers: the second half may overwrite them
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        bu_fence_before();
        __syncthreads();
        bu_pass2_issue(cst, bars, tmem_base, 1);
    }
    uint8_t *o = ob + (16 * half + 8 * h) * 128 + 32 * q + lane;
#pragma unroll
    for (int j = 0; j < 8; j++) o[j * 128] = (uint8_t)((lo[j] + (hi[j] << 8)) >> 16);
}

// Step 4 — after the second half: TMEM freed, the tile's 32 x 128 blurred bytes written with one 16-byte store per thread.
__device__ __forceinline__ void bu_finish(uint32_t tmem_base, const uint8_t *ob, uint8_t *__restrict__ blur_frame, int pitch, int H,
                                          int x0, int y0) {
    const int tid = threadIdx.x;
    bu_fence_before();
    __syncthreads();
    if (tid < 32) {
        bu_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(BU_TMEM_COLS) : "memory");
    }
    const int r = tid >> 3, c = (tid & 7) * 16, gr = y0 + r, gc = x0 + c;
    if (gr < H && gc < pitch)
        *reinterpret_cast<uint4 *>(blur_frame + (uint32_t)(gr * pitch + gc)) = *reinterpret_cast<const uint4 *>(ob + r * 128 + c);
}

// ---- hybrid: pass 1 on the tensor cores, pass 2 from the accumulator registers on the integer pipe -----------------
// After pass 1 (bu_pass1_issue + the mbarrier) a thread holds the row sums of ONE output column: warp w takes lane
// quarter w % 4 and output rows 16 (w / 4) .. + 15, i.e. the 24 sums D1[x, 16 h .. 16 h + 23].  Vertical pairs of them
// are the operands of the IDP.2A vertical pass (yavo_blur_v2_raw, the same arithmetic as the integer-pipe kernel);
// the bytes go to `ob` (32 rows of 128 bytes) for bu_copy_out.  One MMA round trip per tile, no byte split, no
// operands written back to TMEM.
// `part` 0 / 1 = the first / last eight of the warp's sixteen output rows (the accumulator stays in TMEM, so the second
// part can run later — the detect kernel puts it behind the pool reservation's atomic, whose latency it hides).
__device__ __forceinline__ void bu_vpass_from_tmem(uint64_t *bars, uint32_t tmem_base, uint8_t *ob, int part) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = warp & 3, h = warp >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16) + 16 * h + 8 * part;
    if (part == 0) {
        // every thread waits on the mbarrier itself (no CTA barrier: the segment test's first pass has run in between,
        // the first try_wait succeeds) — warps go on as they arrive
        bu_bar_wait(bu_saddr(&bars[1]), 0);
        bu_fence_after();
    }
    uint32_t P[8];
#pragma unroll
    for (int g8 = 0; g8 < 2; g8++) {
        uint32_t v[8];
        bu_ld8(lane_base + 8 * g8, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 4; i++) P[4 * g8 + i] = __byte_perm(v[2 * i], v[2 * i + 1], 0x5410);  // sums < 2^16: low halves
    }
    uint8_t *o = ob + (16 * h + 8 * part) * 128 + 32 * q + lane;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        uint32_t a, b;
        yavo_blur_v2_raw(&P[j], &a, &b);
        o[(2 * j) * 128] = (uint8_t)(a >> 16);
        o[(2 * j + 1) * 128] = (uint8_t)(b >> 16);
    }
}

#endif  // __CUDACC__

}  // namespace bu
}  // namespace yavo

#endif  // YAVO_BLUR_UMMA_CUH
