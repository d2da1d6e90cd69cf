// match_tc4.cuh — K5t4: the tensor-core Hamming matcher of match_tc.cuh on PACKED 4-bit operands
// (tcgen05.mma kind::mxf4, block-scaled, K = 64 per instruction).  Same reference semantics
// (Brief::matchFeatures, reference src/BriefDescriptor.cc:139-183) and the same bit-exact argument; what changes:
//
//   * a descriptor bit becomes one e2m1 nibble: query bit 0 -> +1.0 (0x2), 1 -> -1.0 (0xA); train bit 0 -> -1.0,
//     1 -> +1.0, and the train operand's block scale factors (UE8M0, kept in TMEM) are all 2^7, the query operand's
//     all 1: sum_k a_k * b_k * 128 = 256 * hamming - 32768, every partial sum an integer < 2^16, exact in FP32.
//     Eight nibbles come out of ONE LOP3 ((w << k) & 0x88888888, | or ^ a constant) instead of four bytes, so the
//     expansion — which the integer ALU pipe (64 lanes / clk / SM) bounds in the FP8 version — costs half, the
//     operand tiles are half as large in shared memory, and one instruction covers 64 bit positions:
//     4 + 1 instructions of 112 clk per 128 x 224 tile instead of 8 + 1 of 128 clk per 128 x 256 tile.
//   * all scale factors of an operand are EQUAL, so the TMEM layout of the scale-factor matrices does not matter:
//     two 32-column regions are filled with the bytes 0x7F (1.0) and 0x86 (128.0) once per CTA.
//   * the fifth instruction adds the train column j (scale factors 1 on both sides): query nibbles
//     {1, 4, 4 x4, 4 x16} times train nibbles {j & 3, (j >> 2) & 3, (j >> 4) & 3 x4, (j >> 6) & 3 x16}.
//   * N = 224: two accumulators (448 columns) and the scale factors (64 columns) fill the 512 TMEM columns;
//     2000 keypoints = 9 tiles of 224 with 0.8 % padding.  (-DYAVO_TC4_N=144 / 112 give three / four narrower
//     accumulators; both measured slower on B200: the fixed cost per tile decides, profiles/r1k_match_tc_ncu_summary.md.)
//   * both epilogue warps of a TMEM lane quarter drain every tile, 112 columns each; the role branches are outermost,
//     each role with its own loop over the work items (83 registers per thread).
//
// The -DYAVO_TC_EXP_* macros switch single pieces of work off for the timing experiments recorded in profiles/
// (results are wrong with any of them defined; they are never set by the library build).
//
// Warp roles, barriers and the first-minimum rule are those of match_tc.cuh.
#pragma once
#include "match_tc.cuh"

namespace yavo {
namespace tcm4 {

using namespace tcm;  // barrier / fence / descriptor / tcgen05.ld helpers

constexpr int Q4 = 128;                  // queries per work item (UMMA M)
#ifndef YAVO_TC4_N
#define YAVO_TC4_N 224
#endif
constexpr int T4 = YAVO_TC4_N;           // train descriptors per tile (UMMA N): 224 (two accumulators), 144 (three) or 112 (four)
constexpr int NACC = 448 / T4;           // accumulators in TMEM (448 columns; the scale factors take the other 64)
constexpr int RPW = T4 > 128 ? 64 : 32;  // train rows per expander warp
#ifndef YAVO_TC4_EPI_SPLIT
#define YAVO_TC4_EPI_SPLIT 1
#endif
// epilogue split: true = both warps of a TMEM lane quarter drain EVERY tile, 112 columns each (two tcgen05.ld round trips per
// tile and warp); false = warps 0-3 / 4-7 take alternate tiles whole (four round trips)
constexpr bool EPI_SPLIT = YAVO_TC4_EPI_SPLIT && (T4 == 224 || T4 == 144);
#ifndef YAVO_TC4_EPI_GROUPS
#define YAVO_TC4_EPI_GROUPS 2
#endif
// epilogue groups: every TMEM lane quarter is drained by EG warps (warp & 3 = quarter, warp >> 2 = group), each taking
// T4 / EG columns of every tile: 2 groups of 112 columns (8 warps), or 4 groups of 56 (16 warps: half the drain latency
// per tile, which is what the MMA warp waits on when both accumulators are full)
constexpr int EG = YAVO_TC4_EPI_GROUPS;
static_assert(EG == 2 || (EG == 4 && YAVO_TC4_N == 224 && YAVO_TC4_EPI_SPLIT), "4 epilogue groups: 224-column tiles, split mode");
constexpr int EPI4_WARPS = 4 * EG;
constexpr int MMA4_WARP = EPI4_WARPS;                       // then EXP_WARPS expander warps, then the loader warp
constexpr int LOAD4_WARP = EPI4_WARPS + 1 + EXP_WARPS;
constexpr int THREADS4 = 32 * (EPI4_WARPS + 2 + EXP_WARPS);
#ifndef YAVO_TC4_NCH
#define YAVO_TC4_NCH 4
#endif
constexpr int NCH = YAVO_TC4_NCH;        // independent minimum chains per epilogue thread (4 or 8)
constexpr int CW = EG == 4 ? 56 : (T4 == 144 ? 72 : 112);  // accumulator columns per epilogue pass: 32 + 16 + 8, 32 + 32 + 8 or 32 + 32 + 32 + 16
static_assert(T4 % CW == 0, "tile width");
constexpr int ROWB = 128;                // operand bytes per descriptor (two e2m1 per byte)
constexpr int A4_BYTES = Q4 * ROWB;      // 16 KB
constexpr int B4_BYTES = T4 * ROWB;      // 28 KB / 14 KB
constexpr int AX4_BYTES = Q4 * 32;       // constant index slice, queries
constexpr int BX4_BYTES = T4 * 32;       // constant index slice, train columns
constexpr int RING4_BYTES = T4 * 32;     // packed bits of one train tile
constexpr int NB4 = 4;                   // train-tile operand stages and packed-bit ring entries (even: stage parity = expander group)
#ifndef YAVO_TC4_NR
#define YAVO_TC4_NR 4
#endif
constexpr int NR4 = YAVO_TC4_NR;         // packed-bit ring entries (even); 8 measured no faster than 4
constexpr int SMEM4_BYTES = NSTAGE * A4_BYTES + NB4 * B4_BYTES + AX4_BYTES + BX4_BYTES + NR4 * RING4_BYTES;
constexpr uint32_t SF_ONE_COL = 448, SF_128_COL = 480;  // TMEM columns of the two scale-factor regions

// Block-scaled instruction descriptor (kind::mxf4): A = B = E2M1 (1) at bits 7 / 10, both K-major, N >> 3 at bit 17,
// scale format UE8M0 (1) at bit 23, M >> 4 at bit 24, K = 64 (bit 31 = 0), scale-factor ids 0.
constexpr uint32_t IDESC4 = (1u << 7) | (1u << 10) | ((uint32_t)(T4 >> 3) << 17) | (1u << 23) | ((uint32_t)(Q4 >> 4) << 24);

__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(p));
    return p != 0;
}

__device__ __forceinline__ void mma_f4(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t sfa, uint32_t sfb,
                                       uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC4), "r"(accumulate), "r"(sfa), "r"(sfb)
        : "memory");
}

// 32 TMEM columns of this warp's lane quarter <- one 32-bit value
__device__ __forceinline__ void tmem_fill32(uint32_t taddr, uint32_t v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1};" ::"r"(taddr),
        "r"(v)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// 8 columns into the first half of a 16-register buffer
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&d)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7]), "+r"(d[8]), "+r"(d[9]),
                   "+r"(d[10]), "+r"(d[11]), "+r"(d[12]), "+r"(d[13]), "+r"(d[14]), "+r"(d[15])::"memory");
}
// one tcgen05.wait::ld for two loads in flight; every destination register is tied to it
__device__ __forceinline__ void tmem_wait2(uint32_t (&a)[32], uint32_t (&b)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : YAVO_TM32_RW(a)::"memory");
    asm volatile("" : YAVO_TM32_RW(b)::"memory");
}
__device__ __forceinline__ void tmem_wait2b(uint32_t (&a)[32], uint32_t (&d)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : YAVO_TM32_RW(a)::"memory");
    asm volatile(""
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3]), "+r"(d[4]), "+r"(d[5]), "+r"(d[6]), "+r"(d[7]), "+r"(d[8]), "+r"(d[9]),
                   "+r"(d[10]), "+r"(d[11]), "+r"(d[12]), "+r"(d[13]), "+r"(d[14]), "+r"(d[15])::"memory");
}

// One descriptor (8 words = 256 bits) -> 128 operand bytes = chunks 0..7 of row r of a tile with R rows;
// word i becomes chunk i: nibble q of output word s holds bit 4q + s of the word.
template <bool TRAIN>
__device__ __forceinline__ void expand_row4(uint32_t tile, int R, int r, const uint4 &lo, const uint4 &hi) {
    const uint32_t msk = 0x88888888u, cst = TRAIN ? 0xAAAAAAAAu : 0x22222222u;
    uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    uint32_t p = tile + (uint32_t)(r * 16);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t o[4];
#pragma unroll
        for (int s = 0; s < 4; s++) o[s] = TRAIN ? and_xor(w[i] << (3 - s), msk, cst) : and_or(w[i] << (3 - s), msk, cst);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
        p += R * 16;
    }
}

// e2m1 nibble of an integer 0..4
__device__ __forceinline__ uint32_t e2m1_small_int(uint32_t n) { return n == 0 ? 0u : n == 1 ? 2u : n == 2 ? 4u : n == 3 ? 5u : 6u; }

// the 22 nibbles of the index slice: weights (queries) or digits of j (train column)
__device__ __forceinline__ uint4 index_slice(bool query, uint32_t j) {
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 22; e++) {
        const uint32_t digit = e == 0 ? (j & 3) : e == 1 ? ((j >> 2) & 3) : e < 6 ? ((j >> 4) & 3) : ((j >> 6) & 3);
        const uint32_t v = query ? (e == 0 ? 1u : 4u) : digit;
        w[e >> 3] |= e2m1_small_int(v) << (4 * (e & 7));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <bool FULL>
__device__ __forceinline__ void min_keys4(const uint32_t (&v)[32], float (&m)[NCH], int col0, int nvalid) {
    if (FULL) {
#pragma unroll
        for (int i = 0; i < 32; i += 2 * NCH) {
#pragma unroll
            for (int c = 0; c < NCH; c++) m[c] = fminf(m[c], fminf(__uint_as_float(v[i + 2 * c]), __uint_as_float(v[i + 2 * c + 1])));
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; i++)
            if (col0 + i < nvalid) m[i & (NCH - 1)] = fminf(m[i & (NCH - 1)], __uint_as_float(v[i]));
    }
}

// the two smallest keys of every chain (SECOND: ratio-test extension).  Per two new keys: lo / hi of the pair, then
// m2 = min(m2, hi, max(m1, lo)), m1 = min(m1, lo) — five minimum / maximum operations instead of one.
template <bool FULL>
__device__ __forceinline__ void min2_keys4(const uint32_t (&v)[32], float (&m)[NCH], float (&m2)[NCH], int n, int col0, int nvalid) {
    if (FULL) {
#pragma unroll
        for (int i = 0; i < 32; i += 2 * NCH) {
            if (i >= n) break;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const float a = __uint_as_float(v[i + 2 * c]), b = __uint_as_float(v[i + 2 * c + 1]);
                const float lo = fminf(a, b), hi = fmaxf(a, b);
                m2[c] = fminf(fminf(m2[c], hi), fmaxf(m[c], lo));
                m[c] = fminf(m[c], lo);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; i++) {
            if (i >= n) break;
            if (col0 + i < nvalid) {
                const float a = __uint_as_float(v[i]);
                m2[i & (NCH - 1)] = fminf(m2[i & (NCH - 1)], fmaxf(m[i & (NCH - 1)], a));
                m[i & (NCH - 1)] = fminf(m[i & (NCH - 1)], a);
            }
        }
    }
}

// every role walks the same list of work items: role branches outermost, one item loop per role (83 instead of 96
// registers and 9 % less time than one shared item loop with the role branches inside)
#define YAVO_TC4_FOR_ITEMS \
    for (int item = blockIdx.x; item < items; item += gridDim.x) { \
        const int pair = item / q_tiles, q0 = (item - pair * q_tiles) * Q4; \
        const int nq = nq_all ? nq_all[pair + q_set_offset] : nq_fixed; \
        const int nt = nt_all ? nt_all[pair + t_set_offset] : nt_fixed; \
        if (q0 >= nq) continue; \
        const int n_tiles = (nt + T4 - 1) / T4; \
        const uint32_t *dq = dq_all + (size_t)(pair + q_set_offset) * set_stride_words; \
        const uint32_t *dt = dt_all + (size_t)(pair + t_set_offset) * set_stride_words; \
        (void)dq; (void)dt; (void)nt; (void)n_tiles;

// Arguments as match_tc_kernel.  dbg_acc (test tool only): the 128 x 224 accumulator values (key - 32768) of the
// first tile of work item 0.
template <bool DBG, bool SECOND = false>
__global__ void __launch_bounds__(THREADS4, 1)
match_tc4_kernel(const uint32_t *__restrict__ dq_all, const int *__restrict__ nq_all, int nq_fixed,
                 const uint32_t *__restrict__ dt_all, const int *__restrict__ nt_all, int nt_fixed,
                 size_t set_stride_words, int q_set_offset, int t_set_offset, int pairs, int q_tiles, int out_stride,
                 int32_t *__restrict__ out_idx, int32_t *__restrict__ out_dist, float *__restrict__ dbg_acc,
                 int32_t *__restrict__ out_second = nullptr) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bars[2 * NSTAGE + 2 * NACC + 2 * NB4 + 2 * NR4];
    __shared__ uint32_t tmem_base_s;
    __shared__ int2 comb[2][EG - 1][Q4];
    __shared__ int comb_sec[SECOND ? 2 : 1][EG - 1][SECOND ? Q4 : 1];
    // shared-window addresses, computed once (barriers are 8 bytes apart)
    uint32_t bars_s, smem_s;  // through an opaque move: the compiler otherwise rematerialises the conversion (S2R + LEA) at every use
    asm volatile("mov.u32 %0, %2;\n\tmov.u32 %1, %3;" : "=r"(bars_s), "=r"(smem_s) : "r"(saddr(bars)), "r"(saddr(smem_raw)));
    const uint32_t a_full = bars_s, a_empty = a_full + 8 * NSTAGE, acc_full = a_empty + 8 * NSTAGE, acc_empty = acc_full + 8 * NACC;
    const uint32_t b_full = acc_empty + 8 * NACC, b_empty = b_full + 8 * NB4, r_full = b_empty + 8 * NB4, r_empty = r_full + 8 * NR4;
    const uint32_t sA = smem_s, sB = sA + NSTAGE * A4_BYTES, sAX = sB + NB4 * B4_BYTES, sBX = sAX + AX4_BYTES, sRing = sBX + BX4_BYTES;
    uint8_t *const gAX = smem_raw + NSTAGE * A4_BYTES + NB4 * B4_BYTES, *const gBX = gAX + AX4_BYTES;  // generic pointers (set-up only)
    const uint8_t *const gRing = gBX + BX4_BYTES;
    // warp index through a shuffle: provably warp-uniform, so the role branches and everything the MMA warp computes
    // stay in uniform registers (with a lane-0 branch around them every tcgen05.mma costs an ELECT / R2UR.BROADCAST loop)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            bar_init(a_full + 8 * (s), EXP_WARPS / 2);
            bar_init(a_empty + 8 * (s), 1);
        }
        for (int s = 0; s < NACC; s++) {
            bar_init(acc_full + 8 * (s), 1);
            bar_init(acc_empty + 8 * (s), EPI_SPLIT ? EPI4_WARPS : EPI4_WARPS / 2);
        }
        for (int s = 0; s < NB4; s++) {
            bar_init(b_full + 8 * (s), EXP_WARPS / 2);
            bar_init(b_empty + 8 * (s), 1);
        }
        for (int s = 0; s < NR4; s++) {
            bar_init(r_full + 8 * (s), 1);
            bar_init(r_empty + 8 * (s), EXP_WARPS / 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // constant index slice: chunk 0 = 22 nibbles, chunk 1 = 0
    for (int i = threadIdx.x; i < Q4 + T4; i += THREADS4) {
        const bool q = i < Q4;
        const int r = q ? i : i - Q4;
        uint8_t *p = q ? gAX + r * 16 : gBX + r * 16;
        *reinterpret_cast<uint4 *>(p) = index_slice(q, (uint32_t)r);
        *reinterpret_cast<uint4 *>(p + (q ? Q4 : T4) * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_async_smem();
    if (warp == MMA4_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);
    if (warp < 4) {  // scale factors: every byte of a region is the same UE8M0 value
        const uint32_t lanes = (uint32_t)(warp * 32) << 16;
        tmem_fill32(tmem_base + lanes + SF_ONE_COL, 0x7f7f7f7fu);  // 2^0
        tmem_fill32(tmem_base + lanes + SF_128_COL, 0x86868686u);  // 2^7
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();

    const int items = pairs * q_tiles;
    uint32_t e_cnt = 0, a_cnt = 0, t_cnt = 0;

    if (warp < EPI4_WARPS) {
        YAVO_TC4_FOR_ITEMS
            // ------------------------------------------------ epilogue (two groups on alternate accumulators)
            const int g = warp >> 2, row = (warp & 3) * 32 + lane;
            int best_d = 0x7fffffff, best_j = -1, sec_d = 0x7fffffff;  // sec_d: second smallest distance (SECOND)
            for (int t = EPI_SPLIT ? 0 : (((t_cnt & 1) == (uint32_t)g) ? 0 : 1); t < n_tiles; t += EPI_SPLIT ? 1 : 2) {
                const uint32_t acc = (t_cnt + t) % NACC, ph = ((t_cnt + t) / NACC) & 1;
                bar_wait(acc_full + 8 * (acc), ph);
                fence_after_sync();
                const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc * T4;
                const int nvalid = min(T4, nt - t * T4);
                float m4[NCH], s4[NCH];  // s4: chain seconds (SECOND only; dead code otherwise)
#pragma unroll
                for (int c = 0; c < NCH; c++) m4[c] = 3.0e38f;
                if (SECOND) {
#pragma unroll
                    for (int c = 0; c < NCH; c++) s4[c] = 3.0e38f;
                }
                uint32_t v0[32], v1[32], v2[32], v3[16];
                const bool full = nvalid == T4;
                // CW columns per pass = 32 + 32 | 32 + 16 (or | 8): the second group is loaded while the first is reduced
#pragma unroll 1
                for (int h = EPI_SPLIT ? g : 0; h < (EPI_SPLIT ? g + 1 : T4 / CW); h++) {
                    const uint32_t ta = taddr + h * CW;
                    const int c0 = h * CW;
                    if (CW == 56) {  // four groups: 32 + 16 + 8 columns, one round trip
                        uint32_t v4[16];
                        tmem_ld32(ta, v0);
                        tmem_ld16(ta + 32, v3);
                        tmem_ld8(ta + 48, v4);
                        tmem_wait2b(v0, v3);
                        tmem_wait16(v4);
                        uint32_t vt[32], vu[32];
#pragma unroll
                        for (int i = 0; i < 32; i++) {
                            vt[i] = v3[i & 15];
                            vu[i] = v4[i & 7];
                        }
                        if (SECOND) {
                            if (full) {
                                min2_keys4<true>(v0, m4, s4, 32, 0, 0);
                                min2_keys4<true>(vt, m4, s4, 16, 0, 0);
                                min2_keys4<true>(vu, m4, s4, 8, 0, 0);
                            } else {
                                min2_keys4<false>(v0, m4, s4, 32, c0, nvalid);
                                min2_keys4<false>(vt, m4, s4, 16, c0 + 32, nvalid);
                                min2_keys4<false>(vu, m4, s4, 8, c0 + 48, nvalid);
                            }
                        } else if (full) {
                            min_keys4<true>(v0, m4, 0, 0);
#pragma unroll
                            for (int i = 0; i < 16; i += 2) m4[(i >> 1) & (NCH - 1)] = fminf(m4[(i >> 1) & (NCH - 1)], fminf(__uint_as_float(v3[i]), __uint_as_float(v3[i + 1])));
#pragma unroll
                            for (int i = 0; i < 8; i += 2) m4[(i >> 1) & (NCH - 1)] = fminf(m4[(i >> 1) & (NCH - 1)], fminf(__uint_as_float(v4[i]), __uint_as_float(v4[i + 1])));
                        } else {
                            min_keys4<false>(v0, m4, c0, nvalid);
#pragma unroll
                            for (int i = 0; i < 16; i++)
                                if (c0 + 32 + i < nvalid) m4[i & (NCH - 1)] = fminf(m4[i & (NCH - 1)], __uint_as_float(v3[i]));
#pragma unroll
                            for (int i = 0; i < 8; i++)
                                if (c0 + 48 + i < nvalid) m4[i & (NCH - 1)] = fminf(m4[i & (NCH - 1)], __uint_as_float(v4[i]));
                        }
                        continue;
                    }
#ifndef YAVO_TC_EXP_NO_LD
                    tmem_ld32(ta, v0);
                    tmem_ld32(ta + 32, v1);
                    tmem_wait2(v0, v1);
#endif
#ifndef YAVO_TC_EXP_NO_EPI
                    if (SECOND) {
                        if (full) min2_keys4<true>(v0, m4, s4, 32, 0, 0); else min2_keys4<false>(v0, m4, s4, 32, c0, nvalid);
                    } else {
                        if (full) min_keys4<true>(v0, m4, 0, 0); else min_keys4<false>(v0, m4, c0, nvalid);
                    }
                    if (CW == 112) tmem_ld32(ta + 64, v2);
                    if (SECOND) {
                        if (full) min2_keys4<true>(v1, m4, s4, 32, 0, 0); else min2_keys4<false>(v1, m4, s4, 32, c0 + 32, nvalid);
                    } else {
                        if (full) min_keys4<true>(v1, m4, 0, 0); else min_keys4<false>(v1, m4, c0 + 32, nvalid);
                    }
                    constexpr int TAIL = CW == 112 ? 16 : 8, TOFF = CW - TAIL;
                    if (CW == 112) {
                        tmem_ld16(ta + TOFF, v3);
                        tmem_wait2b(v2, v3);
                        if (SECOND) {
                            if (full) min2_keys4<true>(v2, m4, s4, 32, 0, 0); else min2_keys4<false>(v2, m4, s4, 32, c0 + 64, nvalid);
                        } else {
                            if (full) min_keys4<true>(v2, m4, 0, 0); else min_keys4<false>(v2, m4, c0 + 64, nvalid);
                        }
                    } else {
                        tmem_ld8(ta + TOFF, v3);
                        tmem_wait16(v3);
                    }
                    if (SECOND) {
                        uint32_t vt[32];
#pragma unroll
                        for (int i = 0; i < 32; i++) vt[i] = i < TAIL ? v3[i & 15] : 0u;
                        if (full) min2_keys4<true>(vt, m4, s4, TAIL, 0, 0); else min2_keys4<false>(vt, m4, s4, TAIL, c0 + TOFF, nvalid);
                    } else if (full) {
#pragma unroll
                        for (int i = 0; i < TAIL; i += 2) m4[(i >> 1) & (NCH - 1)] = fminf(m4[(i >> 1) & (NCH - 1)], fminf(__uint_as_float(v3[i]), __uint_as_float(v3[i + 1])));
                    } else {  // last tile of a train set
#pragma unroll
                        for (int i = 0; i < TAIL; i++)
                            if (c0 + TOFF + i < nvalid) m4[i & (NCH - 1)] = fminf(m4[i & (NCH - 1)], __uint_as_float(v3[i]));
                    }
#endif
                }
#ifndef YAVO_TC_EXP_NO_EPI
                if (DBG && dbg_acc && item == 0 && t == 0) {  // (test tool) re-read the tile
                    for (int c = 0; c < T4 / 16; c++) {
                        tmem_ld16(taddr + 16 * c, v3);
                        tmem_wait16(v3);
                        for (int i = 0; i < 16; i++) dbg_acc[row * T4 + 16 * c + i] = __uint_as_float(v3[i]);
                    }
                }
#endif
                float m = fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3]));
                if (NCH == 8) m = fminf(m, fminf(fminf(m4[NCH - 4], m4[NCH - 3]), fminf(m4[NCH - 2], m4[NCH - 1])));
                float m2 = 3.0e38f;
                if (SECOND) {
                    // second smallest key of the tile part: the smallest chain second, or the second smallest chain minimum
                    float a = m4[0], b = 3.0e38f;  // a <= b: the two smallest chain minima
#pragma unroll
                    for (int c = 1; c < NCH; c++) {
                        b = fminf(b, fmaxf(a, m4[c]));
                        a = fminf(a, m4[c]);
                    }
                    m2 = b;
#pragma unroll
                    for (int c = 0; c < NCH; c++) m2 = fminf(m2, s4[c]);
                }
                fence_before_sync();
                __syncwarp();
                if (lane == 0) bar_arrive(acc_empty + 8 * (acc));
                const int ki = (int)m + 32768;  // 256 * distance + column, exact
                const bool any = m < 1.0e30f;   // this warp's columns may all lie beyond the train set
                if (SECOND && any) {
                    // the running second: the old best when this tile part beats it, else this part's smallest; and its second
                    const int d1 = ki >> 8, d2 = m2 < 1.0e30f ? (((int)m2 + 32768) >> 8) : 0x7fffffff;
                    sec_d = min(sec_d, d1 < best_d ? min(best_d, d2) : d1);
                }
                if (any && (ki >> 8) < best_d) {
                    best_d = ki >> 8;
                    best_j = t * T4 + (ki & 255);
                }
            }
            t_cnt += n_tiles;
            int2 (*cb)[Q4] = comb[e_cnt & 1];
            int (*cs)[SECOND ? Q4 : 1] = comb_sec[SECOND ? (e_cnt & 1) : 0];
            e_cnt++;
            if (g >= 1) {
                cb[g - 1][row] = make_int2(best_d, best_j);
                if (SECOND) cs[g - 1][row] = sec_d;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI4_WARPS * 32) : "memory");
            if (g == 0) {
#pragma unroll
                for (int og = 0; og < EG - 1; og++) {
                    const int2 o = cb[og][row];
                    if (SECOND) {
                        // two parts (best, second): overall second = min(the two seconds, the larger of the two bests)
                        const int os = cs[og][row];
                        sec_d = min(min(sec_d, os), max(best_d, o.x));
                    }
                    if (o.x < best_d || (o.x == best_d && o.y < best_j)) {
                        best_d = o.x;
                        best_j = o.y;
                    }
                }
                const int q = q0 + row;
                if (q < nq) {
                    out_idx[(size_t)pair * out_stride + q] = best_j;  // empty train set: -1 / INT_MAX
                    out_dist[(size_t)pair * out_stride + q] = best_d;
                    if (SECOND) out_second[(size_t)pair * out_stride + q] = sec_d;
                }
            }
        }
    } else if (warp == MMA4_WARP) {
        YAVO_TC4_FOR_ITEMS
            // ------------------------------------------------ MMA issue: the whole warp runs the loop (uniform
            // control flow and operands), one elected lane issues tcgen05.mma / tcgen05.commit
            if (n_tiles > 0) {
                const uint32_t as = a_cnt & 1, aph = (a_cnt >> 1) & 1;
                const uint64_t adesc0 = smem_desc(sA + as * A4_BYTES, Q4 * 16u, 128u);
                const uint64_t adescx = smem_desc(sAX, Q4 * 16u, 128u), bdescx = smem_desc(sBX, T4 * 16u, 128u);
                const uint32_t sf1 = tmem_base + SF_ONE_COL, sf128 = tmem_base + SF_128_COL;
                bar_wait(a_full + 8 * (as), aph);
                for (int t = 0; t < n_tiles; t++, t_cnt++) {
                    const uint32_t s = t_cnt % NACC, ph = (t_cnt / NACC) & 1;    // accumulator
                    const uint32_t bs = t_cnt % NB4, bph = (t_cnt / NB4) & 1;    // operand stage
                    bar_wait(b_full + 8 * (bs), bph);
                    bar_wait(acc_empty + 8 * (s), ph ^ 1);
                    fence_after_sync();
                    const uint64_t bdesc0 = smem_desc(sB + bs * B4_BYTES, T4 * 16u, 128u);
                    const uint32_t tacc = tmem_base + s * T4;
                    if (elect_one()) {
#ifndef YAVO_TC_EXP_NO_MMA
#pragma unroll
                        for (int k = 0; k < ROWB / 32; k++)  // K = 64 nibbles = 32 operand bytes (two chunks) per instruction
                            mma_f4(tacc, adesc0 + (uint64_t)((k * 2 * Q4 * 16) >> 4), bdesc0 + (uint64_t)((k * 2 * T4 * 16) >> 4), sf1,
                                   sf128, k > 0);
                        mma_f4(tacc, adescx, bdescx, sf1, sf1, 1);  // + column index
#endif
                        mma_commit(b_empty + 8 * (bs));
                        mma_commit(acc_full + 8 * (s));
                        if (t == n_tiles - 1) mma_commit(a_empty + 8 * (as));
                    }
                    __syncwarp();
                }
                a_cnt++;
            }
        }
    } else if (warp == LOAD4_WARP) {
        YAVO_TC4_FOR_ITEMS
            // ------------------------------------------------ loader: packed bits of the train tiles -> ring
            if (n_tiles > 0 && lane == 0) {
                for (int t = 0; t < n_tiles; t++, t_cnt++) {
                    const uint32_t s = t_cnt % NR4, ph = (t_cnt / NR4) & 1;
                    const uint32_t bytes = 32u * (uint32_t)min(T4, nt - t * T4);
#ifndef YAVO_TC_EXP_NO_TMA
                    bar_wait(r_empty + 8 * (s), ph ^ 1);
                    bar_expect_tx(r_full + 8 * (s), bytes);
                    bulk_load(sRing + s * RING4_BYTES, dt + (size_t)t * T4 * 8, bytes, r_full + 8 * (s));
#endif
                }
            } else {
                t_cnt += n_tiles;
            }
        }
    } else if (warp < LOAD4_WARP) {
        YAVO_TC4_FOR_ITEMS
            // ------------------------------------------------ expanders (two groups on alternate operand stages)
            if (n_tiles > 0) {
                const int ew = warp - (EPI4_WARPS + 1), ge = ew >> 2, k = ew & 3;
                const uint32_t as = a_cnt & 1, aph = (a_cnt >> 1) & 1;
                const uint4 zero = make_uint4(0, 0, 0, 0);
                if (ge == (int)as) {  // query tile: rows k*32 .. +31
                    const int r = k * 32 + lane;
                    const bool in = q0 + r < nq;
                    const uint4 w0 = in ? __ldg(reinterpret_cast<const uint4 *>(dq + (size_t)(q0 + r) * 8)) : zero;
                    const uint4 w1 = in ? __ldg(reinterpret_cast<const uint4 *>(dq + (size_t)(q0 + r) * 8) + 1) : zero;
                    bar_wait(a_empty + 8 * (as), aph ^ 1);
                    expand_row4<false>(sA + as * A4_BYTES, Q4, r, w0, w1);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) bar_arrive(a_full + 8 * (as));
                }
                for (int t = 0; t < n_tiles; t++, t_cnt++) {
                    if (ge != (int)(t_cnt & 1)) continue;
                    const uint32_t s = t_cnt % NB4, ph = (t_cnt / NB4) & 1;    // operand stage
                    const uint32_t rs = t_cnt % NR4, rph = (t_cnt / NR4) & 1;  // ring entry
                    const int r = k * RPW + lane;  // train tile: rows k*RPW .. +RPW-1, as far as the tile goes
#ifndef YAVO_TC_EXP_NO_TMA
                    bar_wait(r_full + 8 * (rs), rph);
#endif
                    const bool in0 = r < T4 && t * T4 + r < nt, in1 = RPW > 32 && r + 32 < T4 && t * T4 + r + 32 < nt;
                    const uint8_t *src = gRing + rs * RING4_BYTES + r * 32;
                    const uint4 c0 = in0 ? *reinterpret_cast<const uint4 *>(src) : zero;
                    const uint4 c1 = in0 ? *reinterpret_cast<const uint4 *>(src + 16) : zero;
                    const uint4 c2 = in1 ? *reinterpret_cast<const uint4 *>(src + 1024) : zero;
                    const uint4 c3 = in1 ? *reinterpret_cast<const uint4 *>(src + 1040) : zero;
                    bar_wait(b_empty + 8 * (s), ph ^ 1);
#ifndef YAVO_TC_EXP_NO_EXP
                    if (r < T4) expand_row4<true>(sB + s * B4_BYTES, T4, r, c0, c1);
                    if (RPW > 32 && r + 32 < T4) expand_row4<true>(sB + s * B4_BYTES, T4, r + 32, c2, c3);
                    fence_async_smem();
#endif
                    __syncwarp();
                    if (lane == 0) {
#ifndef YAVO_TC_EXP_NO_TMA
                        bar_arrive(r_empty + 8 * (rs));  // only now: the ring reads above have certainly completed (their values were used)
#endif
                        bar_arrive(b_full + 8 * (s));
                    }
                }
                a_cnt++;
            }
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == MMA4_WARP) {
        fence_after_sync();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace tcm4
}  // namespace yavo
