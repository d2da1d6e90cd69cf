// match_tc.cuh — K5t: brute-force Hamming match on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Reference semantics: Brief::matchFeatures / hammingDistance (reference src/BriefDescriptor.cc:139-183):
// for every query descriptor the train descriptor with the smallest 256-bit Hamming distance, the LOWEST
// train index among equal minima.
//
// Formulation.  A query bit b is written as the FP8 (e4m3) number +1.0 (b = 0, byte 0x38) or -1.0 (b = 1, byte
// 0xB8), a train bit as -128.0 (b = 0, byte 0xF0) or +128.0 (b = 1, byte 0x70).  Over the 256 bit positions
//   sum_k a_k * b_k = -128 * (256 - 2 * hamming) = 256 * hamming - 32768,
// and a ninth K slice adds the train column: query bytes {1, 16, 0, ...} times train bytes {j & 15, j >> 4, 0, ...}
// = j, so the accumulator itself holds  key - 32768  with  key = 256 * hamming + j  — the (distance, index) pair
// whose minimum is the reference's "first minimum wins".  Every product and every partial sum is an integer of
// magnitude < 2^16, exact in the FP32 accumulator: this is bit-exact integer arithmetic carried out by
// tcgen05.mma (kind::f8f6f4, M = 128 queries x N = 256 train x K = 32 per instruction, 9 instructions per
// 128 x 256 tile) instead of 4.2 M XOR/POPC chains per frame pair, and the epilogue is a bare minimum.
//
// One persistent CTA per SM, three warp roles connected by mbarriers:
//   * one loader thread: cp.async.bulk (TMA engine) of the packed descriptors (32 bytes each) of the next operand
//     tiles into a small shared-memory ring, completion on an mbarrier;
//   * expander warps: read the packed bits from the ring and write the FP8 bytes into shared memory directly in the canonical K-major, non-swizzled UMMA operand layout
//     (8 x 16-byte core matrices: chunk c of row r at  c * rows * 16 + r * 16);  one SHF + one LOP3 per four
//     operand bytes.  The order of the 256 bit positions along K is permuted (bit 8b+s of word i sits at
//     k = 32 i + 4 s + b), identically for both operands, which a dot product does not see.
//   * one MMA thread: 8 tcgen05.mma per train tile into one of two 128 x 256 FP32 accumulators in TMEM,
//     tcgen05.commit onto the mbarriers that free the operand stage and publish the accumulator.
//   * four epilogue warps (one per TMEM lane quarter, thread = query row): tcgen05.ld 32 columns at a time,
//     one 3-input minimum per two keys on four independent chains; tiles are visited in ascending order and
//     only a strictly smaller distance replaces the best one, which extends the first-minimum rule across tiles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace yavo {
namespace tcm {

constexpr int TQ = 128;                 // queries per work item  (UMMA M)
constexpr int TT = 256;                 // train descriptors per tile (UMMA N)
constexpr int KBYTES = 256;             // operand bytes per descriptor (one FP8 per bit)
constexpr int A_BYTES = TQ * KBYTES;    // 32 KB
constexpr int B_BYTES = TT * KBYTES;    // 64 KB
constexpr int NSTAGE = 2;               // operand stages (A: per work item, B: per train tile) and accumulators
constexpr int EPI_WARPS = 8;            // warps 0..7: TMEM lane quarter == warp & 3, accumulator == warp >> 2
constexpr int MMA_WARP = EPI_WARPS;     // warp 8
constexpr int EXP_WARPS = 8;            // warps 9..16
constexpr int LOAD_WARP = EPI_WARPS + 1 + EXP_WARPS;  // warp 17: one thread issues the bulk copies of the packed bits
constexpr int THREADS = 32 * (EPI_WARPS + 2 + EXP_WARPS);
constexpr int NRING = 2;                // packed-descriptor ring: one entry = the bits of one operand tile
constexpr int RING_BYTES = TT * 32;     // 8 KB
constexpr int AX_BYTES = TQ * 32;        // constant ninth K slice of the queries  {1, 16, 0, ...}
constexpr int BX_BYTES = TT * 32;        // constant ninth K slice of a train tile {j & 15, j >> 4, 0, ...}
constexpr int SMEM_BYTES = NSTAGE * (A_BYTES + B_BYTES) + AX_BYTES + BX_BYTES + NRING * RING_BYTES;
constexpr uint32_t TMEM_COLS = 512;     // two 128 x 256 FP32 accumulators

__device__ __forceinline__ uint32_t saddr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr(bar)) : "memory");
}
#ifndef YAVO_TC_WAIT_HINT_NS
#define YAVO_TC_WAIT_HINT_NS 0
#endif
// mbarrier wait: try_wait suspends the thread in hardware until the phase completes or a time limit passes (with
// YAVO_TC_WAIT_HINT_NS > 0 that limit is given explicitly, so that waiting warps poll less often)
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity) {
#if YAVO_TC_WAIT_HINT_NS > 0
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(saddr(bar)),
        "r"(parity), "r"((uint32_t)YAVO_TC_WAIT_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(saddr(bar)),
        "r"(parity)
        : "memory");
#endif
}
__device__ __forceinline__ void bar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
// bulk asynchronous copy global -> shared (TMA engine), completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr(dst)),
                 "l"(src), "r"(bytes), "r"(saddr(bar))
                 : "memory");
}
// tcgen05.commit: the mbarrier receives one arrival when every tcgen05.mma issued so far by this thread is done
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy that tcgen05.mma reads operands through
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// The same on shared-window addresses computed once per kernel: taking the address of a __shared__ object inside the
// loops costs an S2R (SR_CgaCtaId) + LEA per barrier operation.
__device__ __forceinline__ void bar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle, descriptor version 1 (sm_100):
// bits [0,14) start address >> 4, [16,30) leading-dimension byte offset >> 4 (distance between the two 16-byte
// K chunks of one instruction), [32,46) stride byte offset >> 4 (distance between 8-row groups), [46,48) = 1.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor of kind::f8f6f4: D = F32 (bits [4,6) = 1), A = B = E4M3 (0), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(TT >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);

__device__ __forceinline__ void mma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
        : "memory");
}

#define YAVO_TM32(v) \
    "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
    "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), \
    "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), \
    "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define YAVO_TM32_RW(v) \
    "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), \
    "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), \
    "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), \
    "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])

// 32 consecutive accumulator columns of this thread's TMEM lane (asynchronous: see tmem_wait)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : YAVO_TM32(v)
        : "r"(taddr)
        : "memory");
}
// tcgen05.wait::ld; the registers are listed as read-write operands so that no use of them is scheduled above it
__device__ __forceinline__ void tmem_wait(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : YAVO_TM32_RW(v)::"memory");
}

// (a & b) | c  and  (a & b) ^ c  as ONE LOP3 each (two different immediates would cost the compiler two)
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t and_xor(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x6A;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Words 4h..4h+3 of one descriptor -> operand chunks 8h..8h+7 of row r of a tile with R rows.
// TRAIN = false: bytes +-1.0 (0x38 | sign);  TRAIN = true: bytes -+128.0 (0xF0 ^ sign).
template <bool TRAIN>
__device__ __forceinline__ void expand_half(uint32_t tile /* shared-window address */, int R, int r, int h, const uint4 &w, uint32_t msk, uint32_t cst) {
    uint32_t w0 = w.x, w1 = w.y, w2 = w.z, w3 = w.w;
    uint32_t p = tile + (uint32_t)(8 * h * R * 16 + r * 16);
    // a real loop over the four words: the stores of one word go out before the ALU work of the next one (fully
    // unrolled, ptxas gathers all 16 stores of a row at the end and every expander warp stalls on the store
    // queue at the same time)
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        uint32_t o[8];
#pragma unroll
        for (int s = 0; s < 8; s++) o[s] = TRAIN ? and_xor(w0 << (7 - s), msk, cst) : and_or(w0 << (7 - s), msk, cst);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p + R * 16), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
        p += 2 * R * 16;
        w0 = w1;
        w1 = w2;
        w2 = w3;
    }
}

// e4m3 byte of an integer 0..16 (exact: at most four significant bits)
__device__ __forceinline__ uint32_t e4m3_small_int(uint32_t n) {
    if (n == 0) return 0;
    const uint32_t e = 31 - __clz(n);
    return ((e + 7) << 3) | (((n << 3) >> e) & 7);
}

// running minima (four independent chains) over 32 accumulator columns that already hold key - 32768;
// FULL = false: col0 = first column of the chunk inside the tile, nvalid = train descriptors in this tile
template <bool FULL>
__device__ __forceinline__ void min_keys(const uint32_t (&v)[32], float (&m)[4], int col0, int nvalid) {
    if (FULL) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
#pragma unroll
            for (int c = 0; c < 4; c++) m[c] = fminf(m[c], fminf(__uint_as_float(v[i + 2 * c]), __uint_as_float(v[i + 2 * c + 1])));
        }
    } else {  // columns beyond the set do not take part
#pragma unroll
        for (int i = 0; i < 32; i++)
            if (col0 + i < nvalid) m[i & 3] = fminf(m[i & 3], __uint_as_float(v[i]));
    }
}

// Work item = (pair, tile of TQ queries).  Set addressing as in match_partial_kernel: pair p takes its queries
// from set p + q_set_offset and its train descriptors from set p + t_set_offset of arrays with set_stride_words
// words per set; n*_all == nullptr means every set holds n*_fixed descriptors.  out_* rows have out_stride entries.
// dbg_dots (test tool only): the 128 x 256 accumulator values (key - 32768) of the first tile of work item 0.
template <bool DBG>
__global__ void __launch_bounds__(THREADS, 1)
match_tc_kernel(const uint32_t *__restrict__ dq_all, const int *__restrict__ nq_all, int nq_fixed,
                const uint32_t *__restrict__ dt_all, const int *__restrict__ nt_all, int nt_fixed,
                size_t set_stride_words, int q_set_offset, int t_set_offset, int pairs, int q_tiles, int out_stride,
                int32_t *__restrict__ out_idx, int32_t *__restrict__ out_dist, float *__restrict__ dbg_dots) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t bars[6 * NSTAGE + 2 * NRING];
    __shared__ uint32_t tmem_base_s;
    __shared__ int2 comb[2][TQ];  // (distance, index) of accumulator group 1, double-buffered over work items
    uint32_t bars_s, smem_s;  // shared-window addresses, computed once (through an opaque move: no rematerialisation)
    asm volatile("mov.u32 %0, %2;\n\tmov.u32 %1, %3;" : "=r"(bars_s), "=r"(smem_s) : "r"(saddr(bars)), "r"(saddr(smem_raw)));
    const uint32_t a_full = bars_s, a_empty = a_full + 8 * NSTAGE, b_full = a_empty + 8 * NSTAGE, b_empty = b_full + 8 * NSTAGE;
    const uint32_t acc_full = b_empty + 8 * NSTAGE, acc_empty = acc_full + 8 * NSTAGE, r_full = acc_empty + 8 * NSTAGE, r_empty = r_full + 8 * NRING;
    const uint32_t sA = smem_s, sB = sA + NSTAGE * A_BYTES, sAX = sB + NSTAGE * B_BYTES, sBX = sAX + AX_BYTES, sRing = sBX + BX_BYTES;
    uint8_t *const gAX = smem_raw + NSTAGE * (A_BYTES + B_BYTES), *const gBX = gAX + AX_BYTES;  // generic pointers (set-up only)
    const uint8_t *const gRing = gBX + BX_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            bar_init(a_full + 8 * (s), EXP_WARPS / 2);
            bar_init(a_empty + 8 * (s), 1);
            bar_init(b_full + 8 * (s), EXP_WARPS / 2);
            bar_init(b_empty + 8 * (s), 1);
            bar_init(acc_full + 8 * (s), 1);
            bar_init(acc_empty + 8 * (s), EPI_WARPS / 2);
        }
        for (int s = 0; s < NRING; s++) {
            bar_init(r_full + 8 * (s), 1);
            bar_init(r_empty + 8 * (s), EXP_WARPS / 2);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // constant ninth K slice (two 16-byte chunks per row, the second one zero)
    for (int i = threadIdx.x; i < TQ + TT; i += THREADS) {
        const bool q = i < TQ;
        const int r = q ? i : i - TQ;
        uint8_t *p = q ? gAX + r * 16 : gBX + r * 16;
        const uint32_t w0 = q ? (0x38u | (0x58u << 8)) : (e4m3_small_int(r & 15) | (e4m3_small_int(r >> 4) << 8));
        *reinterpret_cast<uint4 *>(p) = make_uint4(w0, 0, 0, 0);
        *reinterpret_cast<uint4 *>(p + (q ? TQ : TT) * 16) = make_uint4(0, 0, 0, 0);
    }
    fence_async_smem();
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    const int items = pairs * q_tiles;
    uint32_t e_cnt = 0;             // work items seen by the epilogue warps
    uint32_t a_cnt = 0, t_cnt = 0;  // A stages used so far (work items with train tiles); train tiles so far

    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int pair = item / q_tiles, q0 = (item - pair * q_tiles) * TQ;
        const int nq = nq_all ? nq_all[pair + q_set_offset] : nq_fixed;
        const int nt = nt_all ? nt_all[pair + t_set_offset] : nt_fixed;
        if (q0 >= nq) continue;
        const int n_tiles = (nt + TT - 1) / TT;
        const uint32_t *dq = dq_all + (size_t)(pair + q_set_offset) * set_stride_words;
        const uint32_t *dt = dt_all + (size_t)(pair + t_set_offset) * set_stride_words;

        if (warp < EPI_WARPS) {
            // ------------------------------------------------ epilogue: thread = query row; warps 0-3 drain
            // accumulator 0 (every other tile), warps 4-7 accumulator 1, lane quarter = warp & 3
            const int g = warp >> 2, row = (warp & 3) * 32 + lane;
            int best_d = 0x7fffffff, best_j = -1;
            for (int t = ((t_cnt & 1) == (uint32_t)g) ? 0 : 1; t < n_tiles; t += 2) {
                const uint32_t ph = ((t_cnt + t) >> 1) & 1;
                bar_wait(acc_full + 8 * (g), ph);
                fence_after_sync();
                const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + g * TT;
                const int nvalid = min(TT, nt - t * TT);
                float m4[4] = {3.0e38f, 3.0e38f, 3.0e38f, 3.0e38f};
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr, v0);
                if (nvalid == TT) {
#pragma unroll 1
#if defined(YAVO_TC_EXP_EPI_HALF)
                    for (int c = 0; c < TT / 128; c++) {
#elif defined(YAVO_TC_EXP_NO_EPI)
                    for (int c = 0; c < 0; c++) {
#else
                    for (int c = 0; c < TT / 64; c++) {
#endif
                        tmem_wait(v0);
                        tmem_ld32(taddr + c * 64 + 32, v1);
                        if (DBG && dbg_dots && item == 0 && t == 0)
                            for (int i = 0; i < 32; i++) dbg_dots[row * TT + c * 64 + i] = __uint_as_float(v0[i]);
                        min_keys<true>(v0, m4, 0, 0);
                        tmem_wait(v1);
                        if (c + 1 < TT / 64) tmem_ld32(taddr + c * 64 + 64, v0);
                        if (DBG && dbg_dots && item == 0 && t == 0)
                            for (int i = 0; i < 32; i++) dbg_dots[row * TT + c * 64 + 32 + i] = __uint_as_float(v1[i]);
                        min_keys<true>(v1, m4, 0, 0);
                    }
                } else {  // last tile of a train set
#pragma unroll 1
                    for (int c = 0; c < TT / 32; c++) {
                        tmem_wait(v0);
                        min_keys<false>(v0, m4, c * 32, nvalid);
                        if (c + 1 < TT / 32) tmem_ld32(taddr + c * 32 + 32, v0);
                    }
                }
                const float m = fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3]));
                fence_before_sync();
                __syncwarp();
                if (lane == 0) bar_arrive(acc_empty + 8 * (g));
                const int ki = (int)m + 32768;  // 256 * distance + column, exact
                if ((ki >> 8) < best_d) {
                    best_d = ki >> 8;
                    best_j = t * TT + (ki & 255);
                }
            }
            t_cnt += n_tiles;
            // combine the two accumulator groups: smallest distance, then smallest index
            int2 *cb = comb[e_cnt & 1];
            e_cnt++;
            if (g == 1) cb[row] = make_int2(best_d, best_j);
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
            if (g == 0) {
                const int2 o = cb[row];
                if (o.x < best_d || (o.x == best_d && o.y < best_j)) {
                    best_d = o.x;
                    best_j = o.y;
                }
                const int q = q0 + row;
                if (q < nq) {
                    out_idx[(size_t)pair * out_stride + q] = best_j;  // empty train set: -1 / INT_MAX as the reference leaves it
                    out_dist[(size_t)pair * out_stride + q] = best_d;
                }
            }
        } else if (warp == MMA_WARP) {
            // ------------------------------------------------ MMA issue: one thread
            if (n_tiles > 0) {
                if (lane == 0) {
                    const uint32_t as = a_cnt & 1, aph = (a_cnt >> 1) & 1;
                    const uint64_t adesc0 = smem_desc(sA + as * A_BYTES, TQ * 16u, 128u);
                    const uint64_t adescx = smem_desc(sAX, TQ * 16u, 128u), bdescx = smem_desc(sBX, TT * 16u, 128u);
                    bar_wait(a_full + 8 * (as), aph);
                    for (int t = 0; t < n_tiles; t++, t_cnt++) {
                        const uint32_t s = t_cnt & 1, ph = (t_cnt >> 1) & 1;
                        bar_wait(b_full + 8 * (s), ph);
                        bar_wait(acc_empty + 8 * (s), ph ^ 1);
                        fence_after_sync();
                        const uint64_t bdesc0 = smem_desc(sB + s * B_BYTES, TT * 16u, 128u);
                        const uint32_t tacc = tmem_base + s * TT;
#pragma unroll
#if defined(YAVO_TC_EXP_MMA_HALF)
                        for (int k = 0; k < KBYTES / 64; k++)
#elif defined(YAVO_TC_EXP_NO_MMA)
                        for (int k = 0; k < 0; k++)
#else
                        for (int k = 0; k < KBYTES / 32; k++)  // K = 32 operand bytes (two 16-byte chunks) per instruction
#endif
                            mma_f8(tacc, adesc0 + (uint64_t)((k * 2 * TQ * 16) >> 4), bdesc0 + (uint64_t)((k * 2 * TT * 16) >> 4), k > 0);
#ifndef YAVO_TC_EXP_NO_MMA
                        mma_f8(tacc, adescx, bdescx, 1);  // + column index
#endif
                        mma_commit(b_empty + 8 * (s));
                        mma_commit(acc_full + 8 * (s));
                    }
                    mma_commit(a_empty + 8 * (as));
                } else {
                    t_cnt += n_tiles;
                }
                a_cnt++;
                __syncwarp();
            }
        } else if (warp == LOAD_WARP) {
            // ------------------------------------------------ loader: packed bits of the operand tiles -> ring
            // (the expanders must not have global loads in flight: their proxy fence waits for them)
            if (n_tiles > 0 && lane == 0) {
                for (int t = 0; t < n_tiles; t++, t_cnt++) {  // ring stage = operand stage = expander group
                    const uint32_t s = t_cnt & 1, ph = (t_cnt >> 1) & 1;
                    const uint32_t bytes = 32u * (uint32_t)min(TT, nt - t * TT);
                    bar_wait(r_empty + 8 * (s), ph ^ 1);
                    bar_expect_tx(r_full + 8 * (s), bytes);
                    bulk_load(sRing + s * RING_BYTES, dt + (size_t)t * TT * 8, bytes, r_full + 8 * (s));
                }
            } else {
                t_cnt += n_tiles;
            }
        } else {
            // ------------------------------------------------ expanders: packed bits -> FP8 operand bytes
            // Two groups of four warps take alternate tiles (group = operand stage), so that one group's proxy
            // fence (hundreds of cycles) runs under the other group's ALU work.
            if (n_tiles > 0) {
                const int ew = warp - (EPI_WARPS + 1), ge = ew >> 2, k = ew & 3;
                const uint32_t as = a_cnt & 1, aph = (a_cnt >> 1) & 1;
                const uint4 zero = make_uint4(0, 0, 0, 0);
                if (ge == (int)as) {  // query tile: rows k*32 .. +31, both halves (once per work item: plain loads)
                    const int r = k * 32 + lane;
                    const bool in = q0 + r < nq;
                    const uint4 w0 = in ? __ldg(reinterpret_cast<const uint4 *>(dq + (size_t)(q0 + r) * 8)) : zero;
                    const uint4 w1 = in ? __ldg(reinterpret_cast<const uint4 *>(dq + (size_t)(q0 + r) * 8) + 1) : zero;
                    bar_wait(a_empty + 8 * (as), aph ^ 1);
                    expand_half<false>(sA + as * A_BYTES, TQ, r, 0, w0, 0x80808080u, 0x38383838u);
                    expand_half<false>(sA + as * A_BYTES, TQ, r, 1, w1, 0x80808080u, 0x38383838u);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) bar_arrive(a_full + 8 * (as));
                }
                for (int t = 0; t < n_tiles; t++, t_cnt++) {
                    const uint32_t s = t_cnt & 1, ph = (t_cnt >> 1) & 1;
                    if (ge != (int)s) continue;
                    const uint32_t rs = s, rph = ph;
                    const int r = k * 64 + lane;  // train tile: rows k*64 .. +63
                    bar_wait(r_full + 8 * (rs), rph);
                    const bool in0 = t * TT + r < nt, in1 = t * TT + r + 32 < nt;
                    const uint8_t *src = gRing + rs * RING_BYTES + r * 32;
                    const uint4 c0 = in0 ? *reinterpret_cast<const uint4 *>(src) : zero;
                    const uint4 c1 = in0 ? *reinterpret_cast<const uint4 *>(src + 16) : zero;
                    const uint4 c2 = in1 ? *reinterpret_cast<const uint4 *>(src + 1024) : zero;
                    const uint4 c3 = in1 ? *reinterpret_cast<const uint4 *>(src + 1040) : zero;
                    bar_wait(b_empty + 8 * (s), ph ^ 1);
                    const uint32_t tile = sB + s * B_BYTES;
                    expand_half<true>(tile, TT, r, 0, c0, 0x80808080u, 0xF0F0F0F0u);
                    expand_half<true>(tile, TT, r, 1, c1, 0x80808080u, 0xF0F0F0F0u);
                    expand_half<true>(tile, TT, r + 32, 0, c2, 0x80808080u, 0xF0F0F0F0u);
                    expand_half<true>(tile, TT, r + 32, 1, c3, 0x80808080u, 0xF0F0F0F0u);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        bar_arrive(r_empty + 8 * (rs));  // only now: the ring reads above have certainly completed (their values were used)
                        bar_arrive(b_full + 8 * (s));
                    }
                }
                a_cnt++;
            }
        }
    }

    fence_before_sync();
    __syncthreads();
    if (warp == MMA_WARP) {
        fence_after_sync();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace tcm
}  // namespace yavo
