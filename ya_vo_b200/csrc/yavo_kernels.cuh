// yavo_kernels.cuh — sm_100a kernels of the YA_VO front end (detect+blur, compact+score, exact
// top-K select, BRIEF, integer-pipe Hamming match).  Byte/integer work (VABSDIFF4 / IDP.4A / IDP.2A / POPC / LOP3)
// staged through shared memory; the tensor-core matchers that replace K5 by default live in match_tc4.cuh /
// match_tc.cuh.  See DESIGN.md for the data layout and the roofline that bounds each kernel.
#pragma once
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fast_core.h"
#include "select_serial.h"
#include "blur_umma.cuh"

namespace yavo {

// ================================================================================================
// K0  re-pitch: dense uploaded pixels -> frame slots whose row pitch is a multiple of 128 bytes.
// Image::Image deep-copies the pixels (reference src/Image.cc:8-13); here the copy lands in HBM in
// the layout K1 reads with aligned word loads.  One thread per destination word.
// ================================================================================================
__global__ void repitch_kernel(const uint8_t *__restrict__ src, size_t src_pitch, size_t src_frame,
                               uint8_t *__restrict__ dst, int dst_pitch, size_t dst_frame, int cols) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (4 * w >= dst_pitch) return;
    const uint8_t *s = src + (size_t)blockIdx.z * src_frame + (size_t)blockIdx.y * src_pitch;
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        const int c = 4 * w + b;
        if (c < cols) v |= (uint32_t)__ldg(s + c) << (8 * b);
    }
    reinterpret_cast<uint32_t *>(dst + (size_t)blockIdx.z * dst_frame + (size_t)blockIdx.y * dst_pitch)[w] = v;
}

// ================================================================================================
// K1  detect + score + blur
// One CTA per 128x32 pixel tile of one frame.  The tile plus its halo is staged in shared memory by
// bulk asynchronous copies (cp.async.bulk, one 160-byte row each, completion on an mbarrier); frame
// rows are stored with a pitch that is a multiple of 128 so every staged row is 16-byte aligned.
// Three products come out of the single read of the pixels:
//   * the FAST corners of the tile (reference src/FastDetector.cc:298-320; 4 pixels per thread with
//     byte-SIMD), listed in tile row-major order,
//   * their Harris responses (:244-273) from the 5x5 windows that are already in shared memory —
//     scored densely, 256 corners per pass of the CTA.  The (response, position) entries go to a
//     per-frame pool (one atomicAdd per tile reserves the tile's span; pool order is irrelevant) and
//     a segment table [row][tile column] -> (pool offset, count) lets the select kernel read them
//     back in the reference's scan order without a corner bitmask or a second pass over the pixels,
//   * the 9x9 sigma-2.5 fixed-point Gaussian of the tile (reference src/BriefDescriptor.cc:90),
//     horizontal pass by IDP.4A into 16-bit vertical pairs, vertical pass by IDP.2A
// ================================================================================================
constexpr int TW = 128;            // tile width  (pixels)
constexpr int TH = 32;             // tile height (pixels)
constexpr int HALO = 4;
constexpr int SW = (TW + 2 * HALO) / 4;  // 34 words per staged row that the kernels read
constexpr int SLEAD = 16;                // staged rows start 16 bytes left of the tile (16-byte aligned for bulk copies)
constexpr int SROW = SLEAD + TW + 16;    // 160 bytes staged per row
constexpr int SROW_W = SROW / 4;         // 40 words
constexpr int SPX = SLEAD / 4;           // word index of the tile's first pixel inside a staged row
constexpr int SH = TH + 2 * HALO;        // 40 staged rows
constexpr int K1_THREADS = 256;
#ifndef YAVO_K1_MIN_CTAS
#define YAVO_K1_MIN_CTAS 8  // 32 registers per thread: eight CTAs per SM keep the issue slots of this issue-bound kernel full
#endif
#ifndef YAVO_K1_PREFETCH
#define YAVO_K1_PREFETCH 0  // tiles ahead (in launch order) whose pixels a CTA prefetches into L2; 0 = off
#endif
#ifndef YAVO_BLUR_UMMA
#define YAVO_BLUR_UMMA 0  // 0: IDP.4A / IDP.2A on the integer pipes; 1: both passes on the tensor cores (blur_umma.cuh: bit-exact, measured slower, see
                          // DESIGN.md); 2: hybrid — horizontal pass on the tensor cores, vertical pass from the accumulator registers
#endif
constexpr int K1_LIST = 1024;            // corners of a tile listed (and scored densely) per round; a tile with more takes more rounds
constexpr int SEG_CNT_BITS = 8;          // segment table entry = pool offset << 8 | count (count <= 128 per tile row)
constexpr int SEG_MAX_OFFSET = 1 << 24;  // pool offsets (and so max_cand) stay below this

// ---- bulk asynchronous copy (TMA engine, cp.async.bulk -> SASS UBLKCP) with an mbarrier for completion ----
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
// one tensor copy (TMA, cp.async.bulk.tensor -> SASS UTMALDG) of a box of a 3-D tensor (x = byte in row, y = row, z = frame slot);
// coordinates may lie outside the tensor, the out-of-range part of the box is zero-filled
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tmap, int x, int y, int z, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_addr(dst)),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(z), "r"(smem_addr(bar))
                 : "memory");
}
// L2 prefetch of a box of the 3-D tensor (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *tmap, int x, int y, int z) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y),
                 "r"(z)
                 : "memory");
}
// the same for a 4-D tensor map (16-byte chunk of a row, row, chunk index, frame slot): the box {16, rows, chunks, 1}
// lands in shared memory as [chunk][row][16 bytes], the K-major core-matrix order tcgen05.mma reads (blur_umma.cuh)
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tmap, int c0, int c1, int c2, int c3, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     smem_addr(dst)),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
}

template <bool DO_FAST, bool DO_BLUR>
__global__ void __launch_bounds__(K1_THREADS, YAVO_K1_MIN_CTAS)
detect_blur_kernel(const __grid_constant__ CUtensorMap frames_map, const __grid_constant__ CUtensorMap frames_cmap, int slot_base,
                   const uint8_t *__restrict__ frames, size_t frame_stride, int pitch, int H, int W,
                   uint8_t *__restrict__ blur, yavo_ent *__restrict__ pool, int max_cand, int *__restrict__ ncand,
                   uint32_t *__restrict__ seg, int seg_cols, int rows_alloc, const uint8_t *__restrict__ blur_consts) {
    __shared__ __align__(128) uint32_t tile[SH][SROW_W];  // staged rows: cols x0-16 .. x0+143
    __shared__ __align__(8) uint64_t tile_bar;
#if YAVO_BLUR_UMMA
    constexpr bool UMMA = DO_BLUR;
    __shared__ __align__(128) uint8_t ub[UMMA ? bu::BU_UB_BYTES : 16];      // the staged rows in core-matrix order
    __shared__ __align__(128) uint8_t bcst[UMMA ? bu::BU_CONST_BYTES : 16];  // the two constant band matrices
    __shared__ __align__(8) uint64_t bu_bars[2];
    __shared__ uint32_t tmem_base_s;
    if (UMMA) bu::bu_prologue(bcst, blur_consts, bu_bars, &tmem_base_s);
#else
    constexpr bool UMMA = false;
    __shared__ uint4 hpair[SH / 2][TW / 4];        // [row pair][quad] -> 4 x (h[even] | h[odd] << 16)
#endif
    // corners of the tile: per-row bit masks and counts, the row-major corner list, the tile's span in the pool
    __shared__ uint32_t rmask[TH][TW / 32];
    __shared__ int rcnt[TH];
    __shared__ uint16_t clist[DO_FAST ? K1_LIST : 1];
    __shared__ int s_base, s_total;

    const int f = blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const uint8_t *img = frames + (size_t)f * frame_stride;
    const int tid = threadIdx.x;

    // ---- stage tile rows y0-4 .. y0+35 (reflected into the image), cols x0-16 .. x0+143 -------------
    // Tiles whose 40 staged rows all lie inside the image (every tile row but the first and the last): ONE tensor copy
    // (TMA) of the 160 x 40 byte box, issued and awaited by thread 0; columns left of the image are zero-filled and not
    // used (the blur's reflected columns are patched in below).  First / last tile row: one bulk copy per staged row
    // (reflected row index), issued by the lanes of warp 0.  Either way the other warps park at the CTA barrier
    // below (a barrier stall issues nothing; a try_wait loop in every warp would take issue slots this issue-bound
    // kernel needs).
    if (tid < 32) {
        if (tid == 0) mbar_init(&tile_bar, 1);
        if (y0 - HALO >= 0 && y0 + TH + HALO <= H) {  // CTA-uniform
            if (tid == 0) {
#if YAVO_BLUR_UMMA == 2
                // + the same pixels once more, in core-matrix order, as the tensor cores' operand (unless reflected
                // columns have to be patched in first: then the CTA re-lays the patched tile out itself)
                const bool direct = UMMA && !(x0 == 0 || x0 + TW + HALO > W);
                mbar_expect_tx(&tile_bar, direct ? 2 * SROW * SH : SROW * SH);
                if (direct) tma_load_4d(ub, &frames_cmap, 0, y0 - HALO, (x0 - SLEAD) / 16, slot_base + f, &tile_bar);
#else
                mbar_expect_tx(&tile_bar, SROW * SH);
#endif
                tma_load_3d(&tile[0][0], &frames_map, x0 - SLEAD, y0 - HALO, slot_base + f, &tile_bar);
            }
        } else {
            const int src_col = max(x0 - SLEAD, 0);
            const int dst_off = src_col - (x0 - SLEAD);                    // 16 for the first tile column, else 0
            const uint32_t row_bytes = (uint32_t)min(SROW - dst_off, pitch - src_col);
            if (tid == 0) mbar_expect_tx(&tile_bar, row_bytes * SH);
            __syncwarp();
            for (int tr = tid; tr < SH; tr += 32) {
                const int gr = reflect101(y0 - HALO + tr, H);
                bulk_g2s(reinterpret_cast<uint8_t *>(&tile[tr][0]) + dst_off, img + (size_t)gr * pitch + src_col, row_bytes,
                         &tile_bar);
            }
        }
#if YAVO_K1_PREFETCH > 0
        // the frames come from HBM (a batch is far larger than L2): pull the tile of a CTA that will start a few waves
        // from now into L2, so that its staging wait is an L2 round trip
        if (tid == 0) {
            const int gx = (int)gridDim.x, gy = (int)gridDim.y;
            const int lin = (int)blockIdx.x + gx * ((int)blockIdx.y + gy * (int)blockIdx.z) + YAVO_K1_PREFETCH;
            const int t = lin / gx, fx = lin - t * gx, fz = t / gy, fy = t - fz * gy;
            if (fz < (int)gridDim.z) tma_prefetch_3d(&frames_map, fx * TW - SLEAD, fy * TH - HALO, slot_base + fz);
        }
#endif
        if (tid == 0) mbar_wait(&tile_bar, 0);
    }
    const bool edge_cols = DO_BLUR && (x0 == 0 || x0 + TW + HALO > W);
    if (edge_cols) {
        __syncthreads();
        // BORDER_REFLECT_101 for the columns outside [0,W) that valid outputs can read
        // (x0-4..-1 on the left, W..W+3 on the right)
        uint8_t *tb = reinterpret_cast<uint8_t *>(&tile[0][0]);
        for (int i = tid; i < SH * 8; i += K1_THREADS) {
            const int tr = i >> 3, j = i & 7;
            const int gc = (j < 4) ? (j - 4) : (W + j - 4);
            const int tc = gc - (x0 - SLEAD);
            if (tc < 0 || tc >= SROW) continue;
            if (j < 4 && x0 != 0) continue;
            const int gr = reflect101(y0 - HALO + tr, H);
            tb[tr * SROW + tc] = __ldg(img + (size_t)gr * pitch + reflect101(gc, W));
        }
    }
    __syncthreads();

    const int lane = tid & 31, warp = tid >> 5;

#if YAVO_BLUR_UMMA
    uint32_t tmem_base = 0;
    if (UMMA) {
#if YAVO_BLUR_UMMA == 2
        const bool direct = (y0 - HALO >= 0 && y0 + TH + HALO <= H) && !edge_cols;  // CTA-uniform: the operand arrived by tensor copy
#else
        const bool direct = false;
#endif
        if (!direct) {
            bu::bu_relayout(reinterpret_cast<const uint8_t *>(&tile[0][0]), ub);
            __syncthreads();
        }
        tmem_base = tmem_base_s;
        bu::bu_pass1_issue(ub, bcst, bu_bars, tmem_base);  // the horizontal pass runs under the segment test
    }
#else
    if (DO_BLUR) {
        // horizontal pass: (SH/2) row pairs x 32 quads; thread -> one quad of one row pair
        // (three rounds written out — the third for half of the threads: the loop's address increments become immediates)
#pragma unroll
        for (int it = 0; it < ((SH / 2) * (TW / 4) + K1_THREADS - 1) / K1_THREADS; it++) {
            const int i = tid + it * K1_THREADS;
            if (i >= (SH / 2) * (TW / 4)) break;
            const int rp = i >> 5, q = i & 31;
            const uint32_t *ra = &tile[2 * rp][q + SPX];
            const uint32_t *rb = &tile[2 * rp + 1][q + SPX];
            uint32_t ha[4], hb[4];
            yavo_blur_h4(ra[-1], ra[0], ra[1], ha);
            yavo_blur_h4(rb[-1], rb[0], rb[1], hb);
            // sums < 2^16: the low halves of the two rows' sums side by side, one PRMT per pair
            hpair[rp][q] = make_uint4(__byte_perm(ha[0], hb[0], 0x5410), __byte_perm(ha[1], hb[1], 0x5410),
                                      __byte_perm(ha[2], hb[2], 0x5410), __byte_perm(ha[3], hb[3], 0x5410));
        }
    }
#endif

    if (DO_FAST) {
        // Segment test, warp -> 4 tile rows, lane -> quad (4 pixels).  Pass A runs the cheap necessary
        // condition (ring 0, 4, 7, 8 all differ) on every quad and packs the survivors' addresses into a
        // per-warp list; pass B runs the full 16-position test only on the survivors, densely packed 32 to a
        // warp step (on real frames a few percent of the quads); pass C assembles the 32-pixel mask words.
        constexpr int NW = K1_THREADS / 32, RPW = TH / NW;
        __shared__ __align__(16) uint8_t fnib[NW][RPW][32];
        __shared__ uint8_t flist[NW][RPW * 32];
        __shared__ uint32_t fcore[NW][RPW * 32];  // the survivors' necessary-condition words (pass B starts from them)
        const unsigned lt = (1u << lane) - 1u;
        int cnt = 0;
        // the warp's four rows of nibbles start at zero (one 16-byte store by each of eight lanes; pass B, behind a
        // __syncwarp, fills in the survivors)
        if (lane < RPW * 32 / 16) reinterpret_cast<uint4 *>(&fnib[warp][0][0])[lane] = make_uint4(0u, 0u, 0u, 0u);
        // rows 4 .. H-5 only: a tile whose 32 rows all qualify (every tile row but the first and the last) skips the check
        const bool rows_inside = y0 >= 4 && y0 + TH <= H - 4;  // CTA-uniform
        auto pass_a = [&](int k, bool check_row = true) {
            const int tr = warp + NW * k, gr = y0 + tr, sr = tr + HALO;
            uint32_t core = 0;  // the word itself, tested after the (warp-uniform) branch: no boolean to materialise
            if (!check_row || (gr >= 4 && gr < H - 4))
                core = yavo_fast4_core(&tile[sr][lane + SPX], &tile[sr + 1][lane + SPX], &tile[sr + 3][lane + SPX]);
            const bool live = core != 0;
            const unsigned bl = __ballot_sync(0xffffffffu, live);
            if (live) {
                const int slot = cnt + __popc(bl & lt);
                flist[warp][slot] = (uint8_t)(k * 32 + lane);
                fcore[warp][slot] = core;
            }
            cnt += __popc(bl);
        };
        // With the blur on the tensor cores its steps are interleaved with the balanced parts of the segment test, so
        // that every tcgen05.mma round trip has a few hundred instructions per warp to hide under.
#if YAVO_BLUR_UMMA == 0
        if (rows_inside) {
            pass_a(0, false);
            pass_a(1, false);
            pass_a(2, false);
            pass_a(3, false);
        } else {
            pass_a(0);
            pass_a(1);
            pass_a(2);
            pass_a(3);
        }
#else
        pass_a(0);
        pass_a(1);
#if YAVO_BLUR_UMMA == 1
        if (UMMA) bu::bu_pass1_drain(bcst, bu_bars, tmem_base);  // row sums -> byte operands, first half of the vertical pass issued
#endif
        pass_a(2);
        pass_a(3);
#endif
        static_assert(RPW == 4, "pass A is written out for four rows per warp");
#if YAVO_BLUR_UMMA == 1
        if (UMMA) bu::bu_pass2_drain(bcst, bu_bars, tmem_base, ub, 0);  // first half drained, second half issued
#elif YAVO_BLUR_UMMA == 2
        if (UMMA) bu::bu_vpass_from_tmem(bu_bars, tmem_base, ub, 0);  // the horizontal pass ran under pass A; vertical pass from its accumulator
#endif
        __syncwarp();
        for (int i = lane; i < cnt; i += 32) {
            const int e = flist[warp][i], k = e >> 5, q = e & 31;
            const int sr = warp + NW * k + HALO;
            uint32_t nib = yavo_fast4_rest(&tile[sr - 3][q + SPX], &tile[sr - 2][q + SPX], &tile[sr - 1][q + SPX],
                                           &tile[sr][q + SPX], &tile[sr + 1][q + SPX], &tile[sr + 2][q + SPX],
                                           &tile[sr + 3][q + SPX], fcore[warp][i]);
            if (x0 < 4 || x0 + TW > W - 4) {  // CTA-uniform: only the first / last tile column holds excluded columns
#pragma unroll
                for (int bb = 0; bb < 4; bb++)  // interior columns only: 4 <= col < W-4
                    if (x0 + 4 * q + bb < 4 || x0 + 4 * q + bb >= W - 4) nib &= ~(1u << bb);
            }
            fnib[warp][k][q] = (uint8_t)nib;
        }
        __syncwarp();
        // pass C: lane l < 4*RPW packs the eight nibbles of mask word (row l/4, word l%4) in one go; the words and the
        // per-row corner counts stay in shared memory
        if (lane < 4 * RPW) {
            const int k = lane >> 2, w = lane & 3;
            const uint2 nb = *reinterpret_cast<const uint2 *>(&fnib[warp][k][8 * w]);  // 8 bytes, low nibble used
            // [n0,n1,n2,n3] -> n0 | n1<<4 in byte 0, n2 | n3<<4 in byte 2; then gather the four packed bytes
            const uint32_t lo = (nb.x & 0x000f000fu) | ((nb.x & 0x0f000f00u) >> 4);
            const uint32_t hi = (nb.y & 0x000f000fu) | ((nb.y & 0x0f000f00u) >> 4);
            const uint32_t v = __byte_perm(lo, hi, 0x6420);
            int c = __popc(v);
            rmask[warp + NW * k][w] = v;
            c += __shfl_xor_sync((1u << (4 * RPW)) - 1u, c, 1);
            c += __shfl_xor_sync((1u << (4 * RPW)) - 1u, c, 2);
            if (w == 0) rcnt[warp + NW * k] = c;
        }
    }

    // Corner list of the tile in row-major order (tile row t, then column): thread j < 128 owns mask word (row j/4,
    // word j%4) and writes its corners at [start of row + corners in the earlier words of the row].  Entries
    // [round * K1_LIST, (round+1) * K1_LIST) are kept; every one of the four warps computes the row starts itself.
    auto build_list = [&](int round, int *total_out, int *row_start_out, int *row_cnt_out) {
        const int lane = tid & 31;
        const int c = rcnt[lane];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        *total_out = __shfl_sync(0xffffffffu, incl, 31);
        *row_start_out = incl - c;  // of tile row `lane`
        *row_cnt_out = c;
        const int t = tid >> 2, w = tid & 3;
        uint32_t m = rmask[t][w];
        const int cw = __popc(m);
        int pre = cw;  // inclusive prefix over the row's four words (lanes 4i .. 4i+3)
        int u = __shfl_up_sync(0xffffffffu, pre, 1, 4);
        if (w >= 1) pre += u;
        u = __shfl_up_sync(0xffffffffu, pre, 2, 4);
        if (w >= 2) pre += u;
        int off = __shfl_sync(0xffffffffu, incl - c, t & 31) + pre - cw - round * K1_LIST;
        while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            if (off >= 0 && off < K1_LIST) clist[off] = (uint16_t)((t << 7) | (w * 32 + b));
            off++;
        }
    };

#if YAVO_BLUR_UMMA
    if (UMMA) {
#if YAVO_BLUR_UMMA == 1
        if (!DO_FAST) {
            bu::bu_pass1_drain(bcst, bu_bars, tmem_base);
            bu::bu_pass2_drain(bcst, bu_bars, tmem_base, ub, 0);
        }
        // second half of the vertical pass drained (its CTA barrier also publishes the masks and counts of the segment
        // test), TMEM freed, the blurred tile written
        bu::bu_pass2_drain(bcst, bu_bars, tmem_base, ub, 1);
        // (the CTA barrier inside also publishes the masks and counts of the segment test)
        bu::bu_finish(tmem_base, ub, blur + (size_t)f * frame_stride, pitch, H, x0, y0);
#else
        if (!DO_FAST) {
            bu::bu_vpass_from_tmem(bu_bars, tmem_base, ub, 0);
            bu::bu_vpass_from_tmem(bu_bars, tmem_base, ub, 1);
            bu::bu_finish(tmem_base, ub, blur + (size_t)f * frame_stride, pitch, H, x0, y0);
        }
#endif
    }
#endif
    if (DO_FAST) {
        if (YAVO_BLUR_UMMA != 1 || !UMMA) __syncthreads();  // masks, counts (and the horizontal blur pass) are complete
        if (tid < 128) {
            int total, row_start, row_cnt;
            build_list(0, &total, &row_start, &row_cnt);
            if (tid < 32) {
                // reserve the tile's span of the frame's pool; publish where each of the tile's rows starts in it
                int base = 0;
                if (tid == 0 && total > 0) base = atomicAdd(&ncand[blockIdx.z], total);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (tid == 0) {
                    s_base = base;
                    s_total = total;
                }
                const int gr = y0 + tid;
                if (gr < H)
                    seg[((size_t)f * rows_alloc + gr) * seg_cols + blockIdx.x] =
                        ((uint32_t)(base + row_start) << SEG_CNT_BITS) | (uint32_t)row_cnt;
            }
        }
    }

#if YAVO_BLUR_UMMA == 2
    if (UMMA && DO_FAST) {
        // second half of the vertical pass behind the pool reservation (its atomic's round trip is hidden here, as it is
        // behind the vertical pass of the integer-pipe kernel); then TMEM freed and the tile written.  The CTA barrier
        // inside also publishes s_base / s_total for the scoring below.
        bu::bu_vpass_from_tmem(bu_bars, tmem_base, ub, 1);
        bu::bu_finish(tmem_base, ub, blur + (size_t)f * frame_stride, pitch, H, x0, y0);
    }
#endif
#if !YAVO_BLUR_UMMA
    if (DO_BLUR) {
        if (!DO_FAST) __syncthreads();
        // vertical pass: output row pairs (tile rows 2j, 2j+1) x 32 quads
        for (int i = tid; i < (TH / 2) * (TW / 4); i += K1_THREADS) {
            const int j = i >> 5, q = i & 31;
            // output tile row t = 2j -> staged row s = t + 4; taps s-4 .. s+5 = staged rows 2j .. 2j+9
            uint4 p[5];
#pragma unroll
            for (int k = 0; k < 5; k++) p[k] = hpair[j + k][q];
            uint32_t o0[4], o1[4];  // rounded fixed-point sums: the output pixel is byte 2 of each
            {
                const uint32_t P[5] = {p[0].x, p[1].x, p[2].x, p[3].x, p[4].x};
                yavo_blur_v2_raw(P, &o0[0], &o1[0]);
            }
            {
                const uint32_t P[5] = {p[0].y, p[1].y, p[2].y, p[3].y, p[4].y};
                yavo_blur_v2_raw(P, &o0[1], &o1[1]);
            }
            {
                const uint32_t P[5] = {p[0].z, p[1].z, p[2].z, p[3].z, p[4].z};
                yavo_blur_v2_raw(P, &o0[2], &o1[2]);
            }
            {
                const uint32_t P[5] = {p[0].w, p[1].w, p[2].w, p[3].w, p[4].w};
                yavo_blur_v2_raw(P, &o0[3], &o1[3]);
            }
            // byte 2 of four sums -> one word: two 2-way gathers, then interleave
            const uint32_t w0 = __byte_perm(__byte_perm(o0[0], o0[1], 0x0062), __byte_perm(o0[2], o0[3], 0x0062), 0x5410);
            const uint32_t w1 = __byte_perm(__byte_perm(o1[0], o1[1], 0x0062), __byte_perm(o1[2], o1[3], 0x0062), 0x5410);
            const int gr = y0 + 2 * j, gc = x0 + 4 * q;
            // (x0 + 4 q < pitch always: the pitch is a multiple of the tile width and covers every tile column)
            uint8_t *dst = blur + (size_t)f * frame_stride + (size_t)gr * pitch + gc;
            if (y0 + TH <= H) {  // CTA-uniform: every row of the tile lies inside the frame
                *reinterpret_cast<uint32_t *>(dst) = w0;
                *reinterpret_cast<uint32_t *>(dst + pitch) = w1;
            } else {
                if (gr < H) *reinterpret_cast<uint32_t *>(dst) = w0;
                if (gr + 1 < H) *reinterpret_cast<uint32_t *>(dst + pitch) = w1;
            }
        }
    }
#endif

    if (DO_FAST) {
        // Harris response of every corner from the staged pixels (the kernel's unbalanced tail: a tile's ~45 corners
        // occupy two warps) (reference src/FastDetector.cc:244-273; the 5x5 window
        // of an interior pixel lies inside the tile + halo), densely: thread i takes corner i of the list
        if (!(YAVO_BLUR_UMMA == 2 && UMMA)) __syncthreads();
        const int total = s_total, base = s_base;
        const bool fits = base + total <= max_cand;  // otherwise the select kernel reports the overflow (ncand > max_cand)
        const uint8_t *tb = reinterpret_cast<const uint8_t *>(&tile[0][0]);
        for (int r0 = 0; r0 < total; r0 += K1_LIST) {
            if (r0 > 0) {  // more than K1_LIST corners in one tile: list the next K1_LIST (CTA-uniform, rare)
                __syncthreads();
                if (tid < 128) {
                    int a, b, c;
                    build_list(r0 / K1_LIST, &a, &b, &c);
                }
                __syncthreads();
            }
            const int n = min(K1_LIST, total - r0);
            for (int i = tid; i < n; i += K1_THREADS) {
                const int e = clist[i], t = e >> 7, c = e & 127;
                const uint8_t *ctr = tb + (t + HALO) * SROW + SLEAD + c;
                int a, bb, cc;
                yavo_structure_tensor([&](int dr, int dc) { return (int)ctr[dr * SROW + dc]; }, 0, 0, &a, &bb, &cc);
                if (fits)
                    pool[(size_t)f * max_cand + base + r0 + i] =
                        yavo_make_ent(yavo_harris_from_tensor(a, bb, cc), ((uint32_t)(y0 + t) << 16) | (uint32_t)(x0 + c));
            }
        }
    }
}

// ================================================================================================
// K2  candidate gather (device function + stand-alone kernel)
// The detect kernel leaves a frame's corners in its pool in no particular tile order; the segment table
// [row][tile column] -> (pool offset, count) gives the reference's scan order back (row-major over the image,
// src/FastDetector.cc:298-324): an exclusive scan of the counts over the table in row-major order is each
// segment's position in the candidate list.  The select kernel calls this while it loads its sort buffer; the
// stand-alone kernel serves yavo_fast_candidates.  All threads of the CTA take part.
// ================================================================================================
// Latency matters here (the select kernel cannot start before its list is loaded, and a thread's segments are few):
// a thread's segment entries are loaded GC at a time, and the first two pool entries of each of those segments are in
// flight together before any of them is stored (a tile row holds 1.4 corners on average on the benchmark's frames;
// longer segments finish in a plain loop).
constexpr int GC = 8;

template <int NTHREADS>
__device__ int gather_candidates(const uint32_t *__restrict__ seg_f, int seg_cols, int H, int ntx,
                                 const yavo_ent *__restrict__ pool_f, yavo_ent *dst, int dst_cap, int *s_wtot /* NTHREADS / 32 ints */) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = H * ntx, per = (S + NTHREADS - 1) / NTHREADS;
    const int e0 = min(S, tid * per), e1 = min(S, e0 + per);
    constexpr uint32_t CNT_MASK = (1u << SEG_CNT_BITS) - 1u;
    int mine = 0;
    {
        int row = e0 / ntx, tx = e0 - row * ntx;
        for (int b = e0; b < e1; b += GC) {
            uint32_t sg[GC];
#pragma unroll
            for (int j = 0; j < GC; j++) {
                sg[j] = b + j < e1 ? seg_f[(size_t)row * seg_cols + tx] : 0u;
                if (++tx == ntx) { tx = 0; row++; }
            }
#pragma unroll
            for (int j = 0; j < GC; j++) mine += (int)(sg[j] & CNT_MASK);
        }
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    int pre = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NTHREADS / 32; w++) {
        const int t = s_wtot[w];
        if (w < warp) pre += t;
        total += t;
    }
    if (total <= dst_cap) {
        int o = pre + incl - mine;
        int row = e0 / ntx, tx = e0 - row * ntx;
        for (int b = e0; b < e1; b += GC) {
            uint32_t sg[GC];
#pragma unroll
            for (int j = 0; j < GC; j++) {
                sg[j] = b + j < e1 ? seg_f[(size_t)row * seg_cols + tx] : 0u;
                if (++tx == ntx) { tx = 0; row++; }
            }
            yavo_ent v0[GC], v1[GC];
#pragma unroll
            for (int j = 0; j < GC; j++) {
                const int cnt = (int)(sg[j] & CNT_MASK);
                const yavo_ent *src = pool_f + (sg[j] >> SEG_CNT_BITS);
                v0[j] = cnt > 0 ? src[0] : 0ull;
                v1[j] = cnt > 1 ? src[1] : 0ull;
            }
#pragma unroll
            for (int j = 0; j < GC; j++) {
                const int cnt = (int)(sg[j] & CNT_MASK);
                if (cnt > 0) dst[o] = v0[j];
                if (cnt > 1) dst[o + 1] = v1[j];
                if (cnt > 2) {
                    const yavo_ent *src = pool_f + (sg[j] >> SEG_CNT_BITS);
                    for (int k = 2; k < cnt; k++) dst[o + k] = src[k];
                }
                o += cnt;
            }
        }
    }
    __syncthreads();
    return total;
}

constexpr int K2_THREADS = 512;

// one CTA per frame: cand[f][0 .. ncand) = the frame's candidates in scan order
__global__ void __launch_bounds__(K2_THREADS)
gather_kernel(const uint32_t *__restrict__ seg, int seg_cols, int rows_alloc, int H, int ntx,
              const yavo_ent *__restrict__ pool, yavo_ent *__restrict__ cand, int max_cand, const int *__restrict__ ncand) {
    __shared__ int wtot[K2_THREADS / 32];
    const int f = blockIdx.x;
    if (ncand[f] > max_cand) return;  // overflowed pool: the caller reports it
    gather_candidates<K2_THREADS>(seg + (size_t)f * rows_alloc * seg_cols, seg_cols, H, ntx, pool + (size_t)f * max_cand,
                                  cand + (size_t)f * max_cand, max_cand, wtot);
}

// ================================================================================================
// K3  exact top-K select  (std::sort replay; see select_serial.h)
// One CTA per frame.  Replays libstdc++'s introsort partition tree over the candidate list (in scan
// order), pruned to ranges that start below K, so the first K outputs come out in exactly the order
// the reference's std::sort leaves them — tied responses included.
//   phase 1  ranges larger than SEL_WARP_MAX: partitioned by the whole CTA, level by level;
//   phase 2  everything else: a shared work queue of ranges served by the 16 warps independently.
//            A warp partitions its range, queues the right child and keeps the left; ranges of
//            <= 16 elements (where std::sort's final insertion sort is the only thing left, i.e. a
//            stable sort) are finished by a warp-wide rank sort.
// A partition is the parallel form of __unguarded_partition: the i-th element from the left that
// does not sort before the pivot (L_i) is swapped with the i-th element from the right that the
// pivot does not sort before (R_i) while L_i < R_i; with m such swaps the cut is
// min(L_{m+1}, R_m).  Stopper positions are found by rank (ballot + scan) and scattered into a
// scratch list; the swaps are independent.  The depth-limit fallback (heapsort) is replayed by a
// single lane (never reached on real score lists; kept for exactness).
// ================================================================================================
#ifndef YAVO_SEL_THREADS
#define YAVO_SEL_THREADS 384  // measured on B200 (round 2, scores arriving from K1): 384 x 2 CTAs/SM 0.386 ms per 1024 frames, 512 x 2 0.398, 256 x 3 0.424, 256 x 4 0.433
#endif
#ifndef YAVO_SEL_SLEEP
#define YAVO_SEL_SLEEP 64
#endif
constexpr int SEL_THREADS_BATCH = YAVO_SEL_THREADS;  // CTA size of the batch instance (many frames in flight, two CTAs per SM)
constexpr int SEL_THREADS_SINGLE = 768;              // the single-frame instance: one CTA on an idle GPU, twice the warps for phase 2
#ifndef YAVO_SEL_SMEM_ENTS
#define YAVO_SEL_SMEM_ENTS 6144  // measured on B200: 6144 entries at 2 CTAs/SM beat 4096@3, 4096@2 and 8192@2 (leaves L1 for the scoring loads)
#endif
#ifndef YAVO_SEL_MIN_CTAS
#define YAVO_SEL_MIN_CTAS 2
#endif
constexpr int SEL_SMEM_ENTS = YAVO_SEL_SMEM_ENTS;  // candidates kept in shared memory once the active prefix fits
#ifndef YAVO_SEL_WARP_MAX
#define YAVO_SEL_WARP_MAX 1024  // measured on B200 (1024 frames): 512 0.389 ms, 768 0.371, 1024 0.365, 1536 0.361 (no longer two CTAs' worth of shared memory), 2048 0.53
#endif
constexpr int SEL_WARP_MAX = YAVO_SEL_WARP_MAX;  // ranges up to this size are partitioned by one warp
constexpr int SEL_QCAP = 512;           // shared work queue (ring)
constexpr int SEL_STACK = 48;           // per-warp private stack
#ifndef YAVO_SEL_LOCAL_SINGLE
#define YAVO_SEL_LOCAL_SINGLE 16  // the single-frame instance hands almost every right child to the queue (idle warps): 62.6 -> 58 us
#endif
#ifndef YAVO_SEL_LOCAL
#define YAVO_SEL_LOCAL 32  // measured at the end of round 2: 0 0.309 ms, 16 0.291, 24 0.287, 32 0.286, 64 0.298, 128 0.345 per 1024 frames
#endif
constexpr int SEL_LOCAL = YAVO_SEL_LOCAL;           // right children up to this size stay with the warp that produced them
constexpr int SEL_BIG = 64;             // per-level list of CTA-partitioned ranges

struct SelRange {
    int f, l, d;
};

template <int NT>
struct SelSharedT {
    static constexpr int SEL_WARPS = NT / 32;
    SelRange ring[SEL_QCAP];
    int ready[SEL_QCAP];
    SelRange stack[SEL_WARPS][SEL_STACK];
    uint16_t wscratch[SEL_WARPS][2][SEL_WARP_MAX / 2 + 2];
    uint16_t bscratch[2][SEL_SMEM_ENTS / 2 + 2];  // CTA-partition stopper lists while the data is in smem
    SelRange big[2][SEL_BIG];
    int nbig[2];
    int q_head, q_tail, pending, watchdog;
    int wtot[SEL_WARPS];
    int bcast[4];
};

// queue a range for phase 2 (one thread).  Returns false when the ring is full.
template <int NT>
__device__ __forceinline__ bool sel_push(SelSharedT<NT> &S, const SelRange &r) {
    constexpr int SEL_WARPS = NT / 32;
    const int head = *(volatile int *)&S.q_head, tail = *(volatile int *)&S.q_tail;
    if (tail - head >= SEL_QCAP - 2 * SEL_WARPS) return false;
    const int idx = atomicAdd(&S.q_tail, 1);
    S.ring[idx % SEL_QCAP] = r;
    __threadfence_block();
    atomicExch(&S.ready[idx % SEL_QCAP], idx + 1);
    return true;
}

// warp-collective pop; returns false once no work is left anywhere in the CTA
template <int NT>
__device__ __forceinline__ bool sel_pop(SelSharedT<NT> &S, SelRange &out) {
    const int lane = threadIdx.x & 31;
    int got = -1;
    if (lane == 0) {
        int spins = 0;
        for (;;) {
            const int h = *(volatile int *)&S.q_head, t = *(volatile int *)&S.q_tail;
            if (h < t) {
                if (atomicCAS(&S.q_head, h, h + 1) == h) {
                    got = h;
                    break;
                }
                continue;
            }
            if (*(volatile int *)&S.pending <= 0) break;
            __nanosleep(YAVO_SEL_SLEEP);
            if (++spins > (1 << 22)) {  // watchdog (~0.5 s): never hang the GPU on a scheduling bug; report instead
                atomicExch(&S.watchdog, 1);
                atomicExch(&S.pending, 0);
                break;
            }
        }
        if (got >= 0) {
            while (*(volatile int *)&S.ready[got % SEL_QCAP] != got + 1) {
            }
            __threadfence_block();
        }
    }
    got = __shfl_sync(0xffffffffu, got, 0);
    if (got < 0) return false;
    out = S.ring[got % SEL_QCAP];
    return true;
}

// whole-CTA partition of [f,l); returns the cut to every thread.  Lpos/Rpos: global scratch.
// Each thread classifies SEL_ITEMS consecutive positions per pass (one block-wide scan per 4096 elements).
#ifndef YAVO_SEL_ITEMS
#define YAVO_SEL_ITEMS 5  // measured with the ballot form: 4 0.299 ms, 6 0.300, 8 0.315, 12 0.339 per 1024 frames; at the end of the round 4 0.2864, 5 0.2849
#endif
constexpr int SEL_ITEMS = YAVO_SEL_ITEMS;

template <int NT, typename PosT>
__device__ int sel_block_partition(SelSharedT<NT> &S, yavo_ent *A, int f, int l, PosT *Lpos, PosT *Rpos) {
    constexpr int SEL_THREADS = NT, SEL_WARPS = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = l - f;
    // __move_median_to_first(first, first+1, mid, last-1): every thread works the pivot out for itself (four broadcast
    // loads) and classifies against the list as it will look after the move — the entry at `mpos` read as the old
    // first element — so no barrier and no single-thread section precede the classification; thread 0 performs the
    // move itself after the first barrier below (every thread has read the four entries by then; the pair swaps, which
    // may touch `mpos`, come after the loop's last barrier)
    const int pa = f + 1, pb = f + n / 2, pc = l - 1;
    const yavo_ent v0 = A[f], va = A[pa], vb = A[pb], vc = A[pc];
    int mpos;
    if (yavo_before(va, vb)) mpos = yavo_before(vb, vc) ? pb : (yavo_before(va, vc) ? pc : pa);
    else mpos = yavo_before(va, vc) ? pa : (yavo_before(vb, vc) ? pc : pb);
    const yavo_ent piv = mpos == pa ? va : (mpos == pb ? vb : vc);
    const int cap = n / 2 + 1;
    int runL = 0, runR = 0;
    // A warp classifies a contiguous chunk of 32 * SEL_ITEMS scan positions per pass, lane l taking positions l, l + 32, ...
    // of it: consecutive lanes read consecutive entries (eight consecutive entries per THREAD, as before, are 64 bytes
    // apart across the lanes of a load: a 16-way bank conflict on every shared-memory read).  Ranks inside the chunk come
    // from ballots (one per side and item), the chunk totals from their population counts, the order across warps from
    // the per-warp totals as before.
    const unsigned lt = (1u << lane) - 1u;
    for (int base = 0; base < n - 1; base += SEL_THREADS * SEL_ITEMS) {
        const int w0 = base + warp * (32 * SEL_ITEMS) + lane;
        unsigned bL[SEL_ITEMS], bR[SEL_ITEMS];
        int packed = 0;
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; e++) {
            const int i = w0 + 32 * e;
            bool sL = false, sR = false;
            if (i < n - 1) {
                const int pl = f + 1 + i, pr = l - 1 - i;
                const yavo_ent el = A[pl], er = A[pr];
                sL = !yavo_before(pl == mpos ? v0 : el, piv);
                sR = !yavo_before(piv, pr == mpos ? v0 : er);
            }
            bL[e] = __ballot_sync(0xffffffffu, sL);
            bR[e] = __ballot_sync(0xffffffffu, sR);
            packed += __popc(bL[e]) | (__popc(bR[e]) << 16);
        }
        if (lane == 0) S.wtot[warp] = packed;  // the warp's totals (every lane holds them)
        __syncthreads();
        if (base == 0 && tid == 0) {  // the move of the median to the front
            A[f] = piv;
            A[mpos] = v0;
        }
        int pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < SEL_WARPS; w++) {
            const int t = S.wtot[w];
            if (w < warp) pre += t;
            tot += t;
        }
        int rL = runL + (pre & 0xffff), rR = runR + (pre >> 16);  // ranks of the warp's first left / right stopper
#pragma unroll
        for (int e = 0; e < SEL_ITEMS; e++) {
            const int i = w0 + 32 * e;
            if ((bL[e] >> lane) & 1u) {
                const int r = rL + __popc(bL[e] & lt);
                if (r < cap) Lpos[r] = (PosT)(1 + i);  // positions relative to f
            }
            if ((bR[e] >> lane) & 1u) {
                const int r = rR + __popc(bR[e] & lt);
                if (r < cap) Rpos[r] = (PosT)(n - 1 - i);
            }
            rL += __popc(bL[e]);
            rR += __popc(bR[e]);
        }
        runL += tot & 0xffff;
        runR += tot >> 16;
        __syncthreads();
    }
    const int nL = min(runL, cap), nR = min(runR, cap);
    const int npairs = min(nL, nR);
    int cnt = 0;
    for (int i = tid; i < npairs; i += SEL_THREADS * 4) {
        // four independent swaps per thread in flight (the pairs are disjoint)
        uint32_t a[4], b[4];
        yavo_ent va[4], vb[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = i + u * SEL_THREADS;
            ok[u] = false;
            if (k < npairs) {
                a[u] = (uint32_t)Lpos[k];
                b[u] = (uint32_t)Rpos[k];
                ok[u] = a[u] < b[u];
                a[u] += f;
                b[u] += f;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (ok[u]) {
                va[u] = A[a[u]];
                vb[u] = A[b[u]];
            }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (ok[u]) {
                A[a[u]] = vb[u];
                A[b[u]] = va[u];
                cnt++;
            }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) S.wtot[warp] = cnt;
    __syncthreads();
    int m = 0;  // every thread sums the swap counts and reads the two stopper positions that bound the cut
#pragma unroll
    for (int w = 0; w < SEL_WARPS; w++) m += S.wtot[w];
    uint32_t cut = 0xffffffffu;
    if (m < nL) cut = (uint32_t)Lpos[m];
    if (m >= 1) cut = min(cut, (uint32_t)Rpos[m - 1]);
    __syncthreads();  // the next partition reuses the counts and the stopper lists
    return f + (int)cut;
}

// single-warp partition of [f,l), 16 < n <= SEL_WARP_MAX; positions relative to f in 16-bit scratch.
// Written for latency (a frame's phase 2 is ~400 of these, each a chain of dependent shared-memory round trips on a
// warp that has nothing else to do): every lane reads the four median candidates itself (broadcast loads, no
// shuffles) and classifies against the list as it will look after __move_median_to_first — the entry at `mpos` read
// as the old first element — so the move itself (lane 0) happens off the critical path; two 32-element chunks are
// loaded before either is classified.
template <int NT>
__device__ int sel_warp_partition(SelSharedT<NT> &S, yavo_ent *A, int f, int l) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint16_t *Lpos = S.wscratch[warp][0], *Rpos = S.wscratch[warp][1];
    const int n = l - f;
    // __move_median_to_first(first, first+1, mid, last-1), libstdc++'s comparison sequence
    const int pa = f + 1, pb = f + n / 2, pc = l - 1;
    const yavo_ent v0 = A[f], va = A[pa], vb = A[pb], vc = A[pc];
    int mpos;
    if (yavo_before(va, vb)) mpos = yavo_before(vb, vc) ? pb : (yavo_before(va, vc) ? pc : pa);
    else mpos = yavo_before(va, vc) ? pa : (yavo_before(vb, vc) ? pc : pb);
    const yavo_ent piv = mpos == pa ? va : (mpos == pb ? vb : vc);
    const int cap = n / 2 + 1;
    const unsigned lt = (1u << lane) - 1u;
    int runL = 0, runR = 0;
    for (int base = 0; base < n - 1; base += 64) {
        const int i0 = base + lane, i1 = i0 + 32;
        const bool in0 = i0 < n - 1, in1 = i1 < n - 1;
        yavo_ent l0 = 0ull, r0 = 0ull, l1 = 0ull, r1 = 0ull;
        if (in0) {
            l0 = A[f + 1 + i0];
            r0 = A[l - 1 - i0];
        }
        if (in1) {
            l1 = A[f + 1 + i1];
            r1 = A[l - 1 - i1];
        }
        if (f + 1 + i0 == mpos) l0 = v0;
        if (l - 1 - i0 == mpos) r0 = v0;
        if (f + 1 + i1 == mpos) l1 = v0;
        if (l - 1 - i1 == mpos) r1 = v0;
        const bool sL0 = in0 && !yavo_before(l0, piv), sR0 = in0 && !yavo_before(piv, r0);
        const bool sL1 = in1 && !yavo_before(l1, piv), sR1 = in1 && !yavo_before(piv, r1);
        const unsigned bL0 = __ballot_sync(0xffffffffu, sL0), bR0 = __ballot_sync(0xffffffffu, sR0);
        const unsigned bL1 = __ballot_sync(0xffffffffu, sL1), bR1 = __ballot_sync(0xffffffffu, sR1);
        const int cL0 = __popc(bL0), cR0 = __popc(bR0);
        const int rL0 = runL + __popc(bL0 & lt), rR0 = runR + __popc(bR0 & lt);
        const int rL1 = runL + cL0 + __popc(bL1 & lt), rR1 = runR + cR0 + __popc(bR1 & lt);
        if (sL0 && rL0 < cap) Lpos[rL0] = (uint16_t)(1 + i0);
        if (sR0 && rR0 < cap) Rpos[rR0] = (uint16_t)(n - 1 - i0);
        if (sL1 && rL1 < cap) Lpos[rL1] = (uint16_t)(1 + i1);
        if (sR1 && rR1 < cap) Rpos[rR1] = (uint16_t)(n - 1 - i1);
        runL += cL0 + __popc(bL1);
        runR += cR0 + __popc(bR1);
    }
    if (lane == 0) {  // the move of the median to the front (every lane holds the four entries it read above)
        A[f] = piv;
        A[mpos] = v0;
    }
    __syncwarp();
    const int nL = min(runL, cap), nR = min(runR, cap);
    const int npairs = min(nL, nR);
    int m = 0;  // number of swaps: the pairs with L_i < R_i form a prefix
    for (int i0 = 0; i0 < npairs; i0 += 32) {
        const int i = i0 + lane;
        bool ok = false;
        int a = 0, b = 0;
        if (i < npairs) {
            a = Lpos[i];
            b = Rpos[i];
            ok = a < b;
        }
        if (ok) {
            const yavo_ent t = A[f + a];
            A[f + a] = A[f + b];
            A[f + b] = t;
        }
        const unsigned bo = __ballot_sync(0xffffffffu, ok);
        m += __popc(bo);
        if (bo != 0xffffffffu) break;
    }
    int cut = 0x7fffffff;
    if (m < nL) cut = Lpos[m];
    if (m >= 1) cut = min(cut, (int)Rpos[m - 1]);
    __syncwarp();
    return f + cut;
}

// stable sort of a range of <= 16 elements by one warp (what __final_insertion_sort does to it):
// rank = elements that sort strictly before + equal elements that come earlier.  (Sixteen unrolled shuffles instead of
// the loop over the range's length: measured slower, 0.398 vs 0.366 ms per 1024 frames — the leaves are short; eight
// shuffles per half-warp + one to combine: 0.348 vs 0.334.  The kernel is sensitive to the number of shuffle / shared-
// memory instructions, not only to their latency.)
__device__ __forceinline__ void sel_warp_leaf(yavo_ent *A, int f, int l) {
    const int lane = threadIdx.x & 31, n = l - f;
    const yavo_ent e = (lane < n) ? A[f + lane] : 0ull;
    const float se = yavo_ent_score(e);
    int rank = 0;
    for (int j = 0; j < n; j++) {
        const float sj = __shfl_sync(0xffffffffu, se, j);
        rank += (sj > se) || (sj == se && j < lane);
    }
    __syncwarp();
    if (lane < n) A[f + rank] = e;
    __syncwarp();
}

// phase 2: one warp works a range down to its leaves, handing right children to the queue
#ifdef YAVO_SEL_TIMING
__device__ long long g_sel_dbg[8];  // [0] partition cycles [1] partitions [2] leaf cycles [3] leaves [4] elements partitioned
#define SEL_DBG_ADD(i, v) do { if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) atomicAdd((unsigned long long *)&g_sel_dbg[i], (unsigned long long)(v)); } while (0)
#else
#define SEL_DBG_ADD(i, v) do { } while (0)
#endif

template <int NT>
__device__ void sel_warp_work(SelSharedT<NT> &S, yavo_ent *A, SelRange cur, int K) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // Right children of at most SEL_LOCAL elements stay on this warp's private stack: no queue traffic, no
    // atomics — most partitions are of small ranges.  Larger ones go to the shared queue so idle warps can take
    // them.  S.pending counts queue-level tasks only: this warp's task ends when its private stack is empty.
    int sp = 0;
    for (;;) {
        // cur satisfies cur.f < K, cur.l - cur.f > 1
        const int n = cur.l - cur.f;
        bool finished = false;
        if (n <= YAVO_SORT_THRESHOLD) {
#ifdef YAVO_SEL_TIMING
            const long long c0 = clock64();
#endif
            sel_warp_leaf(A, cur.f, cur.l);
#ifdef YAVO_SEL_TIMING
            SEL_DBG_ADD(2, clock64() - c0);
            SEL_DBG_ADD(3, 1);
#endif
            finished = true;
        } else if (cur.d == 0) {
            if (lane == 0) yavo_serial_heapsort(A, cur.f, cur.l);  // libstdc++'s depth-limit fallback
            __syncwarp();
            finished = true;
        } else if (n > SEL_WARP_MAX) {  // cannot happen: phase 1 leaves only ranges <= SEL_WARP_MAX; stay exact anyway
            if (lane == 0) yavo_serial_introsort(A, cur.f, cur.l, cur.d, K);
            __syncwarp();
            finished = true;
        } else {
#ifdef YAVO_SEL_TIMING
            const long long c0 = clock64();
#endif
            const int cut = sel_warp_partition(S, A, cur.f, cur.l);
#ifdef YAVO_SEL_TIMING
            SEL_DBG_ADD(0, clock64() - c0);
            SEL_DBG_ADD(1, 1);
            SEL_DBG_ADD(4, n);
#endif
            const SelRange left = {cur.f, cut, cur.d - 1}, right = {cut, cur.l, cur.d - 1};
            const bool vL = cut - cur.f > 1;               // left starts at cur.f < K
            const bool vR = cut < K && cur.l - cut > 1;
            if (vL && vR) {
                bool stacked = false;
                if (right.l - right.f <= (NT == SEL_THREADS_BATCH ? SEL_LOCAL : YAVO_SEL_LOCAL_SINGLE) && sp < SEL_STACK) {
                    stacked = true;
                } else {
                    int pushed = 0;
                    if (lane == 0) {
                        atomicAdd(&S.pending, 1);
                        pushed = sel_push(S, right) ? 1 : 0;
                        if (!pushed) atomicSub(&S.pending, 1);
                    }
                    pushed = __shfl_sync(0xffffffffu, pushed, 0);
                    if (!pushed) {
                        if (sp < SEL_STACK) {
                            stacked = true;
                        } else {  // queue and stack full: finish the child serially (exact, practically unreachable)
                            if (lane == 0) yavo_serial_introsort(A, right.f, right.l, right.d, K);
                            __syncwarp();
                        }
                    }
                }
                if (stacked) {
                    if (lane == 0) S.stack[warp][sp] = right;
                    sp++;
                    __syncwarp();
                }
                cur = left;
            } else if (vL) {
                cur = left;
            } else if (vR) {
                cur = right;
            } else {
                finished = true;
            }
        }
        if (finished) {
            if (sp == 0) {
                if (lane == 0) {
                    __threadfence_block();
                    atomicSub(&S.pending, 1);
                }
                return;
            }
            sp--;
            cur = S.stack[warp][sp];
        }
    }
}

// ================================================================================================
// K3a  the top of the partition tree of a LARGE candidate list, on a thread-block cluster
// One CTA per frame is the right shape for KITTI-size frames (5 k candidates, hundreds of frames in flight); a 4K frame
// has 300 k candidates and a batch only a handful of frames, so its first partitions — which scan hundreds of
// thousands of elements in global memory — are spread over the BIG_CL CTAs of a cluster (one cluster per frame):
//   gather  every CTA takes a slice of the segment table; slice totals are exchanged, then each copies its slice
//   per partition of a range [f, l):  (the same parallel form of __unguarded_partition as sel_block_partition)
//     rank 0: median-of-three to the front                                                     | cluster barrier
//     pass 1: every CTA counts the left / right stoppers of its slice of the scan index range  | cluster barrier
//     pass 2: ... and scatters their positions at [slice base + rank] into the stopper lists   | cluster barrier
//     swaps: pair k (k-th left stopper, k-th right stopper) while L_k < R_k, spread over all threads; swap count
//                                                                                              | cluster barrier
//     cut = min(L_m, R_{m-1}); children above BIG_MIN elements go to the next level, the others (that start below K)
//     are handed to the select kernel, which continues from this partially partitioned state.
// Counts travel through a small global exchange block: the cluster barrier (release / acquire at cluster scope)
// orders them, and the 8 CTAs of a cluster are co-scheduled by construction, so the barrier cannot deadlock.
// ================================================================================================
#ifndef YAVO_BIG_CL
#define YAVO_BIG_CL 8
#endif
constexpr int BIG_CL = YAVO_BIG_CL;  // CTAs per cluster (8 = the portable maximum; 16 needs the non-portable opt-in)
constexpr int BIG_THREADS = 512;
#ifndef YAVO_BIG_ITEMS
#define YAVO_BIG_ITEMS 8
#endif
constexpr int BIG_ITEMS = YAVO_BIG_ITEMS;
constexpr int BIG_PRE = 64;        // ranges handed to the select kernel per frame
constexpr int BIG_XCHG = 4 * BIG_CL;  // ints per frame in the exchange block: cntL | cntR | swaps | gather totals

struct BigShared {
    int wtot[BIG_THREADS / 32];
    SelRange lvl[2][SEL_BIG];
    int nlvl[2];
    int bc[2];
};

// block-wide exclusive scan of a packed (low 16 bits | high 16 bits) per-thread count; returns this thread's exclusive
// prefix and the block total (both packed)
__device__ __forceinline__ void big_block_scan(BigShared &S, int packed, int *excl, int *total) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int incl = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // the previous use of wtot is over
    if (lane == 31) S.wtot[warp] = incl;
    __syncthreads();
    int pre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < BIG_THREADS / 32; w++) {
        const int t = S.wtot[w];
        if (w < warp) pre += t;
        tot += t;
    }
    *excl = pre + incl - packed;
    *total = tot;
}

__device__ __forceinline__ int big_block_sum(BigShared &S, int v) {
    int e, t;
    big_block_scan(S, v, &e, &t);
    return t;
}

__global__ void __cluster_dims__(BIG_CL, 1, 1) __launch_bounds__(BIG_THREADS)
select_big_kernel(const uint32_t *__restrict__ seg, int seg_cols, int rows_alloc, int ntx, const yavo_ent *__restrict__ pool,
                  yavo_ent *__restrict__ cand_all, int max_cand, const int *__restrict__ ncand,
                  uint32_t *__restrict__ scratch_all, int *__restrict__ xchg_all, int K, int H, int big_min,
                  SelRange *__restrict__ pre_all, int *__restrict__ n_pre) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ BigShared S;
    // cluster barrier + a device-scope fence on the way out: what other CTAs wrote to global memory before the barrier
    // (entries, stopper lists, counts) must not be served from this SM's L1
    auto csync = [&]() {
        cluster.sync();
        __threadfence();
    };
    const int rank = (int)cluster.block_rank(), f = blockIdx.x / BIG_CL;
    const int tid = threadIdx.x;
    const int N = ncand[f];
    if (N <= big_min || N > max_cand) return;  // the select kernel does everything (uniform over the cluster)
    yavo_ent *A = cand_all + (size_t)f * max_cand;
    uint32_t *Lpos = scratch_all + (size_t)f * (size_t)(2 * max_cand + 8);
    uint32_t *Rpos = Lpos + (max_cand / 2 + 2);
    volatile int *X = xchg_all + (size_t)f * BIG_XCHG;
    const uint32_t *seg_f = seg + (size_t)f * rows_alloc * seg_cols;
    const yavo_ent *pool_f = pool + (size_t)f * max_cand;

    // ---- gather: this CTA's slice of the segment table, one contiguous run of segments per thread -----------------
    {
        const int Sg = H * ntx, per_cta = (Sg + BIG_CL - 1) / BIG_CL;
        const int c0 = min(Sg, rank * per_cta), c1 = min(Sg, c0 + per_cta);
        const int per = (c1 - c0 + BIG_THREADS - 1) / BIG_THREADS;
        const int e0 = min(c1, c0 + tid * per), e1 = min(c1, e0 + per);
        int mine = 0;
        {
            int row = e0 / ntx, tx = e0 - row * ntx;
            for (int e = e0; e < e1; e++) {
                mine += (int)(seg_f[(size_t)row * seg_cols + tx] & ((1u << SEG_CNT_BITS) - 1u));
                if (++tx == ntx) { tx = 0; row++; }
            }
        }
        int excl, total;
        // counts can exceed 16 bits here: scan the plain value (the packed form is for the stopper counts)
        big_block_scan(S, mine, &excl, &total);
        if (tid == 0) X[3 * BIG_CL + rank] = total;
        __threadfence();
        csync();
        int base = 0;
        for (int c = 0; c < rank; c++) base += X[3 * BIG_CL + c];
        int o = base + excl;
        int row = e0 / ntx, tx = e0 - row * ntx;
        for (int e = e0; e < e1; e++) {
            const uint32_t sg = seg_f[(size_t)row * seg_cols + tx];
            const int cnt = (int)(sg & ((1u << SEG_CNT_BITS) - 1u));
            const yavo_ent *src = pool_f + (sg >> SEG_CNT_BITS);
            for (int k = 0; k < cnt; k++) A[o + k] = src[k];
            o += cnt;
            if (++tx == ntx) { tx = 0; row++; }
        }
        __threadfence();
        csync();
    }

    if (tid == 0) {
        S.lvl[0][0] = {0, N, 2 * (31 - __clz(N))};
        S.nlvl[0] = 1;
        S.nlvl[1] = 0;
    }
    __syncthreads();
    int npre = 0;  // handed-over ranges so far (every CTA counts; rank 0 writes them)
    SelRange *pre = pre_all + (size_t)f * BIG_PRE;
    int cur = 0;
    while (S.nlvl[cur] > 0) {
        const int nb = S.nlvl[cur], nxt = cur ^ 1;
        for (int ri = 0; ri < nb; ri++) {
            const SelRange r = S.lvl[cur][ri];
            const int rf = r.f, rl = r.l, n = rl - rf;
            if (r.d == 0) {  // depth limit (never reached on score lists): the select kernel's heapsort takes the range
                if (rank == 0 && tid == 0 && npre < BIG_PRE) pre[npre] = r;
                npre++;
                continue;
            }
            if (rank == 0 && tid == 0) {
                yavo_median_to_first(A, rf, rl);
                __threadfence();
            }
            csync();
            const yavo_ent piv = A[rf];
            const int cap = n / 2 + 1;
            const int per_cta = (n - 1 + BIG_CL - 1) / BIG_CL;
            const int i_lo = min(n - 1, rank * per_cta), i_hi = min(n - 1, i_lo + per_cta);
            // pass 1: stoppers in this CTA's slice of the scan index range
            {
                int cL = 0, cR = 0;
                for (int i = i_lo + tid; i < i_hi; i += BIG_THREADS) {
                    cL += !yavo_before(A[rf + 1 + i], piv);
                    cR += !yavo_before(piv, A[rl - 1 - i]);
                }
                const int tL = big_block_sum(S, cL), tR = big_block_sum(S, cR);
                if (tid == 0) {
                    X[rank] = tL;
                    X[BIG_CL + rank] = tR;
                }
                __threadfence();
            }
            csync();
            int baseL = 0, baseR = 0, nL = 0, nR = 0;
            for (int c = 0; c < BIG_CL; c++) {
                const int a = X[c], b = X[BIG_CL + c];
                if (c < rank) { baseL += a; baseR += b; }
                nL += a;
                nR += b;
            }
            // pass 2: positions of the stoppers at [base + rank inside the slice]
            {
                int runL = baseL, runR = baseR;
                // warp-contiguous chunks, ranks from ballots (as sel_block_partition): every load of a warp is one
                // contiguous 256-byte piece of the list
                const int lane = tid & 31, warp = tid >> 5;
                const unsigned lt = (1u << lane) - 1u;
                for (int base = i_lo; base < i_hi; base += BIG_THREADS * BIG_ITEMS) {
                    const int w0 = base + warp * (32 * BIG_ITEMS) + lane;
                    unsigned bL[BIG_ITEMS], bR[BIG_ITEMS];
                    int packed = 0;
#pragma unroll
                    for (int e = 0; e < BIG_ITEMS; e++) {
                        const int i = w0 + 32 * e;
                        bool sL = false, sR = false;
                        if (i < i_hi) {
                            sL = !yavo_before(A[rf + 1 + i], piv);
                            sR = !yavo_before(piv, A[rl - 1 - i]);
                        }
                        bL[e] = __ballot_sync(0xffffffffu, sL);
                        bR[e] = __ballot_sync(0xffffffffu, sR);
                        packed += __popc(bL[e]) | (__popc(bR[e]) << 16);
                    }
                    __syncthreads();  // the previous use of wtot is over
                    if (lane == 0) S.wtot[warp] = packed;
                    __syncthreads();
                    int pre = 0, tot = 0;
#pragma unroll
                    for (int w = 0; w < BIG_THREADS / 32; w++) {
                        const int t = S.wtot[w];
                        if (w < warp) pre += t;
                        tot += t;
                    }
                    int rL = runL + (pre & 0xffff), rR = runR + (pre >> 16);
#pragma unroll
                    for (int e = 0; e < BIG_ITEMS; e++) {
                        const int i = w0 + 32 * e;
                        if ((bL[e] >> lane) & 1u) {
                            const int r = rL + __popc(bL[e] & lt);
                            if (r < cap) Lpos[r] = (uint32_t)(1 + i);  // positions relative to rf
                        }
                        if ((bR[e] >> lane) & 1u) {
                            const int r = rR + __popc(bR[e] & lt);
                            if (r < cap) Rpos[r] = (uint32_t)(n - 1 - i);
                        }
                        rL += __popc(bL[e]);
                        rR += __popc(bR[e]);
                    }
                    runL += tot & 0xffff;
                    runR += tot >> 16;
                }
                __threadfence();
            }
            csync();
            // swaps: the pairs with L_k < R_k form a prefix of the pair list
            const int nLc = min(nL, cap), nRc = min(nR, cap), npairs = min(nLc, nRc);
            {
                int cnt = 0;
                for (int k = rank * BIG_THREADS + tid; k < npairs; k += BIG_CL * BIG_THREADS) {
                    const uint32_t a = Lpos[k], b = Rpos[k];
                    if (a < b) {
                        const yavo_ent va = A[rf + a], vb = A[rf + b];
                        A[rf + a] = vb;
                        A[rf + b] = va;
                        cnt++;
                    }
                }
                const int t = big_block_sum(S, cnt);
                if (tid == 0) X[2 * BIG_CL + rank] = t;
                __threadfence();
            }
            csync();
            if (tid == 0) {
                int m = 0;
                for (int c = 0; c < BIG_CL; c++) m += X[2 * BIG_CL + c];
                uint32_t cut = 0xffffffffu;
                if (m < nLc) cut = Lpos[m];
                if (m >= 1) cut = min(cut, Rpos[m - 1]);
                S.bc[0] = rf + (int)cut;
            }
            __syncthreads();
            const int cut = S.bc[0];
            if (tid == 0) {
                const SelRange ch[2] = {{rf, cut, r.d - 1}, {cut, rl, r.d - 1}};
                for (int c = 0; c < 2; c++) {
                    if (ch[c].f >= K || ch[c].l - ch[c].f <= 1) continue;
                    if (ch[c].l - ch[c].f > big_min && S.nlvl[nxt] < SEL_BIG) {
                        S.lvl[nxt][S.nlvl[nxt]++] = ch[c];
                    } else {
                        if (rank == 0 && npre < BIG_PRE) pre[npre] = ch[c];
                        npre++;
                    }
                }
                S.bc[1] = npre;
            }
            __syncthreads();
            npre = S.bc[1];
        }
        __syncthreads();
        if (tid == 0) S.nlvl[cur] = 0;
        cur = nxt;
        __syncthreads();
    }
    if (rank == 0 && tid == 0) {
        __threadfence();
        n_pre[f] = min(npre, BIG_PRE);
    }
}

// checkBoundry of reference src/BriefDescriptor.cc:128-136 as computeBrief calls it (:97)
__device__ __forceinline__ bool brief_admits(int row, int col, int H, int W) {
    return !(col - 8 < 0 || col + 8 > W || row - 8 < 0 || row + 8 > H);
}

// Two instances: NT = SEL_THREADS_BATCH for batches (hundreds of frames in flight, two CTAs per SM) and
// NT = SEL_THREADS_SINGLE for the single-frame path (yavo_frame_features / yavo_fast_detect on one slot: the GPU is
// otherwise idle, so the one CTA takes twice the warps for the work queue of phase 2).
template <int NT>
__global__ void __launch_bounds__(NT, NT == SEL_THREADS_BATCH ? YAVO_SEL_MIN_CTAS : 1)
select_topk_kernel(const uint32_t *__restrict__ seg, int seg_cols, int rows_alloc, int ntx,
                   const yavo_ent *__restrict__ pool, yavo_ent *__restrict__ cand_all, int max_cand,
                   const int *__restrict__ ncand, uint32_t *__restrict__ scratch_all, int K, int H, int W, int kp_stride,
                   int32_t *__restrict__ kp_row, int32_t *__restrict__ kp_col, float *__restrict__ kp_score,
                   int *__restrict__ nkp,
                   // compacted (checkBoundry-admitted) list that BRIEF / the matcher consume
                   int32_t *__restrict__ bk_row, int32_t *__restrict__ bk_col, float *__restrict__ bk_score,
                   int32_t *__restrict__ bk_id, int *__restrict__ nbk, int *__restrict__ status,
                   const SelRange *__restrict__ pre_all = nullptr, const int *__restrict__ n_pre = nullptr,
                   int team = 1 /* CTAs per frame (> 1 only together with the cluster pre-partition) */,
                   int *__restrict__ team_done = nullptr) {
    extern __shared__ __align__(16) unsigned char sel_smem_raw[];
    constexpr int SEL_THREADS = NT, SEL_WARPS = NT / 32;
    using SelShared = SelSharedT<NT>;
    SelShared &S = *reinterpret_cast<SelShared *>(sel_smem_raw);
    yavo_ent *sbuf = reinterpret_cast<yavo_ent *>(sel_smem_raw + ((sizeof(SelShared) + 15) & ~size_t(15)));

    // team > 1: the frame's handed-over ranges (disjoint spans of the list) are dealt round-robin to `team` independent
    // CTAs, each of which replays its ranges to the leaves in global memory; the CTA that finishes last writes the outputs
    const int f = blockIdx.x / team, member = blockIdx.x - f * team;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef YAVO_SEL_TIMING
    long long tmark[6];
    tmark[0] = clock64();
#define SEL_MARK(i) do { __syncthreads(); tmark[i] = clock64(); } while (0)
#else
#define SEL_MARK(i) do { } while (0)
#endif
    const int N = ncand[f];
    if (N > max_cand) {  // candidate buffer overflow: report, never return a silently truncated order
        if (tid == 0) { atomicExch(status, 1); nkp[f] = 0; nbk[f] = 0; }
        return;
    }
    yavo_ent *G = cand_all + (size_t)f * max_cand;
    // stopper lists of CTA-wide partitions in global memory: per frame 2 * max_cand + 8 entries; a range [rf, rl) uses
    // [2 rf, 2 rf + n + 4), so concurrent partitions of disjoint ranges (team members) never share entries
    uint32_t *Lbase = scratch_all + (size_t)f * (size_t)(2 * max_cand + 8);
    const int np = n_pre ? n_pre[f] : 0;
    if (member > 0 && np == 0) return;  // a short list: member 0 does everything
    const bool solo = team == 1 || np == 0;

    for (int i = tid; i < SEL_QCAP; i += SEL_THREADS) S.ready[i] = 0;
    if (tid == 0) { S.nbig[0] = S.nbig[1] = 0; S.q_head = S.q_tail = 0; S.pending = 0; S.watchdog = 0; }
    __syncthreads();

    // load the candidate list in scan order: the detect kernel scored the corners and left them in the frame's pool;
    // the segment table puts them back in the order the reference appends retCorners (src/FastDetector.cc:298-324)
    // a large list arrives gathered and partitioned at the top by the cluster kernel (K3a): continue from its ranges
    const bool in_smem_at_start = np == 0 && N <= SEL_SMEM_ENTS;
    yavo_ent *A = in_smem_at_start ? sbuf : G;
    bool in_smem = in_smem_at_start;
    if (np == 0)
        gather_candidates<SEL_THREADS>(seg + (size_t)f * rows_alloc * seg_cols, seg_cols, H, ntx, pool + (size_t)f * max_cand, A,
                                       in_smem ? SEL_SMEM_ENTS : max_cand, S.wtot);
    if (tid == 0 && np > 0) {
        const SelRange *pre = pre_all + (size_t)f * BIG_PRE;
        for (int i = member; i < np; i += team) {
            const SelRange r = pre[i];
            if (r.l - r.f > SEL_WARP_MAX && S.nbig[0] < SEL_BIG) S.big[0][S.nbig[0]++] = r;
            else if (r.l - r.f > SEL_WARP_MAX || !sel_push(S, r)) yavo_serial_introsort(A, r.f, r.l, r.d, K);  // lists full: exact, serial
        }
    } else if (tid == 0 && N > 1) {
        const SelRange r0 = {0, N, 2 * (31 - __clz(N))};
        if (N > SEL_WARP_MAX) S.big[0][S.nbig[0]++] = r0;
        else sel_push(S, r0);
    }
    __syncthreads();

    SEL_MARK(1);
    // ---- phase 1: ranges larger than SEL_WARP_MAX, whole CTA, level by level -------------------------
    int cur = 0;
    while (S.nbig[cur] > 0) {
        const int nb = S.nbig[cur], nxt = cur ^ 1;
        if (!in_smem && solo) {  // move the active prefix into shared memory as soon as it fits
            int E = min(N, K);
            for (int i = 0; i < nb; i++) E = max(E, S.big[cur][i].l);
            const int qt = S.q_tail;
            for (int i = 0; i < qt; i++) E = max(E, S.ring[i].l);
            if (E <= SEL_SMEM_ENTS) {
                for (int i = tid; i < E; i += SEL_THREADS) sbuf[i] = G[i];
                A = sbuf;
                in_smem = true;
                __syncthreads();
            }
        }
        for (int i = 0; i < nb; i++) {
            const SelRange r = S.big[cur][i];
            int cut;
            if (r.d == 0) {  // depth limit: heapsort, nothing below it
                if (tid == 0) yavo_serial_heapsort(A, r.f, r.l);
                __syncthreads();
                continue;
            }
            cut = in_smem ? sel_block_partition<NT, uint16_t>(S, A, r.f, r.l, S.bscratch[0], S.bscratch[1])
                          : sel_block_partition<NT, uint32_t>(S, A, r.f, r.l, Lbase + 2 * (size_t)r.f,
                                                          Lbase + 2 * (size_t)r.f + ((r.l - r.f) / 2 + 2));
            if (tid == 0) {
                const SelRange ch[2] = {{r.f, cut, r.d - 1}, {cut, r.l, r.d - 1}};
                for (int c = 0; c < 2; c++) {
                    if (ch[c].f >= K || ch[c].l - ch[c].f <= 1) continue;
                    if (ch[c].l - ch[c].f > SEL_WARP_MAX && S.nbig[nxt] < SEL_BIG) S.big[nxt][S.nbig[nxt]++] = ch[c];
                    else if (ch[c].l - ch[c].f > SEL_WARP_MAX || !sel_push(S, ch[c]))
                        yavo_serial_introsort(A, ch[c].f, ch[c].l, ch[c].d, K);  // lists full: exact, serial
                }
            }
        }
        __syncthreads();
        if (tid == 0) S.nbig[cur] = 0;
        cur = nxt;
        __syncthreads();
    }
    SEL_MARK(2);
    // move the active prefix into shared memory if it fits (phase 2 then never touches global memory)
    if (!in_smem && solo) {
        int E = min(N, K);
        const int qt = S.q_tail;
        for (int i = 0; i < qt && i < SEL_QCAP; i++) E = max(E, S.ring[i].l);
        if (E <= SEL_SMEM_ENTS) {
            for (int i = tid; i < E; i += SEL_THREADS) sbuf[i] = G[i];
            A = sbuf;
            in_smem = true;
        }
    }
    if (tid == 0) S.pending = S.q_tail;
    __syncthreads();

    SEL_MARK(3);
    // ---- phase 2: warps drain the queue ------------------------------------------------------------------
    {
        SelRange t;
#ifdef YAVO_SEL_TIMING
        long long t_pop = 0, t_work = 0, n_task = 0, c0 = clock64();
        for (;;) {
            const bool got = sel_pop(S, t);
            const long long c1 = clock64();
            t_pop += c1 - c0;
            if (!got) break;
            sel_warp_work(S, A, t, K);
            c0 = clock64();
            t_work += c0 - c1;
            n_task++;
        }
        if (lane == 0 && f < 8) {
            long long *dbg = reinterpret_cast<long long *>(scratch_all) + 64 * 8 + (f * SEL_WARPS + warp) * 4;
            dbg[0] = t_pop; dbg[1] = t_work; dbg[2] = n_task; dbg[3] = 0;
        }
#else
        while (sel_pop(S, t)) sel_warp_work(S, A, t, K);
#endif
    }
    __syncthreads();
    if (tid == 0 && S.watchdog) atomicExch(status, 2);
    SEL_MARK(4);

    if (!solo) {
        // the member that finishes last sees every member's ranges in place and writes the outputs
        __threadfence();
        if (tid == 0) S.bcast[1] = atomicAdd(&team_done[f], 1);
        __syncthreads();
        if (S.bcast[1] != team - 1) return;
        __threadfence();  // lines shared with neighbouring ranges may sit in this SM's L1 with other members' old entries
    }
    // ---- outputs: first min(N,K) in order, plus the checkBoundry-compacted list -------------------
    const int nout = min(N, K);
    int run = 0;
    for (int base = 0; base < nout; base += SEL_THREADS) {
        const int i = base + tid;
        bool ok = false;
        int row = 0, col = 0;
        float sc = 0.f;
        if (i < nout) {
            const yavo_ent e = A[i];
            row = (int)((uint32_t)e >> 16);
            col = (int)((uint32_t)e & 0xffffu);
            sc = yavo_ent_score(e);
            kp_row[(size_t)f * kp_stride + i] = row;
            kp_col[(size_t)f * kp_stride + i] = col;
            kp_score[(size_t)f * kp_stride + i] = sc;
            ok = brief_admits(row, col, H, W);
        }
        const unsigned b = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) S.wtot[warp] = __popc(b);
        __syncthreads();
        int pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < SEL_WARPS; w++) {
            const int t = S.wtot[w];
            if (w < warp) pre += t;
            tot += t;
        }
        if (ok) {
            const int j = run + pre + __popc(b & ((1u << lane) - 1u));
            bk_row[(size_t)f * kp_stride + j] = row;
            bk_col[(size_t)f * kp_stride + j] = col;
            bk_score[(size_t)f * kp_stride + j] = sc;
            bk_id[(size_t)f * kp_stride + j] = i;
        }
        run += tot;
        __syncthreads();
    }
    if (tid == 0) { nkp[f] = nout; nbk[f] = run; }
#ifdef YAVO_SEL_TIMING
    SEL_MARK(5);
    if (tid == 0 && f < 64)
        for (int i = 0; i < 6; i++) reinterpret_cast<long long *>(scratch_all)[f * 8 + i] = tmark[i];
#endif
}

// ================================================================================================
// K4  BRIEF  (reference src/BriefDescriptor.cc:86-124)
// A warp handles BP_KPW keypoints one after the other; for each it stages the 17 x 17 smoothed neighbourhood
// in shared memory, lane l evaluates tests l, l+32, ..., l+224 and a ballot per group of 32 tests yields
// descriptor word w directly (bit j of the descriptor is bit
// j%32 of word j/32, i.e. byte j/8 bit j%8 little-endian — the reference layout).
// Reads follow Image::getPixelVal's unchecked linear indexing (src/Image.cc:15-17): a column index
// equal to W wraps to column 0 of the next row; a linear index >= H*W (undefined behaviour in the
// reference) reads as 0 and is counted.
// ================================================================================================
#ifndef YAVO_K4_THREADS
#define YAVO_K4_THREADS 256
#endif
constexpr int K4_THREADS = YAVO_K4_THREADS;

__device__ __forceinline__ int brief_sample(const uint8_t *S, int pitch, int H, int W, int r, int c, bool *oob) {
    if (c >= W) { c -= W; r += 1; }
    if (r >= H) { *oob = true; return 0; }
    return (int)__ldg(S + (size_t)r * pitch + c);
}

// offsets packed one test per word: byte0 = drow1, byte1 = dcol1, byte2 = drow2, byte3 = dcol2 (int8);
// spos: the same tests as byte positions inside a staged 17 x 32-byte patch (lo16 = first sample, hi16 = second),
// both tables prepared by the host when the offset table is set.
constexpr int BP_ROWB = 48;              // bytes per staged patch row: the tensor copy starts at a 16-byte aligned column (measured on
                                         // B200, tools/microbench/tma_probe.cu: any other innermost coordinate raises "illegal
                                         // instruction"), so the 17 columns col-8 .. col+8 start 0..15 bytes into the row
constexpr int BP_ROWS = 17;              // rows row-8 .. row+8
constexpr int BP_BYTES = BP_ROWS * BP_ROWB;                  // 816 bytes arrive per patch
constexpr int BP_ENTRY = (BP_BYTES + 127) & ~127;            // ring entries are 128-byte aligned (tensor-copy destination)
#ifndef YAVO_BP_KPW
#define YAVO_BP_KPW 8
#endif
constexpr int BP_KPW = YAVO_BP_KPW;      // keypoints per warp
#ifndef YAVO_BP_D
#define YAVO_BP_D 4
#endif
constexpr int BP_D = YAVO_BP_D;          // patches in flight per warp (ring of tensor copies): 4 or 8

// The patch of an interior keypoint arrives by ONE tensor copy (TMA: box of 48 x 17 bytes at ((col-8) & ~15, row-8) of
// the blurred plane), tracked by an mbarrier per ring entry: no per-lane gather loads, and a warp keeps BP_D patches
// in flight.  The keypoint's phase (col-8) & 15 is warp-uniform (built from ballots, so it lives in a uniform
// register) and is added to the lane's fixed sample addresses.  What bounds the kernel is then the shared-memory sampling
// itself (16 byte loads per lane and keypoint at table-defined, i.e. random, bank positions).
#ifndef YAVO_K4_MIN_CTAS
#define YAVO_K4_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(K4_THREADS, YAVO_K4_MIN_CTAS)
brief_kernel(const __grid_constant__ CUtensorMap blur_map, int slot_base,
             const uint8_t *__restrict__ blur, size_t frame_stride, int pitch, int H, int W,
             const uint32_t *__restrict__ offs, const uint32_t *__restrict__ spos, const int32_t *__restrict__ rows,
             const int32_t *__restrict__ cols, const int *__restrict__ n_per_frame, int n_fixed,
             int kp_stride, uint32_t *__restrict__ desc, uint8_t *__restrict__ valid,
             int *__restrict__ n_oob) {
    __shared__ __align__(128) uint8_t patch[K4_THREADS / 32][BP_D][BP_ENTRY];
    __shared__ __align__(8) uint64_t pbar[K4_THREADS / 32][BP_D];
    const int f = blockIdx.y;
    const int n = n_per_frame ? n_per_frame[f] : n_fixed;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kp0 = (blockIdx.x * (K4_THREADS / 32) + warp) * BP_KPW;
    if (kp0 >= n) return;
    const int nk = min(BP_KPW, n - kp0);
    if (lane < BP_D) mbar_init(&pbar[warp][lane], 1);
    // this lane's eight tests (j = 32 w + lane) as shared-memory byte addresses inside ring entry 0 of this warp
    uint32_t sa[8], sb[8];
    {
        const uint32_t pbase = smem_addr(&patch[warp][0][0]);
#pragma unroll
        for (int w = 0; w < 8; w++) {
            const uint32_t ps = __ldg(spos + 32 * w + lane);
            sa[w] = pbase + (ps & 0xffffu);
            sb[w] = pbase + (ps >> 16);
        }
    }
    // coordinates of this warp's keypoints (lanes 0..nk-1 load, shuffles broadcast)
    int my_row = 0, my_col = 0;
    if (lane < nk) {
        my_row = rows[(size_t)f * kp_stride + kp0 + lane];
        my_col = cols[(size_t)f * kp_stride + kp0 + lane];
    }
    // admitted by checkBoundry / interior (no sample can wrap to the next row or leave the buffer): bit i = keypoint i
    const unsigned adm = __ballot_sync(0xffffffffu, lane < nk && brief_admits(my_row, my_col, H, W));
    const unsigned inter = __ballot_sync(0xffffffffu, lane < nk && my_row + 8 < H && my_col + 8 < W) & adm;
    // phase bits of every keypoint as ballots: bit i of phb[b] = bit b of (col_i - 8) & 15
    unsigned phb[4];
#pragma unroll
    for (int b = 0; b < 4; b++) phb[b] = __ballot_sync(0xffffffffu, ((my_col - 8) >> b) & 1);
    __syncwarp();  // barriers initialised
    auto issue = [&](int i) {  // lane i issues the copy of keypoint i's patch
        if (lane == i && ((inter >> i) & 1u)) {
            uint64_t *bar = &pbar[warp][i % BP_D];
            mbar_expect_tx(bar, BP_BYTES);
            tma_load_3d(&patch[warp][i % BP_D][0], &blur_map, (my_col - 8) & ~15, my_row - 8, slot_base + f, bar);
        }
    };
#pragma unroll
    for (int i = 0; i < BP_D; i++) issue(i);
    const uint8_t *S = blur + (size_t)f * frame_stride;
    bool oob = false;
#pragma unroll
    for (int i = 0; i < BP_KPW; i++) {
        if (i >= nk) break;
        const int kp = kp0 + i;
        uint32_t *d = desc + ((size_t)f * kp_stride + kp) * 8;
        const bool ok = (adm >> i) & 1u;
        if (valid && lane == 0) valid[(size_t)f * kp_stride + kp] = ok ? 1 : 0;
        uint32_t mine = 0;
        if ((inter >> i) & 1u) {
            // parity of the entry's barrier = earlier copies into the same entry (keypoints i-BP_D, i-2*BP_D, ... that took this path)
            constexpr uint32_t same_entry = BP_D == 4 ? 0x11111111u : 0x01010101u;
            static_assert(BP_D == 4 || BP_D == 8, "same_entry mask");
            const uint32_t par = __popc(inter & (same_entry << (i % BP_D)) & ((1u << i) - 1u)) & 1u;
            if (lane == 0) mbar_wait(&pbar[warp][i % BP_D], par);
            __syncwarp();
            const uint32_t so = (uint32_t)((i % BP_D) * BP_ENTRY) +  // compile-time after unrolling: an immediate offset
                                (((phb[0] >> i) & 1u) | (((phb[1] >> i) & 1u) << 1) | (((phb[2] >> i) & 1u) << 2) | (((phb[3] >> i) & 1u) << 3));
#pragma unroll
            for (int w = 0; w < 8; w++) {
                uint32_t va, vb;
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(va) : "r"(sa[w] + so) : "memory");
                asm volatile("ld.shared.u8 %0, [%1];" : "=r"(vb) : "r"(sb[w] + so) : "memory");
                const unsigned word = __ballot_sync(0xffffffffu, va > vb);  // consumes the samples of every lane
                if (lane == w) mine = word;
            }
        } else if (ok) {
            const int row = __shfl_sync(0xffffffffu, my_row, i), col = __shfl_sync(0xffffffffu, my_col, i);
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint32_t o = __ldg(offs + 32 * w + lane);
                const int r1 = row + (int)(int8_t)(o & 0xff), c1 = col + (int)(int8_t)((o >> 8) & 0xff);
                const int r2 = row + (int)(int8_t)((o >> 16) & 0xff), c2 = col + (int)(int8_t)(o >> 24);
                const int va = brief_sample(S, pitch, H, W, r1, c1, &oob);
                const int vb = brief_sample(S, pitch, H, W, r2, c2, &oob);
                const unsigned word = __ballot_sync(0xffffffffu, va > vb);
                if (lane == w) mine = word;
            }
            if (n_oob) {
                const unsigned any = __ballot_sync(0xffffffffu, oob);
                if (lane == 0 && any) atomicAdd(n_oob, 1);
                oob = false;
            }
        }
        if (lane < 8) d[lane] = mine;  // not admitted: zeros
        // every lane's samples of this ring entry have been used (the ballots consumed them): it may be overwritten
        if (i + BP_D < BP_KPW) issue(i + BP_D);
    }
}

// ================================================================================================
// K4b  result pack of the single-frame path: everything a caller of getFastFeatures + computeBrief reads back, gathered
// into one contiguous buffer so that it crosses PCIe as ONE copy:
//   [n_kp, n_desc, n_cand, status | kp_row[K] | kp_col[K] | bk_row[K] | bk_col[K] | bk_id[K] | kp_score[K] | desc[8K words]]
// ================================================================================================
__global__ void pack_frame_kernel(const int *__restrict__ nkp, const int *__restrict__ nbk, const int *__restrict__ ncand,
                                  const int *__restrict__ status, const int32_t *__restrict__ kp_row,
                                  const int32_t *__restrict__ kp_col, const float *__restrict__ kp_score,
                                  const int32_t *__restrict__ bk_row, const int32_t *__restrict__ bk_col,
                                  const int32_t *__restrict__ bk_id, const uint32_t *__restrict__ desc, int K,
                                  uint32_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        out[0] = (uint32_t)nkp[0];
        out[1] = (uint32_t)nbk[0];
        out[2] = (uint32_t)ncand[0];
        out[3] = (uint32_t)status[0];
    }
    const int n = nkp[0], nb = nbk[0];
    uint32_t *o = out + 4;
    if (i < K) {
        if (i < n) {
            o[i] = (uint32_t)kp_row[i];
            o[K + i] = (uint32_t)kp_col[i];
            o[5 * K + i] = __float_as_uint(kp_score[i]);
        }
        if (i < nb) {
            o[2 * K + i] = (uint32_t)bk_row[i];
            o[3 * K + i] = (uint32_t)bk_col[i];
            o[4 * K + i] = (uint32_t)bk_id[i];
        }
    }
    if (i < 8 * K && i < 8 * nb) o[6 * K + i] = desc[i];
}

// ================================================================================================
// K5  Hamming match  (reference src/BriefDescriptor.cc:139-183)
// Brute force over 256-bit descriptors held as 8 x u32.  A CTA owns MQ queries (one per thread,
// descriptor in registers) and one chunk of the train set, staged through shared memory in tiles
// and read as broadcast 128-bit loads.  Every thread keeps key = dist << 22 | j, whose minimum
// (one VIMNMX) is the reference's "first minimum wins" rule; the second smallest distance is
// tracked only when the caller asks for it (SECOND).  Chunk partials are combined by a small
// reduce kernel (deterministic, no atomics).
//
// Pipe balance: the POPC (XU) pipe issues 16 lanes/clk/SM, the integer ALU 64, the FMA pipe 64.  A plain
// 8 x (XOR, POPC) per pair is XU-bound at 2 pairs/clk/SM (measured 1.75, ncu: XU 91 %).  Carry-save
// adders (2 LOP3 each) first compress the eight XOR words: with three CSAs a pair costs 5 POPC + 14 LOP3
// (measured 2.81 pairs/clk/SM, XU 93 %, ALU 79 %); with four CSAs 4 POPC + 16 LOP3, and with the
// popcount sums and the key built by IMADs on the otherwise idle FMA pipe the ALU (17/64 clk) and XU
// (4/16 clk) are balanced at a bound of ~3.8 pairs/clk/SM.
// ================================================================================================
constexpr int MQ = 128;        // queries per CTA == threads
constexpr int MT = 128;        // train descriptors per shared-memory tile
constexpr uint32_t MATCH_NONE = 0xffffffffu;

// carry-save adder on 32 independent bit columns: a + b + c = sum + 2 * carry
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t &sum, uint32_t &carry) {
    sum = a ^ b ^ c;                  // LOP3 0x96
    carry = (a & b) | (c & (a ^ b));  // LOP3 0xE8
}

// a * b + c on the FMA pipe (IMAD): keeps the additions off the integer ALU, which the LOP3s saturate
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// 256-bit Hamming distance: 8 XOR, four carry-save adders (8 LOP3) leave two weight-1 words, one weight-2
// and one weight-4 word -> 4 POPC, summed with IMADs.  Per pair: ALU 8+8 (+1 min), XU 4, FMA 4.
__device__ __forceinline__ uint32_t hamming256(const uint4 &a0, const uint4 &a1, const uint4 &b0, const uint4 &b1) {
    const uint32_t x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
    const uint32_t x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
    uint32_t s1, c1, s2, c2, s3, c3, s4, c4;
    csa(x0, x1, x2, s1, c1);
    csa(x3, x4, x5, s2, c2);
    csa(s1, s2, x6, s3, c3);
    csa(c1, c2, c3, s4, c4);
    return imad(__popc(c4), 4u, imad(__popc(s4), 2u, imad(__popc(s3), 1u, __popc(x7))));
}

template <bool SECOND>
__global__ void __launch_bounds__(MQ)
match_partial_kernel(const uint32_t *__restrict__ dq_all, const int *__restrict__ nq_all, int nq_fixed,
                     const uint32_t *__restrict__ dt_all, const int *__restrict__ nt_all, int nt_fixed,
                     size_t set_stride_words, int q_set_offset, int t_set_offset, int chunk,
                     int n_chunks, int out_stride, uint32_t *__restrict__ part_key,
                     uint32_t *__restrict__ part_sec, uint32_t key_mul /* 1 << 22, passed at run time so that
                     key = d * key_mul + j stays an IMAD on the FMA pipe instead of a LEA on the ALU */) {
    __shared__ uint4 st[MT][2];
    const int pair = blockIdx.z;
    const uint32_t *dq = dq_all + (size_t)(pair + q_set_offset) * set_stride_words;
    const uint32_t *dt = dt_all + (size_t)(pair + t_set_offset) * set_stride_words;
    const int nq = nq_all ? nq_all[pair + q_set_offset] : nq_fixed;
    const int nt = nt_all ? nt_all[pair + t_set_offset] : nt_fixed;
    const int q0 = blockIdx.x * MQ;
    if (q0 >= nq) return;
    const int c = blockIdx.y;
    const int j0 = c * chunk, j1 = min(nt, j0 + chunk);
    const int q = q0 + threadIdx.x;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (q < nq) {
        a0 = __ldg(reinterpret_cast<const uint4 *>(dq + (size_t)q * 8));
        a1 = __ldg(reinterpret_cast<const uint4 *>(dq + (size_t)q * 8) + 1);
    }
    uint32_t best = MATCH_NONE, sec = MATCH_NONE;  // best: packed key; sec: distance only
    for (int t0 = j0; t0 < j1; t0 += MT) {
        const int nt_tile = min(MT, j1 - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < nt_tile * 2; i += MQ)
            st[i >> 1][i & 1] = __ldg(reinterpret_cast<const uint4 *>(dt + (size_t)t0 * 8) + i);
        __syncthreads();
        // warps whose 32 queries all lie beyond nq (three of the four warps of a frame's last CTA at ~1950
        // keypoints) only help staging the tiles: their issue slots go to the other resident CTAs
        if (q0 + (int)(threadIdx.x & ~31u) >= nq) continue;
        if (SECOND) {
#pragma unroll 4
            for (int j = 0; j < nt_tile; j++) {
                const uint32_t d = hamming256(a0, a1, st[j][0], st[j][1]);
                const uint32_t key = imad(d, key_mul, (uint32_t)(t0 + j));
                // ascending j and strict '<' == min over key; sec = second smallest distance
                if (key < best) {
                    sec = (best == MATCH_NONE) ? sec : (best >> 22);
                    best = key;
                } else if (d < sec) {
                    sec = d;
                }
            }
        } else {
            // two pairs per step: one 3-input minimum (VIMNMX3) per two keys, keys built by IMAD (FMA pipe)
            const int even = nt_tile & ~1;
#pragma unroll 2
            for (int j = 0; j < even; j += 2) {
                const uint32_t d0 = hamming256(a0, a1, st[j][0], st[j][1]);
                const uint32_t d1 = hamming256(a0, a1, st[j + 1][0], st[j + 1][1]);
                best = __vimin3_u32(best, imad(d0, key_mul, (uint32_t)(t0 + j)), imad(d1, key_mul, (uint32_t)(t0 + j + 1)));
            }
            if (even < nt_tile) {
                const uint32_t d = hamming256(a0, a1, st[even][0], st[even][1]);
                best = min(best, imad(d, key_mul, (uint32_t)(t0 + even)));
            }
        }
    }
    if (q < nq) {
        part_key[((size_t)pair * out_stride + q) * n_chunks + c] = best;
        if (SECOND) part_sec[((size_t)pair * out_stride + q) * n_chunks + c] = sec;
    }
}

__global__ void match_reduce_kernel(const uint32_t *__restrict__ part_key, const uint32_t *__restrict__ part_sec,
                                    const int *__restrict__ nq_all, int nq_fixed, int q_set_offset,
                                    const int *__restrict__ nt_all, int nt_fixed, int t_set_offset, int chunk,
                                    int n_chunks, int out_stride, int32_t *__restrict__ out_idx,
                                    int32_t *__restrict__ out_dist, int32_t *__restrict__ out_second) {
    const int pair = blockIdx.y;
    const int nq = nq_all ? nq_all[pair + q_set_offset] : nq_fixed;
    const int nt = nt_all ? nt_all[pair + t_set_offset] : nt_fixed;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int used = (nt + chunk - 1) / chunk;  // chunks that held any train descriptor
    uint32_t best = MATCH_NONE, sec = MATCH_NONE;
    for (int c = 0; c < used && c < n_chunks; c++) {
        const uint32_t k = part_key[((size_t)pair * out_stride + q) * n_chunks + c];
        const uint32_t s = out_second ? part_sec[((size_t)pair * out_stride + q) * n_chunks + c] : MATCH_NONE;
        if (k == MATCH_NONE) continue;
        const uint32_t kd = k >> 22;
        if (k < best) {
            // old best distance becomes a second-best candidate, as does this chunk's second
            uint32_t ns = (best == MATCH_NONE) ? sec : min(sec, best >> 22);
            sec = min(ns, s);
            best = k;
        } else {
            sec = min(sec, min(kd, s));
        }
    }
    const size_t o = (size_t)pair * out_stride + q;
    if (best == MATCH_NONE) {  // empty train set: reference leaves kp2 = (0,0,0), distance INT_MAX
        out_idx[o] = -1;
        out_dist[o] = 0x7fffffff;
        if (out_second) out_second[o] = 0x7fffffff;
    } else {
        out_idx[o] = (int32_t)(best & 0x3fffffu);
        out_dist[o] = (int32_t)(best >> 22);
        if (out_second) out_second[o] = (sec == MATCH_NONE) ? 0x7fffffff : (int32_t)sec;
    }
}

// ================================================================================================
// K6  outlier filter + point pairs  (reference src/BriefDescriptor.cc:213-231 and the conversion of the
// kept matches to point pairs in src/LoopHandler.cc:232-237,251-254)
// One CTA per frame pair (f-1, f): min-reduce of the match distances, keep distance < max(2*min, threshold)
// in order, and write the kept matches as compact point pairs — what cv::findEssentialMat consumes — so a
// caller that only needs the filtered correspondences never downloads descriptors or the full match list.
// out_pairs rows: {q_row, q_col, t_row, t_col, dist, q_index, t_index, 0}
// ================================================================================================
constexpr int K6_THREADS = 256;

__global__ void __launch_bounds__(K6_THREADS)
filter_pairs_kernel(const int32_t *__restrict__ midx, const int32_t *__restrict__ mdist,
                    const int32_t *__restrict__ bk_row, const int32_t *__restrict__ bk_col,
                    const int *__restrict__ nbk, int kp_stride, int threshold, int32_t *__restrict__ out_pairs,
                    int *__restrict__ out_n, int *__restrict__ out_min) {
    __shared__ int red[K6_THREADS / 32];
    __shared__ int s_min;
    const int p = blockIdx.x;  // pair index: queries = slot p, train = slot p+1, matches stored at slot p+1
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nq = nbk[p], nt = nbk[p + 1];
    const int32_t *mi = midx + (size_t)(p + 1) * kp_stride, *md = mdist + (size_t)(p + 1) * kp_stride;
    int mn = 0x7fffffff;
    for (int i = tid; i < nq; i += K6_THREADS) mn = min(mn, md[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if (lane == 0) red[warp] = mn;
    __syncthreads();
    if (tid == 0) {
        int m = 0x7fffffff;
        for (int w = 0; w < K6_THREADS / 32; w++) m = min(m, red[w]);
        s_min = m;
    }
    __syncthreads();
    const int lim = max((int)(2u * (unsigned)s_min), threshold);  // int arithmetic of the reference: 2 * INT_MAX wraps to -2
    int run = 0;
    for (int base = 0; base < nq; base += K6_THREADS) {
        const int i = base + tid;
        const bool keep = i < nq && nt > 0 && md[i] < lim;
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        __syncthreads();
        if (lane == 0) red[warp] = __popc(b);
        __syncthreads();
        int pre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < K6_THREADS / 32; w++) {
            if (w < warp) pre += red[w];
            tot += red[w];
        }
        if (keep) {
            const int j = run + pre + __popc(b & ((1u << lane) - 1u));
            const int t = mi[i];
            int32_t *o = out_pairs + ((size_t)(p + 1) * kp_stride + j) * 8;
            o[0] = bk_row[(size_t)p * kp_stride + i];
            o[1] = bk_col[(size_t)p * kp_stride + i];
            o[2] = bk_row[(size_t)(p + 1) * kp_stride + t];
            o[3] = bk_col[(size_t)(p + 1) * kp_stride + t];
            o[4] = md[i];
            o[5] = i;
            o[6] = t;
            o[7] = 0;
        }
        run += tot;
    }
    if (tid == 0) {
        out_n[p + 1] = run;
        out_min[p + 1] = s_min;
    }
}

}  // namespace yavo
