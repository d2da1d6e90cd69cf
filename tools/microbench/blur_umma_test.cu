// blur_umma_test.cu — bring-up of the 9x9 fixed-point Gaussian (reference src/BriefDescriptor.cc:90, OpenCV's u8 path)
// on the 5th-generation tensor cores: both passes as banded Toeplitz products with tcgen05.mma kind::i8.
//
//   pass 1 (horizontal):  D1[x, n]  = sum_c  T1[x, c] * P[n, c]      M = 128 outputs x, N = 48 staged rows (40 used),
//                                                                    K = 160 staged bytes of a row (5 instructions)
//       A = T1: row x holds the nine taps at staged bytes x+12 .. x+20.  The K slice k of T1 is the SAME 256 x 32 band
//           matrix F read from row 128-32k on (A_k[x][j] = F[x+128-32k][j], F[r][j] = g[j+116-r]), so 8 KB of
//           constants serve all five instructions (K-major, no swizzle: the row shift is a start-address shift).
//       B = the staged pixel rows in K-major core-matrix order [16-byte chunk][row][16].
//   D1 (int32, < 2^16) is split into low / high bytes by the CTA's warps and written back to TMEM as the A operand of
//   pass 2 (vertical):    D2lo/hi[x, r] = sum_n  Hlo/hi[x, n] * T2[r, n]     M = 128, N = 16 output rows per pass,
//                                                                             K = 32 staged rows (window r0 .. r0+31)
//       out(r, x) = (D2lo + 256 * D2hi + 32768) >> 16 — exact integer arithmetic, bit-identical to the separable
//       u16 / u32 form OpenCV runs.
//   TMEM: 64 columns per CTA (eight CTAs per SM): D1 at [0,48), A_lo at [40,52) (overlaps only D1's unused pad
//   rows), A_hi at [52,64), D2 lo / hi of a pass at [0,16) / [16,32).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o blur_umma_test blur_umma_test.cu
//   ./blur_umma_test [frames=4] [reps=3]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../ya_vo_b200/csrc/blur_umma.cuh"

using namespace yavo::bu;

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                      \
        }                                                                                  \
    } while (0)

static int reflect101h(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * (n - 1) - p;
    return p;
}

__global__ void __launch_bounds__(256, 8)
blur_only_kernel(const uint8_t *__restrict__ frames, size_t frame_stride, int pitch, int H, int W, uint8_t *__restrict__ blur,
                 const uint8_t *__restrict__ consts, long long *__restrict__ clk) {
    __shared__ __align__(128) uint8_t tile[BU_SH * BU_SROW];
    __shared__ __align__(128) uint8_t ub[BU_UB_BYTES];
    __shared__ __align__(128) uint8_t cst[BU_CONST_BYTES];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, f = blockIdx.z, x0 = blockIdx.x * 128, y0 = blockIdx.y * 32;
    const uint8_t *img = frames + (size_t)f * frame_stride;
    long long t[8];
    t[0] = clock64();
    bu_prologue(cst, consts, bars, &tmem_base_s);
    // stage (plain loads; the product kernel uses its tensor copies)
    for (int i = tid; i < BU_SH * (BU_SROW / 4); i += 256) {
        const int tr = i / (BU_SROW / 4), wq = i % (BU_SROW / 4);
        const int gr = (y0 - 4 + tr < 0) ? -(y0 - 4 + tr) : (y0 - 4 + tr >= H ? 2 * (H - 1) - (y0 - 4 + tr) : y0 - 4 + tr);
        uint32_t v = 0;
        for (int b = 0; b < 4; b++) {
            int gc = x0 - 16 + 4 * wq + b;
            if (gc < 0) gc = -gc;
            if (gc >= W) gc = 2 * (W - 1) - gc;
            if (gc < 0) gc = 0;
            v |= (uint32_t)img[(size_t)gr * pitch + gc] << (8 * b);
        }
        reinterpret_cast<uint32_t *>(tile)[i] = v;
    }
    __syncthreads();
    bu_relayout(tile, ub);
    __syncthreads();
    const uint32_t tmem_base = tmem_base_s;
    t[1] = clock64();
    bu_pass1_issue(ub, cst, bars, tmem_base);
    t[2] = clock64();
    if (tid == 0) bu_bar_wait(bu_saddr(&bars[1]), 0);
    t[3] = clock64();
    bu_pass1_drain(cst, bars, tmem_base);
    t[4] = clock64();
    if (tid == 0) bu_bar_wait(bu_saddr(&bars[1]), 1);
    t[5] = clock64();
    bu_pass2_drain(cst, bars, tmem_base, ub, 0);
    t[6] = clock64();
    bu_pass2_drain(cst, bars, tmem_base, ub, 1);
    bu_finish(tmem_base, ub, blur + (size_t)f * frame_stride, pitch, H, x0, y0);
    t[7] = clock64();
    if (clk && tid == 0 && blockIdx.x == 3 && blockIdx.y == 5 && (f == 0 || f == gridDim.z / 2))
        for (int i = 0; i < 8; i++) clk[(f ? 8 : 0) + i] = t[i] - t[0];
}

int main(int argc, char **argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 4, reps = argc > 2 ? atoi(argv[2]) : 3;
    const int H = 376, W = 1241, pitch = 1280;
    const size_t fs = (size_t)pitch * H;
    std::vector<uint8_t> h(fs * B);
    uint32_t s = 12345;
    for (size_t i = 0; i < h.size(); i++) {
        s = s * 1664525u + 1013904223u;
        h[i] = (uint8_t)(s >> 24);
    }
    for (int r = 0; r < 40 && B > 1; r++)  // saturated block: the largest sums
        for (int c = 0; c < 300; c++) h[fs + (size_t)(100 + r) * pitch + 500 + c] = 255;
    std::vector<uint8_t> cst(BU_CONST_BYTES);
    bu_fill_constants(cst.data());
    uint8_t *d_f, *d_b, *d_c;
    long long *d_clk;
    CK(cudaMalloc(&d_clk, 16 * 8));
    CK(cudaMemset(d_clk, 0, 16 * 8));
    CK(cudaMalloc(&d_f, h.size()));
    CK(cudaMalloc(&d_b, h.size()));
    CK(cudaMalloc(&d_c, cst.size()));
    CK(cudaMemcpy(d_f, h.data(), h.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_c, cst.data(), cst.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_b, 0xee, h.size()));
    dim3 grid((W + 127) / 128, (H + 31) / 32, B);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e9f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        blur_only_kernel<<<grid, 256>>>(d_f, fs, pitch, H, W, d_b, d_c, d_clk);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    std::vector<uint8_t> g(h.size());
    CK(cudaMemcpy(g.data(), d_b, g.size(), cudaMemcpyDeviceToHost));
    const int taps[9] = {12, 22, 31, 41, 44, 41, 31, 22, 12};
    long bad = 0, checked = 0;
    const int nchk = B < 3 ? B : 3;
    for (int f = 0; f < nchk; f++) {
        std::vector<uint32_t> hs((size_t)H * W);
        for (int r = 0; r < H; r++)
            for (int c = 0; c < W; c++) {
                uint32_t a = 0;
                for (int k = 0; k < 9; k++) a += taps[k] * h[f * fs + (size_t)r * pitch + reflect101h(c + k - 4, W)];
                hs[(size_t)r * W + c] = a;
            }
        for (int r = 0; r < H; r++)
            for (int c = 0; c < W; c++) {
                uint32_t a = 32768;
                for (int k = 0; k < 9; k++) a += taps[k] * hs[(size_t)reflect101h(r + k - 4, H) * W + c];
                const uint8_t e = (uint8_t)(a >> 16), v = g[f * fs + (size_t)r * pitch + c];
                checked++;
                if (e != v) {
                    if (bad < 12) printf("mismatch f=%d r=%d c=%d exp=%d got=%d\n", f, r, c, e, v);
                    bad++;
                }
            }
    }
    long long hc[16];
    CK(cudaMemcpy(hc, d_clk, sizeof hc, cudaMemcpyDeviceToHost));
    printf("clocks (thread 0 of one CTA; start, staged+relayout, pass1 issued, pass1 done, drained+pass2a issued, pass2a done, half0 drained, end):\n");
    for (int k = 0; k < 2; k++) {
        for (int i = 0; i < 8; i++) printf(" %lld", hc[8 * k + i]);
        printf("\n");
    }
    printf("frames=%d  blur kernel %.4f ms (best of %d)  checked=%ld mismatches=%ld\n", B, best, reps, checked, bad);
    return bad ? 1 : 0;
}
