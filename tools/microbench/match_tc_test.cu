// Bring-up / timing tool for the tensor-core Hamming matcher (ya_vo_b200/csrc/match_tc.cuh).
// Builds random descriptor sets with planted near-duplicates and exact duplicates (ties), runs the kernel and
// compares every (index, distance) with a scalar popcount loop on the host (first minimum wins).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o match_tc_test match_tc_test.cu
//   ./match_tc_test [pairs=8] [n=1950] [verify=1] [reps=5]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../../ya_vo_b200/csrc/match_tc4.cuh"

#ifdef TC4   // packed 4-bit operands (kind::mxf4), 128 x 224 tiles
#define KERNEL yavo::tcm4::match_tc4_kernel<true>
constexpr int TILE_N = yavo::tcm4::T4, SMEM = yavo::tcm4::SMEM4_BYTES, NTHR = yavo::tcm4::THREADS4;
#else        // FP8 operands (kind::f8f6f4), 128 x 256 tiles
#define KERNEL yavo::tcm::match_tc_kernel<true>
constexpr int TILE_N = yavo::tcm::TT, SMEM = yavo::tcm::SMEM_BYTES, NTHR = yavo::tcm::THREADS;
#endif

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                      \
        }                                                                                  \
    } while (0)

int main(int argc, char **argv) {
    const int sets = (argc > 1 ? atoi(argv[1]) : 8) + 1;
    const int nmax = argc > 2 ? atoi(argv[2]) : 1950;
    const int verify = argc > 3 ? atoi(argv[3]) : 1;
    const int reps = argc > 4 ? atoi(argv[4]) : 5;
    const int stride = ((nmax + 7) / 8) * 8 + 50;  // descriptors per set slot
    std::mt19937 rng(1234);
    std::vector<uint32_t> desc((size_t)sets * stride * 8);
    std::vector<int> n(sets);
    for (int s = 0; s < sets; s++) {
        n[s] = (s == 3 && sets > 4) ? 0 : (s == 5 && sets > 6) ? 7 : nmax - (int)(rng() % 60);
        if (n[s] < 0) n[s] = 0;
        for (int i = 0; i < stride; i++)
            for (int w = 0; w < 8; w++) desc[((size_t)s * stride + i) * 8 + w] = rng();
        if (s > 0)
            for (int i = 0; i < n[s]; i++) {
                const int kind = rng() % 4;
                if (kind < 2 && n[s - 1] > 0) {  // near-duplicate / duplicate of a descriptor of the previous set
                    const int src = rng() % n[s - 1];
                    memcpy(&desc[((size_t)s * stride + i) * 8], &desc[((size_t)(s - 1) * stride + src) * 8], 32);
                    const int flips = kind == 0 ? 0 : rng() % 12;
                    for (int f = 0; f < flips; f++) {
                        const int b = rng() % 256;
                        desc[((size_t)s * stride + i) * 8 + b / 32] ^= 1u << (b % 32);
                    }
                }
            }
    }
    const int pairs = sets - 1, q_tiles = (stride + yavo::tcm::TQ - 1) / yavo::tcm::TQ;
    uint32_t *d_desc;
    int *d_n;
    int32_t *d_idx, *d_dist;
    float *d_dots;
    CK(cudaMalloc(&d_desc, desc.size() * 4));
    CK(cudaMalloc(&d_n, sets * 4));
    CK(cudaMalloc(&d_idx, (size_t)pairs * stride * 4));
    CK(cudaMalloc(&d_dist, (size_t)pairs * stride * 4));
    CK(cudaMalloc(&d_dots, 128 * 256 * 4));
    CK(cudaMemcpy(d_desc, desc.data(), desc.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_n, n.data(), sets * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_idx, 0xee, (size_t)pairs * stride * 4));
    CK(cudaMemset(d_dist, 0xee, (size_t)pairs * stride * 4));
    CK(cudaMemset(d_dots, 0, 128 * 256 * 4));
    CK(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = std::min(sms, pairs * q_tiles);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < reps; rep++) {
        CK(cudaEventRecord(e0));
        KERNEL<<<grid, NTHR, SMEM>>>(
            d_desc, d_n, 0, d_desc, d_n, 0, (size_t)stride * 8, 0, 1, pairs, q_tiles, stride, d_idx, d_dist,
            rep == 0 ? d_dots : nullptr);
        CK(cudaEventRecord(e1));
        CK(cudaGetLastError());
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 || reps == 1) best_ms = std::min(best_ms, ms);
    }
    std::vector<int32_t> idx((size_t)pairs * stride), dist((size_t)pairs * stride);
    std::vector<float> dots(128 * 256);
    CK(cudaMemcpy(idx.data(), d_idx, idx.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dist.data(), d_dist, dist.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dots.data(), d_dots, dots.size() * 4, cudaMemcpyDeviceToHost));
    long long bad = 0, checked = 0, npairs = 0, dot_bad = 0;
    for (int p = 0; p < pairs; p++) npairs += (long long)n[p] * n[p + 1];
    for (int p = 0; p < pairs; p++) {
        if (!verify && p > 0) break;
        const uint32_t *Q = &desc[(size_t)p * stride * 8], *T = &desc[(size_t)(p + 1) * stride * 8];
        if (p == 0)
            for (int i = 0; i < std::min(128, n[0]); i++)
                for (int j = 0; j < std::min(TILE_N, n[1]); j++) {
                    int d = 0;
                    for (int w = 0; w < 8; w++) d += __builtin_popcount(Q[i * 8 + w] ^ T[j * 8 + w]);
                    if (dots[i * TILE_N + j] != (float)(256 * d + j - 32768)) {
                        if (dot_bad < 8) printf("acc[%d][%d] = %g, expected %d\n", i, j, dots[i * TILE_N + j], 256 * d + j - 32768);
                        dot_bad++;
                    }
                }
        for (int i = 0; i < n[p]; i++) {
            int bd = 0x7fffffff, bj = -1;
            for (int j = 0; j < n[p + 1]; j++) {
                int d = 0;
                for (int w = 0; w < 8; w++) d += __builtin_popcount(Q[i * 8 + w] ^ T[j * 8 + w]);
                if (d < bd) bd = d, bj = j;
            }
            checked++;
            if (idx[(size_t)p * stride + i] != bj || dist[(size_t)p * stride + i] != bd) {
                if (bad < 8)
                    printf("pair %d query %d: got (%d, %d), expected (%d, %d)\n", p, i, idx[(size_t)p * stride + i],
                           dist[(size_t)p * stride + i], bj, bd);
                bad++;
            }
        }
        // rows beyond the set must stay untouched
        for (int i = n[p]; i < stride; i++)
            if (idx[(size_t)p * stride + i] != (int32_t)0xeeeeeeee) bad++;
    }
    printf("{\"verified\": %d, \"pairs\": %d, \"n\": %d, \"queries_checked\": %lld, \"mismatches\": %lld, \"dot_mismatches\": %lld, "
           "\"ms\": %.4f, \"gpairs_per_s\": %.1f, \"grid\": %d}\n",
           verify, pairs, nmax, checked, bad, dot_bad, best_ms, npairs / (best_ms * 1e6), grid);
    return bad ? 1 : 0;
}
