// tma_probe.cu — bring-up probe: does a tiled tensor copy (cp.async.bulk.tensor.3d) accept an arbitrary (not 16-byte
// aligned) innermost coordinate for a u8 tensor, and which box shapes work?   usage: tma_probe X BOXW BOXH
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int z, int bytes, uint8_t *out) {
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(d),
                     "l"(reinterpret_cast<uint64_t>(&map)), "r"(x), "r"(y), "r"(z), "r"(b)
                     : "memory");
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b) : "memory");
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char **argv) {
    const int x = argc > 1 ? atoi(argv[1]) : 0, bw = argc > 2 ? atoi(argv[2]) : 32, bh = argc > 3 ? atoi(argv[3]) : 17;
    const int pitch = 1280, rows = 376, slots = 2, y = 100, z = 1;
    std::vector<uint8_t> h((size_t)pitch * rows * slots);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *o;
    cudaMalloc(&d, h.size());
    cudaMalloc(&o, bw * bh);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                              const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map;
    const cuuint64_t gd[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)slots}, gs[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * rows};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CUresult r = ((enc_t)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("x=%d box=%dx%d encode=%d ", x, bw, bh, (int)r);
    probe<<<1, 128, bw * bh + 128>>>(map, x, y, z, bw * bh, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("run=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<uint8_t> g(bw * bh);
        cudaMemcpy(g.data(), o, g.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int rr = 0; rr < bh; rr++)
            for (int c = 0; c < bw; c++) {
                const int gx = x + c;
                const uint8_t exp = (gx >= 0 && gx < pitch) ? h[((size_t)z * rows + y + rr) * pitch + gx] : 0;
                bad += g[rr * bw + c] != exp;
            }
        printf("mismatches=%d", bad);
    }
    printf("\n");
    return 0;
}
