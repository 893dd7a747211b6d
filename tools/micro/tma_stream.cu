// TMA fill-rate microbenchmark: one thread per CTA streams boxes {BW, BH, BC} of an fp32 tensor [N, H, W] through a
// ring of NST shared-memory stages (wait full -> reissue), 1 CTA per SM.  Prints GB/s per box shape / ring depth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32, 1) stream_kernel(const __grid_constant__ CUtensorMap tmap, int nst, int stage_bytes, int iters,
                                                      int W, int H, int N, int bw, int bh, int bc, unsigned long long* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nst * stage_bytes);
    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // walk the tensor: CTA b starts at a different (plane group, row)
        const int rows_per_plane = H / bh, groups = N / bc;
        long long pos = (long long)blockIdx.x * 977;
        auto issue = [&](int s, long long q) {
            const int g = (int)((q / rows_per_plane) % groups), r = (int)(q % rows_per_plane);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(stage_bytes) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(smem + (size_t)s * stage_bytes)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(smem_u32(&bars[s])),
                           "r"(4 * (int)(q % 3)), "r"(r * bh), "r"(g * bc) : "memory");
        };
        for (int s = 0; s < nst; ++s) issue(s, pos + s);
        unsigned long long acc = 0;
        for (int it = 0; it < iters; ++it) {
            const int s = it % nst;
            const uint32_t par = (it / nst) & 1;
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bars[s])), "r"(par) : "memory");
            acc += *reinterpret_cast<volatile unsigned*>(smem + (size_t)s * stage_bytes);
            if (it + nst < iters) issue(s, pos + it + nst);
        }
        sink[blockIdx.x] = acc;
    }
}

int main() {
    const int W = 224, H = 224, N = 64 * 16;   // scale-2 f1 of the bench: 205 MB
    float* d; cudaMalloc(&d, (size_t)N * H * W * 4); cudaMemset(d, 0, (size_t)N * H * W * 4);
    unsigned long long* sink; cudaMalloc(&sink, 148 * 8 * 8);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &fn, 12000, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
    const int shapes[][3] = {{64, 32, 2}, {96, 36, 1}, {40, 40, 1}, {56, 40, 1}, {72, 40, 1}, {56, 32, 1}, {56, 64, 1}, {56, 8, 1}, {56, 40, 2}, {56, 20, 2}, {128, 32, 1}, {224, 16, 1}, {32, 40, 1}};
    for (auto& sh : shapes) {
        const int bw = sh[0], bh = sh[1], bc = sh[2];
        CUtensorMap tm;
        cuuint64_t gd[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t gs[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
        cuuint32_t bx[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bc}, es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed for %d %d %d\n", bw, bh, bc); continue; }
        const int stage_bytes = bw * bh * bc * 4;
        for (int ctas_per_sm = 1; ctas_per_sm <= 2; ++ctas_per_sm)
        for (int nst : {2, 4, 8}) {
            const size_t smem = (size_t)nst * stage_bytes + nst * 8;
            if (smem * ctas_per_sm > 220 * 1024) continue;
            cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            const int iters = (int)(8ll * 1024 * 1024 / stage_bytes);   // 8 MB per CTA
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            stream_kernel<<<148 * ctas_per_sm, 32, smem>>>(tm, nst, stage_bytes, iters, W, H, N, bw, bh, bc, sink);
            cudaEventRecord(a);
            stream_kernel<<<148 * ctas_per_sm, 32, smem>>>(tm, nst, stage_bytes, iters, W, H, N, bw, bh, bc, sink);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            cudaError_t e = cudaGetLastError();
            const double bytes = 148.0 * ctas_per_sm * iters * stage_bytes;
            printf("box {%3d,%2d,%2d} stage %6d B  ring %d x %d CTA/SM: %7.1f GB/s  (%.1f GB/s per SM)%s\n", bw, bh, bc, stage_bytes, nst, ctas_per_sm,
                   bytes / ms / 1e6, bytes / ms / 1e6 / 148, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
