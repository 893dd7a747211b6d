// Shared-memory load issue rate on sm_100a: conflict-free LDS.32 / LDS.64 / LDS.128 per clock and SM, with and without an FFMA2 per
// loaded pair (the inner loop of lc_rot_kernel is one LDS.32 per FMA).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_rate lds_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int WIDTH, bool FMA>
__global__ void __launch_bounds__(512) lds_kernel(int iters, float* sink) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) sm[i] = (float)i * 1e-6f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* base = sm + (warp * 64 % 4096) + lane * WIDTH;          // lane-contiguous: conflict free for every width
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float* p = base + ((it + k) & 7) * 1024;
            if (WIDTH == 1) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((uint32_t)__cvta_generic_to_shared(p))); acc[k] = FMA ? fmaf(v, 1.0001f, acc[k]) : acc[k] + v; }
            if (WIDTH == 2) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p))); acc[k] = FMA ? fmaf(v.x, 1.0001f, fmaf(v.y, 0.5f, acc[k])) : acc[k] + v.x + v.y; }
            if (WIDTH == 4) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(p))); acc[k] = FMA ? fmaf(v.x, 1.0001f, fmaf(v.y, 0.5f, fmaf(v.z, 0.25f, fmaf(v.w, 0.125f, acc[k])))) : acc[k] + v.x + v.y + v.z + v.w; }
        }
    }
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += acc[k];
    if (s == 1.2345f) sink[0] = s;
}

template <int WIDTH, bool FMA>
static void run(int sms, int threads, int ctas, float* sink) {
    const int iters = 2000;
    cudaFuncSetAttribute(lds_kernel<WIDTH, FMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    lds_kernel<WIDTH, FMA><<<sms * ctas, threads, 65536>>>(iters, sink);
    cudaEventRecord(e0);
    lds_kernel<WIDTH, FMA><<<sms * ctas, threads, 65536>>>(iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double n_lds = (double)ctas * (threads / 32) * iters * 8;            // warp-level LDS per SM
    const double clk = ms * 1e-3 * 1.965e9;
    printf("LDS.%-3d %s warps/SM %2d: %.3f ms  %.3f LDS/clk/SM  %.1f B/clk/SM\n", 32 * WIDTH, FMA ? "+fma" : "+add", ctas * threads / 32, ms,
           n_lds / clk, n_lds * 128 * WIDTH / clk);
}

int main() {
    float* sink;
    cudaMalloc(&sink, 64);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int threads : {256, 512})
        for (int ctas : {1, 2, 3}) {
            run<1, true>(sms, threads, ctas, sink);
            run<2, true>(sms, threads, ctas, sink);
            run<4, true>(sms, threads, ctas, sink);
        }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
