// Warp-level MMA (mma.sync) and TMEM read-out rates on sm_100a: the numbers the local-correlation kernel design needs.
//   hmma     m16n8k16 bf16 -> f32, register operands only, NI independent accumulators per warp
//   hmma_ldm the same with one ldmatrix.x4 (512 B, conflict-free) per 3 MMAs
//   tf32     m16n8k8 tf32 -> f32
//   ldtm     tcgen05.ld.32x32b.x32 / .x16 by 4 warps per CTA, 1 or 2 CTAs per SM
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rates mma_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <int MODE, int NI>
__global__ void mma_kernel(int iters, float* sink) {
    __shared__ __align__(128) unsigned char sm[16384];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u + i;
    __syncthreads();
    float d[NI][4];
    uint32_t a[4], b[NI][2];
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < 4; ++i) a[i] = 0x3f803f80u + lane + i;
    for (int n = 0; n < NI; ++n) { b[n][0] = 0x3f003f00u + n; b[n][1] = 0x3e803e80u + lane; for (int i = 0; i < 4; ++i) d[n][i] = 0.f; }
    // conflict-free ldmatrix addresses: lane l -> row l (16 B each) of a 512 B block
    const uint32_t base = smem_u32(sm) + (threadIdx.x >> 5) * 512 % 8192 + lane * 16;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 1) {
#pragma unroll
            for (int n = 0; n < NI; n += 2) {           // one ldmatrix.x4 -> two B fragments (hi) ; pretend lo = same regs shifted
                uint32_t r[4];
                ldsm4(r, base + ((it + n) & 7) * 1024);
                b[n][0] = r[0]; b[n][1] = r[1]; b[n + 1][0] = r[2]; b[n + 1][1] = r[3];
            }
        }
#pragma unroll
        for (int rep = 0; rep < 3; ++rep)
#pragma unroll
            for (int n = 0; n < NI; ++n) {
                if (MODE == 2) mma_tf32(d[n], a, b[n]);
                else hmma(d[n], a, b[n]);
            }
    }
    float s = 0.f;
    for (int n = 0; n < NI; ++n) for (int i = 0; i < 4; ++i) s += d[n][i];
    if (s == 123.456f) sink[0] = s;
}

// ---- TMEM read-out -------------------------------------------------------------------------------------------------
template <int X>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&r)[X]);
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}

template <int X>
__global__ void __launch_bounds__(128) ldtm_kernel(int iters, int ncols, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot + ((uint32_t)(warp * 32) << 16);
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        uint32_t r0[X], r1[X];
        tmem_ld<X>(tb + (uint32_t)((it * 2 * X) % (ncols - 2 * X + 1)), r0);
        tmem_ld<X>(tb + (uint32_t)((it * 2 * X + X) % (ncols - X + 1)), r1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc ^= r0[0] ^ r1[X - 1] ^ r0[X / 2];
    }
    if (acc == 0x12345u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(ncols) : "memory");
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float* sink;
    cudaMalloc(&sink, 64);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4000;
    const char* names[3] = {"hmma bf16 m16n8k16 (regs)", "hmma bf16 + ldmatrix.x4 per 6 mma", "mma tf32 m16n8k8 (regs)"};
    for (int mode = 0; mode < 3; ++mode)
        for (int threads : {128, 256, 512, 1024}) {
            float ms = 0.f;
            constexpr int NI = 8;
            if (mode == 0) ms = time_ms([&] { mma_kernel<0, NI><<<sms, threads>>>(iters, sink); });
            if (mode == 1) ms = time_ms([&] { mma_kernel<1, NI><<<sms, threads>>>(iters, sink); });
            if (mode == 2) ms = time_ms([&] { mma_kernel<2, NI><<<sms, threads>>>(iters, sink); });
            const double n_mma = (double)sms * (threads / 32) * iters * 3 * NI;
            const double flop = n_mma * (mode == 2 ? 2048.0 : 4096.0);
            printf("%-36s warps/SM %2d: %.3f ms  %.1f TFLOP/s  %.2f clk/MMA/SMSP @1.965GHz\n", names[mode], threads / 32, ms,
                   flop / ms * 1e-9, ms * 1e-3 * 1.965e9 / (n_mma / (sms * 4.0)));
        }
    for (int ctas : {1, 2})
        for (int x : {32, 16}) {
            const int ncols = 512 / ctas;
            float ms = 0.f;
            if (x == 32) ms = time_ms([&] { ldtm_kernel<32><<<sms * ctas, 128>>>(iters, ncols, reinterpret_cast<uint32_t*>(sink)); });
            else ms = time_ms([&] { ldtm_kernel<16><<<sms * ctas, 128>>>(iters, ncols, reinterpret_cast<uint32_t*>(sink)); });
            const double bytes = (double)sms * ctas * 4 * iters * 2 * x * 128;
            printf("tcgen05.ld.32x32b.x%d, %d CTA/SM x 4 warps: %.3f ms  %.1f TB/s chip  %.1f B/clk/SM\n", x, ctas, ms,
                   bytes / ms * 1e-9, bytes / sms / (ms * 1e-3 * 1.965e9));
        }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
