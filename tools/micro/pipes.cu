// Microbenchmarks that size the local_correlation inner loop on B200:
//   FFMA vs FFMA2 (fma.rn.f32x2) issue/pipe throughput, LDS.128 / LDS.64 / LDS.32 wavefront cost for the
//   lane-address patterns the kernel produces (lane stride s pixels, 16B-aligned segments).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, float a, float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    float2 a2 = make_float2(a, a);
    unsigned long long aa = *reinterpret_cast<unsigned long long*>(&a2);
    unsigned long long bb = *reinterpret_cast<unsigned long long*>(&b);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}

template <int MODE>
__global__ void fma_kernel(float* out, int iters, float a) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    float2 b = make_float2(1.0001f, 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) {
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].x) : "f"(a), "f"(b.x));
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i].y) : "f"(a), "f"(b.y));
            } else {
                ffma2(acc[i], a, b);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// Shared-memory load cost by lane-address pattern.  VEC = floats per lane per load (1, 2, 4).
//   pat 0: all lanes the same address            pat 1: lane l -> floor(l*1.00) floats, aligned down to VEC
//   pat 2: floor(l*1.75)                         pat 3: lane l -> chunk (l % 8) of 16 B (8 distinct chunks = 128 B)
//   pat 4: lane l -> chunk (l / 4) (8 distinct)  pat 5: every lane its own 16 B chunk (512 B contiguous)
//   pat 6: two rows: (l%16)*1.75 floats + (l/16)*264 floats (channel split, 16 cols x 2 channel rows)
template <int VEC>
__global__ void lds_kernel(float* out, int iters, int pat) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int l = threadIdx.x & 31;
    int base;
    switch (pat) {
        case 0: base = 0; break;
        case 1: base = l & ~(VEC - 1); break;
        case 2: base = ((int)(l * 1.75f)) & ~(VEC - 1); break;
        case 3: base = (l % 8) * 4; break;
        case 4: base = (l / 4) * 4; break;
        case 5: base = l * 4; break;
        default: base = (((int)((l % 16) * 1.75f)) & ~(VEC - 1)) + (l / 16) * 264; break;
    }
    base += (threadIdx.x >> 5) * 32;
    unsigned addr = (unsigned)__cvta_generic_to_shared(sm + base);
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int it = 0; it < iters; ++it) {
        float v[8][4];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned ad = addr + k * 512 + (it & 1) * 4096 * 4;
            if (VEC == 4) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[k][0]), "=f"(v[k][1]), "=f"(v[k][2]), "=f"(v[k][3]) : "r"(ad));
            else if (VEC == 2) { asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v[k][0]), "=f"(v[k][1]) : "r"(ad)); v[k][2] = v[k][3] = 0.f; }
            else { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[k][0]) : "r"(ad)); v[k][1] = v[k][2] = v[k][3] = 0.f; }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += (VEC == 4) ? (v[k][0] + v[k][1]) + (v[k][2] + v[k][3]) : (VEC == 2 ? v[k][0] + v[k][1] : v[k][0]);
    }
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("clock attr %d kHz\n", clk_khz);
    const int iters = 20000;
    for (int warps = 4; warps <= 16; warps *= 2) {
        float t0 = time_ms([&] { fma_kernel<0><<<148, warps * 32>>>(out, iters, 1.00001f); });
        float t1 = time_ms([&] { fma_kernel<1><<<148, warps * 32>>>(out, iters, 1.00001f); });
        double fmas = 148.0 * warps * 32 * iters * 32.0;
        printf("warps/SM %2d: FFMA %.1f GFMA/s (%.1f /clk/SM @1.965)  FFMA2 %.1f GFMA/s (%.1f /clk/SM)\n", warps,
               fmas / t0 / 1e6, fmas / t0 / 1e6 / 148 / 1.965, fmas / t1 / 1e6, fmas / t1 / 1e6 / 148 / 1.965);
    }
    for (int pat = 0; pat <= 6; ++pat) {
        const int it2 = 4000, warps = 8;
        float t4 = time_ms([&] { lds_kernel<4><<<148, warps * 32, 40960>>>(out, it2, pat); });
        float t2 = time_ms([&] { lds_kernel<2><<<148, warps * 32, 40960>>>(out, it2, pat); });
        float t1 = time_ms([&] { lds_kernel<1><<<148, warps * 32, 40960>>>(out, it2, pat); });
        double n = (double)warps * it2 * 8;  // warp-level loads per SM
        printf("pattern %d: clk per warp-load  LDS.128 %.2f  LDS.64 %.2f  LDS.32 %.2f\n", pat,
               t4 * 1e-3 * 1.965e9 / n, t2 * 1e-3 * 1.965e9 / n, t1 * 1e-3 * 1.965e9 / n);
    }
    return 0;
}
