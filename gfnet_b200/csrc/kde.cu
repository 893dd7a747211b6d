// K3 -- Gaussian kernel density (reference: utils/kde.py:4-13, call site model/network.py:406-408).
//
// density[b,m] = sum_{m'} exp(-||x[b,m] - y[b,m']||^2 / (2 std^2)),  y = x[::down].
// The reference materialises the [M,M'] distance matrix (cdist -> **2 -> exp -> sum: five passes over
// 0.8-1.6 GB); here nothing is materialised.  Coordinates are pre-scaled by sqrt(log2(e)/(2 std^2)) so
// each evaluation is  ex2(2<x',y'> - |x'|^2 - |y'|^2): 1 FADD + D FFMA + 1 MUFU.EX2 + 1 FADD.  The
// bound is the MUFU pipe (16 ex2/clk/SM), but with scalar arithmetic the kernel ran out of issue slots first
// (7.5 instructions per evaluation, 62 % of the issue rate): rows are therefore processed in pairs with packed
// add.f32x2 / fma.rn.f32x2 (4.5 instructions per evaluation).  y tiles are staged in shared memory and read as
// warp-wide broadcasts, each thread keeps RT rows of x in registers.
#include "common.cuh"

namespace gfb {

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ unsigned long long kpack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long kadd2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long kfma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

constexpr int KDE_THREADS = 128;
constexpr int KDE_TILE = 1024;  // y points per shared-memory tile

template <int RT>
__global__ void __launch_bounds__(KDE_THREADS) kde4_kernel(const float* __restrict__ x, float* __restrict__ density,
                                                           int M, int Mp, int down, float scale) {
    __shared__ float4 sy[KDE_TILE];   // 2 * y'
    __shared__ float sn[KDE_TILE];    // |y'|^2
    const int b = blockIdx.y;
    const float4* xb = reinterpret_cast<const float4*>(x) + (size_t)b * M;
    float4 xr[RT];
    float nneg[RT], acc[RT];
    int row[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        row[r] = (blockIdx.x * RT + r) * KDE_THREADS + threadIdx.x;
        float4 v = row[r] < M ? __ldg(xb + row[r]) : make_float4(0.f, 0.f, 0.f, 0.f);
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        xr[r] = v;
        nneg[r] = -(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w);
        acc[r] = 0.f;
    }
    for (int t0 = 0; t0 < Mp; t0 += KDE_TILE) {
        const int nt = min(KDE_TILE, Mp - t0);
        __syncthreads();
        for (int i = threadIdx.x; i < KDE_TILE; i += KDE_THREADS) {
            if (i < nt) {
                float4 v = __ldg(xb + (size_t)(t0 + i) * down);
                v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
                sn[i] = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
                sy[i] = make_float4(2.f * v.x, 2.f * v.y, 2.f * v.z, 2.f * v.w);
            } else {  // pad: contributes ex2(-huge) = 0
                sn[i] = 3.0e38f;
                sy[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncthreads();
        const int ntp = (nt + 3) & ~3;
        if constexpr (RT % 2 == 0) {
            // rows in pairs: (x_d[r], x_d[r+1]) * (y_d, y_d) on the packed FP32 pipe
            unsigned long long x2[RT / 2][4], nn2[RT / 2];
#pragma unroll
            for (int q = 0; q < RT / 2; ++q) {
                x2[q][0] = kpack2(xr[2 * q].x, xr[2 * q + 1].x);
                x2[q][1] = kpack2(xr[2 * q].y, xr[2 * q + 1].y);
                x2[q][2] = kpack2(xr[2 * q].z, xr[2 * q + 1].z);
                x2[q][3] = kpack2(xr[2 * q].w, xr[2 * q + 1].w);
                nn2[q] = kpack2(nneg[2 * q], nneg[2 * q + 1]);
            }
#pragma unroll 4
            for (int j = 0; j < ntp; ++j) {
                const float4 y = sy[j];
                const float ny = -sn[j];
                const unsigned long long y0 = kpack2(y.x, y.x), y1 = kpack2(y.y, y.y), y2 = kpack2(y.z, y.z), y3 = kpack2(y.w, y.w);
                const unsigned long long nyy = kpack2(ny, ny);
#pragma unroll
                for (int q = 0; q < RT / 2; ++q) {
                    unsigned long long t = kadd2(nn2[q], nyy);
                    t = kfma2(x2[q][0], y0, t);
                    t = kfma2(x2[q][1], y1, t);
                    t = kfma2(x2[q][2], y2, t);
                    t = kfma2(x2[q][3], y3, t);
                    float t0, t1;
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
                    acc[2 * q] += ex2_approx(t0);
                    acc[2 * q + 1] += ex2_approx(t1);
                }
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < ntp; ++j) {
                const float4 y = sy[j];
                const float ny = sn[j];
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    float t = nneg[r] - ny;
                    t = fmaf(xr[r].x, y.x, t);
                    t = fmaf(xr[r].y, y.y, t);
                    t = fmaf(xr[r].z, y.z, t);
                    t = fmaf(xr[r].w, y.w, t);
                    acc[r] += ex2_approx(t);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RT; ++r)
        if (row[r] < M) density[(size_t)b * M + row[r]] = acc[r];
}

// Any D <= 8 (not on the GFNet path; kept so the op is a full drop-in for kde(x) with other widths).
__global__ void __launch_bounds__(KDE_THREADS) kde_generic_kernel(const float* __restrict__ x, float* __restrict__ density,
                                                                  int M, int Mp, int D, int down, float scale) {
    const int b = blockIdx.y;
    const int row = blockIdx.x * KDE_THREADS + threadIdx.x;
    const float* xb = x + (size_t)b * M * D;
    float xr[8];
    for (int d = 0; d < 8; ++d) xr[d] = (row < M && d < D) ? xb[(size_t)row * D + d] * scale : 0.f;
    float acc = 0.f;
    for (int j = 0; j < Mp; ++j) {
        const float* yp = xb + (size_t)j * down * D;
        float d2 = 0.f;
        for (int d = 0; d < D; ++d) { float df = xr[d] - __ldg(yp + d) * scale; d2 = fmaf(df, df, d2); }
        acc += ex2_approx(-d2);
    }
    if (row < M) density[(size_t)b * M + row] = acc;
}

}  // namespace gfb

using namespace gfb;

extern "C" int gfb_kde_f32(const float* x, float* density, int B, int M, int D, int down, float std,
                           gfb_stream_t stream) {
    GFB_CHECK_ARG(x && density && B > 0 && M > 0 && D > 0 && D <= 8 && down >= 1 && std > 0.f);
    GFB_CHECK_ARG(B <= 65535);
    const int Mp = (M + down - 1) / down;
    // exp(-d2/(2 std^2)) = 2^(-d2 * log2(e)/(2 std^2)); scale coordinates by the square root
    const float scale = (float)sqrt(1.4426950408889634 / (2.0 * (double)std * (double)std));
    cudaStream_t st = gfb_cu(stream);
    if (D == 4) {
        if (!gfb_aligned(x, 16)) return GFB_EALIGN;
        int sms = 148;
        const long long rows128 = ((long long)M + KDE_THREADS - 1) / KDE_THREADS;
        // more rows per thread amortise the shared-memory reads, but keep >= ~2 waves of CTAs
        if (rows128 * B >= 8LL * 2 * sms) {
            dim3 grid((unsigned)((rows128 + 3) / 4), B);
            kde4_kernel<4><<<grid, KDE_THREADS, 0, st>>>(x, density, M, Mp, down, scale);
        } else if (rows128 * B >= 4LL * sms) {
            dim3 grid((unsigned)((rows128 + 1) / 2), B);
            kde4_kernel<2><<<grid, KDE_THREADS, 0, st>>>(x, density, M, Mp, down, scale);
        } else {
            dim3 grid((unsigned)rows128, B);
            kde4_kernel<1><<<grid, KDE_THREADS, 0, st>>>(x, density, M, Mp, down, scale);
        }
    } else {
        dim3 grid((unsigned)((M + KDE_THREADS - 1) / KDE_THREADS), B);
        kde_generic_kernel<<<grid, KDE_THREADS, 0, st>>>(x, density, M, Mp, D, down, scale);
    }
    GFB_LAUNCH_RESULT();
}
