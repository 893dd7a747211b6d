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

// ---- symmetric variant (down == 1: y = x) ---------------------------------------------------------------------
// K(x_i, x_j) = K(x_j, x_i), so only the tile pairs (I, J >= I) of 128 x 128 points are evaluated and every value is
// added to the density of its row AND of its column: half the ex2 evaluations.  A CTA owns row block I and up to
// KS_JC column blocks; a thread keeps an 8 x 8 register tile (rows ti*8.., columns tj*8..).  Row sums stay in
// registers until the CTA ends; column sums are folded over the 16 lanes that share the columns with a shuffle
// transpose-reduce after every tile pair.  Partial sums meet in a 64-bit fixed-point accumulator (2^-40 units):
// integer atomics are order-independent, so the density is deterministic although CTAs finish in any order.
constexpr int KS_T = 128;
constexpr int KS_JC = 32;
constexpr int KS_PITCH = 16 * 9;
constexpr float KS_FIX = 1099511627776.f;   // 2^40

__device__ __forceinline__ unsigned long long kde_fix(float v) { return __float2ull_rn(v * KS_FIX); }

// Optional cut-off (cut2 > 0): the points arrive sorted along a Hilbert curve, bb holds the bounding box of every
// 128-point block, and a tile pair whose boxes are further apart than the cut-off radius is skipped -- each of its
// terms is below exp(-cut_sigmas^2 / 2) (2.3e-11 at 7 sigma, far below one fp32 ulp of a density that is >= 1).
__device__ __forceinline__ float bb_dist2(const float* __restrict__ a, const float* __restrict__ b) {
    float d2 = 0.f;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const float g = fmaxf(0.f, fmaxf(a[d] - b[4 + d], b[d] - a[4 + d]));     // [0..3] = lo, [4..7] = hi
        d2 = fmaf(g, g, d2);
    }
    return d2;
}

__global__ void __launch_bounds__(256, 3) kde4_sym_kernel(const float* __restrict__ x, unsigned long long* __restrict__ acc64,
                                                          int M, int nb, float scale, const float* __restrict__ bb, float cut2) {
    // column block, 8 points per tj at a pitch of 9: the two half-warps of a warp (tj, tj + 1) read different banks
    __shared__ float4 sy[2][KS_PITCH];          // 2 * y'
    __shared__ float sn[2][KS_PITCH];           // -|y'|^2
    __shared__ float srow[16][KS_T];            // end of the CTA: row sums of the 16 column groups
    __shared__ int jl[KS_JC + 1];               // column blocks of this CTA that survive the cut-off, then their count
    const int I = blockIdx.x, b = blockIdx.z;
    const int jfirst = I + blockIdx.y * KS_JC;
    if (jfirst >= nb) return;
    const int tid = threadIdx.x, ti = tid & 15, tj = tid >> 4;
    if (tid < 32) {
        int cnt = 0;
        for (int base = 0; base < KS_JC; base += 32) {
            const int J = jfirst + base + tid;
            bool keep = base + tid < KS_JC && J < nb;
            if (keep && cut2 > 0.f && J != I)
                keep = bb_dist2(bb + ((size_t)b * nb + I) * 8, bb + ((size_t)b * nb + J) * 8) <= cut2;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) jl[cnt + __popc(m & ((1u << tid) - 1u))] = J;
            cnt += __popc(m);
        }
        if (tid == 0) jl[KS_JC] = cnt;
    }
    __syncthreads();
    const int nj = jl[KS_JC];
    if (nj == 0) return;
    const int j0 = jl[0];
    const float4* xb = reinterpret_cast<const float4*>(x) + (size_t)b * M;

    unsigned long long x2[4][4], nn2[4], rs2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 v[2];
        float nn[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int row = I * KS_T + ti * 8 + 2 * q + h;
            float4 w = row < M ? __ldg(xb + row) : make_float4(0.f, 0.f, 0.f, 0.f);
            w.x *= scale; w.y *= scale; w.z *= scale; w.w *= scale;
            v[h] = w;
            nn[h] = row < M ? -(w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w) : -3.0e38f;   // masked row: ex2(-inf) = 0
        }
        x2[q][0] = kpack2(v[0].x, v[1].x);
        x2[q][1] = kpack2(v[0].y, v[1].y);
        x2[q][2] = kpack2(v[0].z, v[1].z);
        x2[q][3] = kpack2(v[0].w, v[1].w);
        nn2[q] = kpack2(nn[0], nn[1]);
        rs2[q] = 0ull;
    }
    auto put = [&](int buf, int i, int idx, float4 v) {
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        const int k = (i >> 3) * 9 + (i & 7);
        sn[buf][k] = idx < M ? -(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w) : -3.0e38f;
        sy[buf][k] = make_float4(2.f * v.x, 2.f * v.y, 2.f * v.z, 2.f * v.w);
    };
    if (tid < KS_T) {
        const int idx = j0 * KS_T + tid;
        put(0, tid, idx, idx < M ? __ldg(xb + idx) : make_float4(0.f, 0.f, 0.f, 0.f));
    }
    __syncthreads();
    for (int k = 0; k < nj; ++k) {
        const int J = jl[k];
        const int buf = k & 1;
        // the upper half of the CTA fetches the next column block while everybody computes on this one
        const bool pf = tid >= KS_T && k + 1 < nj;
        const int pidx = (pf ? jl[k + 1] : 0) * KS_T + tid - KS_T;
        float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pf && pidx < M) pv = __ldg(xb + pidx);
        unsigned long long cs2[8];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
            const float4 y = sy[buf][tj * 9 + cc];
            const float ny = sn[buf][tj * 9 + cc];
            const unsigned long long y0 = kpack2(y.x, y.x), y1 = kpack2(y.y, y.y), y2 = kpack2(y.z, y.z), y3 = kpack2(y.w, y.w);
            const unsigned long long nyy = kpack2(ny, ny);
            unsigned long long csum = 0ull;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                unsigned long long t = kadd2(nn2[q], nyy);
                t = kfma2(x2[q][0], y0, t);
                t = kfma2(x2[q][1], y1, t);
                t = kfma2(x2[q][2], y2, t);
                t = kfma2(x2[q][3], y3, t);
                float t0, t1;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
                const unsigned long long e = kpack2(ex2_approx(t0), ex2_approx(t1));
                rs2[q] = kadd2(rs2[q], e);
                csum = kadd2(csum, e);
            }
            cs2[cc] = csum;
        }
        if (J != I) {                            // the diagonal tile is a full square: its row sums already hold everything
            float cs[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                float lo, hi;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(cs2[cc]));
                cs[cc] = lo + hi;
            }
            // fold over the 16 lanes (ti) that share these 8 columns: transpose-reduce, 8 -> 4 -> 2 -> 1 values per lane
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float send = (ti & 1) ? cs[k] : cs[k + 4], keep = (ti & 1) ? cs[k + 4] : cs[k];
                cs[k] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float send = (ti & 2) ? cs[k] : cs[k + 2], keep = (ti & 2) ? cs[k + 2] : cs[k];
                cs[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            {
                const float send = (ti & 4) ? cs[0] : cs[1], keep = (ti & 4) ? cs[1] : cs[0];
                cs[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            cs[0] += __shfl_xor_sync(0xffffffffu, cs[0], 8);
            const int col = 4 * (ti & 1) + 2 * ((ti >> 1) & 1) + ((ti >> 2) & 1);
            const int idx = J * KS_T + tj * 8 + col;
            if (ti < 8 && idx < M) atomicAdd(acc64 + (size_t)b * M + idx, kde_fix(cs[0]));
        }
        if (pf) put(buf ^ 1, tid - KS_T, pidx, pv);
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(rs2[q]));
        srow[tj][ti * 8 + 2 * q] = lo;
        srow[tj][ti * 8 + 2 * q + 1] = hi;
    }
    __syncthreads();
    if (tid < KS_T) {
        const int idx = I * KS_T + tid;
        float sum = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) sum += srow[g][tid];          // fixed order: deterministic
        if (idx < M) atomicAdd(acc64 + (size_t)b * M + idx, kde_fix(sum));
    }
}

// density[b, perm[i]] = accumulated fixed-point sum of sorted point i (perm == nullptr: identity)
__global__ void __launch_bounds__(256) kde_finish_kernel(const unsigned long long* __restrict__ acc64, const long long* __restrict__ perm,
                                                         float* __restrict__ density, int M, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = (float)((double)acc64[i] * (1.0 / 1099511627776.0));
        density[perm ? (i / M) * M + (size_t)perm[i] : i] = v;
    }
}

// Hilbert-curve index (10 + 10 bits) of the first two coordinates as an exactly representable float sort key.  Unlike
// the Morton curve the Hilbert curve never jumps, so ANY 128 consecutive points are spatially compact (tight boxes).
__global__ void __launch_bounds__(256) kde_keys_kernel(const float4* __restrict__ x, float* __restrict__ keys, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        uint32_t qx = (uint32_t)fminf(fmaxf((v.x + 1.f) * 511.5f, 0.f), 1023.f);
        uint32_t qy = (uint32_t)fminf(fmaxf((v.y + 1.f) * 511.5f, 0.f), 1023.f);
        uint32_t d = 0;
#pragma unroll
        for (uint32_t sbit = 512; sbit > 0; sbit >>= 1) {
            const uint32_t rx = (qx & sbit) ? 1u : 0u, ry = (qy & sbit) ? 1u : 0u;
            d += sbit * sbit * ((3u * rx) ^ ry);
            if (ry == 0) {
                if (rx == 1) { qx = 1023u - qx; qy = 1023u - qy; }
                const uint32_t t = qx; qx = qy; qy = t;
            }
        }
        keys[i] = (float)d;
    }
}

// xs[b, i] = x[b, perm[b, i]] and the bounding box of every block of 128 sorted points (one CTA per block)
__global__ void __launch_bounds__(KS_T) kde_gather_bbox_kernel(const float4* __restrict__ x, const long long* __restrict__ perm,
                                                               float4* __restrict__ xs, float* __restrict__ bb, int M, int nb) {
    __shared__ float sm[4][8];
    const int blk = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const int i = blk * KS_T + t;
    float lo[4], hi[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) { lo[d] = 3.0e38f; hi[d] = -3.0e38f; }
    if (i < M) {
        const float4 v = __ldg(x + (size_t)b * M + (size_t)perm[(size_t)b * M + i]);
        xs[(size_t)b * M + i] = v;
        lo[0] = hi[0] = v.x; lo[1] = hi[1] = v.y; lo[2] = hi[2] = v.z; lo[3] = hi[3] = v.w;
    }
#pragma unroll
    for (int d = 0; d < 4; ++d)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    if ((t & 31) == 0)
#pragma unroll
        for (int d = 0; d < 4; ++d) { sm[t >> 5][d] = lo[d]; sm[t >> 5][4 + d] = hi[d]; }
    __syncthreads();
    if (t < 8) {
        float v = sm[0][t];
        for (int w = 1; w < 4; ++w) v = t < 4 ? fminf(v, sm[w][t]) : fmaxf(v, sm[w][t]);
        bb[((size_t)b * nb + blk) * 8 + t] = v;
    }
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

extern "C" size_t gfb_topk_workspace_bytes(int B, long long n, int k);
extern "C" int gfb_topk_desc_f32(const float* keys, int64_t* idx_out, int B, long long n, int k,
                                 void* workspace, size_t workspace_bytes, gfb_stream_t stream);

static inline size_t kde_align(size_t v) { return (v + 255) / 256 * 256; }

extern "C" size_t gfb_kde_sym_workspace_bytes(int B, int M) {
    if (B <= 0 || M <= 0) return 0;
    const size_t n = (size_t)B * M, nb = (size_t)(M + KS_T - 1) / KS_T;
    // fixed-point sums | sort keys | permutation | sorted points | block boxes | top-k scratch
    return kde_align(n * 8) + kde_align(n * 4) + kde_align(n * 8) + kde_align(n * 16) + kde_align((size_t)B * nb * 32) +
           kde_align(gfb_topk_workspace_bytes(B, M, M));
}

// kde(x) with y = x (down == 1), D = 4: symmetric evaluation, see kde4_sym_kernel.  cut_sigmas > 0 additionally sorts
// the points along a Hilbert curve and skips tile pairs further apart than cut_sigmas * std (0 = evaluate every pair).
extern "C" int gfb_kde_sym_f32(const float* x, float* density, int B, int M, float std, float cut_sigmas,
                               void* workspace, size_t workspace_bytes, gfb_stream_t stream) {
    GFB_CHECK_ARG(x && density && B > 0 && M > 0 && std > 0.f && cut_sigmas >= 0.f);
    GFB_CHECK_ARG(B <= 65535);
    if (!gfb_aligned(x, 16) || !gfb_aligned(workspace, 256)) return GFB_EALIGN;
    if (!workspace || workspace_bytes < gfb_kde_sym_workspace_bytes(B, M)) return GFB_EWORKSPACE;
    const float scale = (float)sqrt(1.4426950408889634 / (2.0 * (double)std * (double)std));
    cudaStream_t st = gfb_cu(stream);
    const int nb = (M + KS_T - 1) / KS_T;
    if (nb > 65535) return GFB_EUNSUPPORTED;
    const size_t n = (size_t)B * M;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    unsigned long long* acc = reinterpret_cast<unsigned long long*>(w); w += kde_align(n * 8);
    float* keys = reinterpret_cast<float*>(w); w += kde_align(n * 4);
    long long* perm = reinterpret_cast<long long*>(w); w += kde_align(n * 8);
    float4* xs = reinterpret_cast<float4*>(w); w += kde_align(n * 16);
    float* bb = reinterpret_cast<float*>(w); w += kde_align((size_t)B * nb * 32);
    cudaError_t e = cudaMemsetAsync(acc, 0, n * 8, st);
    if (e != cudaSuccess) return (int)e;
    // the sort reuses the top-k kernel (k = n): it orders at most 20480 keys per row
    const bool cut = cut_sigmas > 0.f && gfb_topk_workspace_bytes(B, M, M) > 0 && M <= 20480 && nb > 16;
    const unsigned gs = (unsigned)min((size_t)148 * 8, (n + 255) / 256);
    if (cut) {
        kde_keys_kernel<<<gs, 256, 0, st>>>(reinterpret_cast<const float4*>(x), keys, n);
        int rc = gfb_topk_desc_f32(keys, reinterpret_cast<int64_t*>(perm), B, M, M, w, gfb_topk_workspace_bytes(B, M, M), stream);
        if (rc != GFB_OK) return rc;
        kde_gather_bbox_kernel<<<dim3((unsigned)nb, (unsigned)B), KS_T, 0, st>>>(reinterpret_cast<const float4*>(x), perm, xs, bb, M, nb);
    }
    dim3 grid((unsigned)nb, (unsigned)((nb + KS_JC - 1) / KS_JC), (unsigned)B);
    const float r = cut_sigmas * std;
    kde4_sym_kernel<<<grid, 256, 0, st>>>(cut ? reinterpret_cast<const float*>(xs) : x, acc, M, nb, scale, bb, cut ? r * r : 0.f);
    kde_finish_kernel<<<gs, 256, 0, st>>>(acc, cut ? perm : nullptr, density, M, n);
    GFB_LAUNCH_RESULT();
}

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
