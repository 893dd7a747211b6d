// Match post-process and balanced-sampling kernels
// (reference: tail of GFNet.match, model/network.py:358-384; GFNet.sample, :385-414).
#include "common.cuh"
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

namespace gfb {

// ---- match post-process ---------------------------------------------------------------------------
// One thread per output lattice cell of warp[Bp, G, Wd, 4] (Wd = 2G symmetric, else G).
__global__ void __launch_bounds__(256) match_post_kernel(const float* __restrict__ flow, const float* __restrict__ logit,
                                                         const float* __restrict__ atten, float* __restrict__ warp,
                                                         float* __restrict__ cert, int b, int G, int symmetric) {
    const int Bp = symmetric ? b / 2 : b;
    const int Wd = symmetric ? 2 * G : G;
    const size_t gg = (size_t)G * G;
    const size_t total = (size_t)Bp * G * Wd;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int xw = (int)(t % Wd);
        const int gy = (int)((t / Wd) % G);
        const int bp = (int)(t / ((size_t)Wd * G));
        const bool second = xw >= G;             // right half: (B->A flow, lattice)
        const int gx = second ? xw - G : xw;
        const int src = second ? bp + Bp : bp;   // model/network.py:213-222: A->B first, then B->A
        const size_t cell = (size_t)gy * G + gx;
        float fx = flow[(size_t)src * 2 * gg + cell], fy = flow[(size_t)src * 2 * gg + gg + cell];
        float lg = logit[(size_t)src * gg + cell];
        if (atten) lg -= atten[(size_t)src * gg + cell];
        float c = 1.f / (1.f + expf(-lg));                       // sigmoid (:360)
        if (fabsf(fx) > 1.f || fabsf(fy) > 1.f) c = 0.f;          // :368-370
        fx = fminf(1.f, fmaxf(-1.f, fx));                         // :371
        fy = fminf(1.f, fmaxf(-1.f, fy));
        // lattice = linspace(-1+1/G, 1-1/G, G) (:362-367)
        const float s0 = (float)(-1.0 + 1.0 / (double)G), s1 = (float)(1.0 - 1.0 / (double)G);  // python doubles -> fp32
        const float step = G > 1 ? (s1 - s0) / (float)(G - 1) : 0.f;
        const float lx = (gx < G / 2) ? s0 + step * (float)gx : s1 - step * (float)(G - 1 - gx);
        const float ly = (gy < G / 2) ? s0 + step * (float)gy : s1 - step * (float)(G - 1 - gy);
        float4 o = second ? make_float4(fx, fy, lx, ly) : make_float4(lx, ly, fx, fy);
        reinterpret_cast<float4*>(warp)[t] = o;
        cert[t] = c;
    }
}

__global__ void __launch_bounds__(256) sample_keys_kernel(const float* __restrict__ cert, const float* __restrict__ noise,
                                                          float* __restrict__ key, long long n, float thresh) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        float c = cert[t];
        c = c > thresh ? 1.f : c;
        key[t] = __fdiv_rn(c, noise[t]);   // ATen: p / q, IEEE division
    }
}

__global__ void __launch_bounds__(256) balance_keys_kernel(const float* __restrict__ rho, const float* __restrict__ noise,
                                                           float* __restrict__ key, long long n, float min_density) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
        const float d = rho[t];
        float p = __fdiv_rn(1.f, d + 1.f);
        if (d < min_density) p = 1e-7f;
        key[t] = __fdiv_rn(p, noise[t]);
    }
}

__global__ void __launch_bounds__(256) gather_matches_kernel(const float* __restrict__ warp, const float* __restrict__ cert,
                                                             const int64_t* __restrict__ idx, float* __restrict__ om,
                                                             float* __restrict__ oc, int B, long long n_src, int n_sel,
                                                             float thresh) {
    const long long total = (long long)B * n_sel;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(t / n_sel);
        const long long s = idx[t];
        reinterpret_cast<float4*>(om)[t] = __ldg(reinterpret_cast<const float4*>(warp) + (size_t)b * n_src + s);
        if (oc) {
            float c = cert[(size_t)b * n_src + s];
            oc[t] = c > thresh ? 1.f : c;
        }
    }
}

// ---- top-k (descending, ties -> lower index first) -------------------------------------------------
// One CTA per row: 3-pass radix select of the k-th largest key, ordered compaction of the k winners
// (index order) into a global scratch row, then a stable block radix sort (descending).
__device__ __forceinline__ uint32_t order_key(float f) {   // monotone float -> uint32
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int TK_THREADS = 512;

template <int IPT>
__global__ void __launch_bounds__(TK_THREADS, 1) topk_kernel(const float* __restrict__ keys, int64_t* __restrict__ idx_out,
                                                             long long n, int k, uint32_t* __restrict__ ws) {
    using Sort = cub::BlockRadixSort<uint32_t, TK_THREADS, IPT, uint32_t>;
    using Scan = cub::BlockScan<int, TK_THREADS>;
    extern __shared__ __align__(16) unsigned char dsm[];
    typename Sort::TempStorage& sort_tmp = *reinterpret_cast<typename Sort::TempStorage*>(dsm);
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ int hist[2048];
    __shared__ uint32_t s_prefix, s_mask;
    __shared__ int s_remaining;

    const int b = blockIdx.x;
    const float* row = keys + (size_t)b * n;
    uint32_t* wk = ws + (size_t)b * 2 * TK_THREADS * IPT;
    uint32_t* wi = wk + TK_THREADS * IPT;
    if (threadIdx.x == 0) { s_prefix = 0; s_mask = 0; s_remaining = k; }

    // ---- radix select, most significant digits first (11 + 11 + 10 bits); k == n (a full sort, e.g. the kde's
    // space-filling-curve ordering) needs neither the select nor the compaction
    const int shifts[3] = {21, 10, 0}, nbits[3] = {11, 11, 10};
    const bool full = (long long)k == n;
    const bool vec4 = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
    for (int pass = 0; pass < (full ? 0 : 3); ++pass) {
        for (int i = threadIdx.x; i < 2048; i += TK_THREADS) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, mask = s_mask;
        const int remaining = s_remaining;
        const int sh = shifts[pass], nb = 1 << nbits[pass];
        if (vec4) {   // 16-byte loads: four keys per thread and iteration in flight
            // four loads in flight per thread: with one CTA per SM the pass is pure load latency otherwise (ncu: the first use
            // of the loaded key was 55 % of this kernel's stall samples)
            const float4* row4 = reinterpret_cast<const float4*>(row);
            const long long n4 = n / 4;
            for (long long i = threadIdx.x; i < n4; i += TK_THREADS * 4) {
                float4 v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (i + q * TK_THREADS < n4) v[q] = __ldg(row4 + i + q * TK_THREADS);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (i + q * TK_THREADS >= n4) break;
                    const uint32_t u[4] = {order_key(v[q].x), order_key(v[q].y), order_key(v[q].z), order_key(v[q].w)};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((u[e] & mask) == prefix) atomicAdd(&hist[(u[e] >> sh) & (nb - 1)], 1);
                }
            }
        } else {
            for (long long i = threadIdx.x; i < n; i += TK_THREADS) {
                const uint32_t u = order_key(row[i]);
                if ((u & mask) == prefix) atomicAdd(&hist[(u >> sh) & (nb - 1)], 1);
            }
        }
        __syncthreads();
        // suffix sums over bins, highest bin first: thread t owns bins nb-1-4t .. nb-4-4t
        int c[4], tot = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) { const int bin = nb - 1 - (threadIdx.x * 4 + e); c[e] = bin >= 0 ? hist[bin] : 0; tot += c[e]; }
        int before;
        Scan(scan_tmp).ExclusiveSum(tot, before);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int bin = nb - 1 - (threadIdx.x * 4 + e);
            if (bin >= 0 && before < remaining && before + c[e] >= remaining) {   // the k-th largest lives here
                s_prefix = prefix | ((uint32_t)bin << sh);
                s_mask = mask | ((uint32_t)(nb - 1) << sh);
                s_remaining = remaining - before;
            }
            before += c[e];
        }
        __syncthreads();
    }
    const uint32_t T = s_prefix;          // order_key of the k-th largest element
    const int need_eq = s_remaining;      // how many elements equal to T are taken (lowest indices)

    // ---- ordered compaction of the winners (index order) into scratch: CK consecutive keys per thread and round (the two
    // block-wide scans and barriers per round are what this phase costs: 16 keys per thread = 4x fewer rounds than 4)
    constexpr int CK = 16;
    int base_sel = 0, base_eq = 0;
    for (long long t0 = 0; t0 < (full ? 0 : n); t0 += (long long)TK_THREADS * CK) {
        const long long i0 = t0 + (long long)threadIdx.x * CK;
        uint32_t u[CK];
        int eq = 0;
        if (vec4 && i0 + CK <= n) {
#pragma unroll
            for (int e = 0; e < CK; e += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(row + i0 + e));
                u[e] = order_key(v.x); u[e + 1] = order_key(v.y); u[e + 2] = order_key(v.z); u[e + 3] = order_key(v.w);
            }
#pragma unroll
            for (int e = 0; e < CK; ++e) eq += (u[e] == T);
        } else {
#pragma unroll
            for (int e = 0; e < CK; ++e) { u[e] = (i0 + e < n) ? order_key(row[i0 + e]) : 0u; eq += (i0 + e < n && u[e] == T); }
        }
        int eq_before, eq_total;
        Scan(scan_tmp).ExclusiveSum(eq, eq_before, eq_total);
        __syncthreads();
        int sel = 0, er = base_eq + eq_before;
        uint32_t take = 0;
#pragma unroll
        for (int e = 0; e < CK; ++e) {
            const bool in = i0 + e < n;
            const bool tk = in && (u[e] > T || (u[e] == T && er < need_eq));
            take |= (uint32_t)tk << e;
            er += (in && u[e] == T);
            sel += tk;
        }
        int sel_before, sel_total;
        Scan(scan_tmp).ExclusiveSum(sel, sel_before, sel_total);
        __syncthreads();
        int pos = base_sel + sel_before;
#pragma unroll
        for (int e = 0; e < CK; ++e)
            if (take >> e & 1) { wk[pos] = u[e]; wi[pos] = (uint32_t)(i0 + e); ++pos; }
        base_sel += sel_total;
        base_eq += eq_total;
    }
    __syncthreads();   // scratch row complete (block-scope visibility of global writes)

    // ---- stable descending sort of the k winners
    uint32_t sk[IPT], sv[IPT];
#pragma unroll
    for (int e = 0; e < IPT; ++e) {
        const int p = threadIdx.x * IPT + e;
        if (full) {
            sk[e] = p < k ? order_key(row[p]) : 0u;
            sv[e] = p < k ? (uint32_t)p : 0u;
        } else {
            sk[e] = p < k ? wk[p] : 0u;   // padding sorts last (every real key has its top bit set or is > 0)
            sv[e] = p < k ? wi[p] : 0u;
        }
    }
    Sort(sort_tmp).SortDescending(sk, sv);
#pragma unroll
    for (int e = 0; e < IPT; ++e) {
        const int p = threadIdx.x * IPT + e;
        if (p < k) idx_out[(size_t)b * k + p] = (int64_t)sv[e];
    }
}

template <int IPT>
static int launch_topk(const float* keys, int64_t* idx, int B, long long n, int k, void* ws, cudaStream_t st) {
    using Sort = cub::BlockRadixSort<uint32_t, TK_THREADS, IPT, uint32_t>;
    const int smem = (int)sizeof(typename Sort::TempStorage);
    cudaError_t e = cudaFuncSetAttribute(topk_kernel<IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    topk_kernel<IPT><<<B, TK_THREADS, smem, st>>>(keys, idx, n, k, reinterpret_cast<uint32_t*>(ws));
    GFB_LAUNCH_RESULT();
}

}  // namespace gfb

using namespace gfb;

static inline int ew_blocks(long long n) { return (int)std::min<long long>(148LL * 16, (n + 255) / 256); }

extern "C" int gfb_match_postprocess_f32(const float* flow, const float* cert_logits, const float* attenuation,
                                         float* warp, float* cert, int b, int G, int symmetric, gfb_stream_t stream) {
    GFB_CHECK_ARG(flow && cert_logits && warp && cert && b > 0 && G > 0);
    GFB_CHECK_ARG(!symmetric || (b % 2 == 0));
    if (!gfb_aligned(warp, 16)) return GFB_EALIGN;
    const long long total = (long long)(symmetric ? b / 2 : b) * G * (symmetric ? 2 * G : G);
    match_post_kernel<<<ew_blocks(total), 256, 0, gfb_cu(stream)>>>(flow, cert_logits, attenuation, warp, cert, b, G, symmetric);
    GFB_LAUNCH_RESULT();
}

extern "C" int gfb_sample_keys_f32(const float* cert, const float* noise, float* key, long long n, float thresh,
                                   gfb_stream_t stream) {
    GFB_CHECK_ARG(cert && noise && key && n > 0);
    sample_keys_kernel<<<ew_blocks(n), 256, 0, gfb_cu(stream)>>>(cert, noise, key, n, thresh);
    GFB_LAUNCH_RESULT();
}

extern "C" int gfb_balance_keys_f32(const float* density, const float* noise, float* key, long long n,
                                    float min_density, gfb_stream_t stream) {
    GFB_CHECK_ARG(density && noise && key && n > 0);
    balance_keys_kernel<<<ew_blocks(n), 256, 0, gfb_cu(stream)>>>(density, noise, key, n, min_density);
    GFB_LAUNCH_RESULT();
}

extern "C" int gfb_gather_matches_f32(const float* warp, const float* cert, const int64_t* idx, float* out_m,
                                      float* out_c, int B, long long n_src, int n_sel, float thresh,
                                      gfb_stream_t stream) {
    GFB_CHECK_ARG(warp && idx && out_m && B > 0 && n_src > 0 && n_sel > 0);
    GFB_CHECK_ARG(!out_c || cert);
    if (!gfb_aligned(warp, 16) || !gfb_aligned(out_m, 16)) return GFB_EALIGN;
    gather_matches_kernel<<<ew_blocks((long long)B * n_sel), 256, 0, gfb_cu(stream)>>>(warp, cert, idx, out_m, out_c, B, n_src, n_sel, thresh);
    GFB_LAUNCH_RESULT();
}

extern "C" size_t gfb_topk_workspace_bytes(int B, long long n, int k) {
    (void)n;
    if (B <= 0 || k <= 0) return 0;
    const int cap = k <= TK_THREADS * 10 ? TK_THREADS * 10 : TK_THREADS * 40;
    return (size_t)B * 2 * cap * sizeof(uint32_t);
}

extern "C" int gfb_topk_desc_f32(const float* keys, int64_t* idx_out, int B, long long n, int k,
                                 void* workspace, size_t workspace_bytes, gfb_stream_t stream) {
    GFB_CHECK_ARG(keys && idx_out && B > 0 && n > 0 && k > 0 && k <= n && n < (1LL << 31));
    if (k > TK_THREADS * 40) return GFB_EUNSUPPORTED;
    if (!workspace || workspace_bytes < gfb_topk_workspace_bytes(B, n, k)) return GFB_EWORKSPACE;
    if (k <= TK_THREADS * 10) return launch_topk<10>(keys, idx_out, B, n, k, workspace, gfb_cu(stream));
    return launch_topk<40>(keys, idx_out, B, n, k, workspace, gfb_cu(stream));
}
