// K2 -- coarse global match (reference: GFNet.corr_volume + GFNet.pos_embed, model/network.py:415-440).
//
//   vol[b,j,i] = <f0[b,:,i], f1[b,:,j]> / sqrt(C);  P = softmax_j(vol);  flow[b,:,i] = sum_j P[j,i] grid[j]
//
// It is attention with d_v = 2.  The reference writes the [B,N1,N0] volume, re-reads it for a strided
// softmax and again for a skinny GEMM; here a CTA owns (batch element, 128 source positions i) and
// sweeps the N1 target positions in tiles of 128:
//   * operands are read once from NCHW global memory, split into TF32 hi/lo parts and stored K-major
//     with the 128B swizzle the tensor core expects,
//   * one thread issues tcgen05.mma (kind::tf32, M=128, N=128, K=8) into a double-buffered TMEM
//     accumulator; the 3xTF32 split (hi*hi + hi*lo + lo*hi) keeps fp32 accuracy, the reference
//     computes this einsum in fp32 (TF32 is off at inference),
//   * four epilogue warps pull their 32 TMEM lanes with tcgen05.ld and run an online softmax fused
//     with the grid expectation, so the volume is never written unless the caller asks for it.
// gm_simt_kernel is the any-shape fp32 CUDA-core version (fallback for C > 64 and cross-check).
#include "common.cuh"
#include <math.h>

namespace gfb {

__device__ __forceinline__ float gm_linspace(int n, int i) {  // torch.linspace(-1+1/n, 1-1/n, n)[i]
    float start = (float)(-1.0 + 1.0 / (double)n), end = (float)(1.0 - 1.0 / (double)n);  // python doubles -> fp32
    if (n <= 1) return start;
    float step = (end - start) / (float)(n - 1);
    return (i < n / 2) ? start + step * (float)i : end - step * (float)(n - 1 - i);
}

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ------------------------------------------------------------------------------------------------
// SIMT version: one thread per source position i, f1 tile broadcast from shared memory.
constexpr int GM_SIMT_T = 128;
constexpr int GM_SIMT_TJ = 32;
__global__ void __launch_bounds__(GM_SIMT_T) gm_simt_kernel(const float* __restrict__ f0, const float* __restrict__ f1,
                                                            float* __restrict__ flow, float* __restrict__ vol,
                                                            int C, int N0, int N1, int H1, int W1) {
    extern __shared__ float sm[];
    float* s0 = sm;                            // [C][GM_SIMT_T]  f0 tile, thread-contiguous
    float* s1 = sm + (size_t)C * GM_SIMT_T;    // [GM_SIMT_TJ][C] f1 tile, channel-contiguous
    const int b = blockIdx.y;
    const int i = blockIdx.x * GM_SIMT_T + threadIdx.x;
    const float* f0b = f0 + (size_t)b * C * N0;
    const float* f1b = f1 + (size_t)b * C * N1;
    for (int c = 0; c < C; ++c) s0[c * GM_SIMT_T + threadIdx.x] = i < N0 ? f0b[(size_t)c * N0 + i] : 0.f;
    const float sc = 1.4426950408889634f / sqrtf((float)C), inv = 1.f / sqrtf((float)C);
    float m = -INFINITY, l = 0.f, sx = 0.f, sy = 0.f;
    for (int j0 = 0; j0 < N1; j0 += GM_SIMT_TJ) {
        __syncthreads();
        for (int e = threadIdx.x; e < GM_SIMT_TJ * C; e += GM_SIMT_T) {
            int jj = e % GM_SIMT_TJ, c = e / GM_SIMT_TJ;
            s1[jj * C + c] = (j0 + jj < N1) ? f1b[(size_t)c * N1 + j0 + jj] : 0.f;
        }
        __syncthreads();
        for (int jj = 0; jj < GM_SIMT_TJ && j0 + jj < N1; ++jj) {
            float d = 0.f;
            for (int c = 0; c < C; ++c) d = fmaf(s0[c * GM_SIMT_T + threadIdx.x], s1[jj * C + c], d);
            const int j = j0 + jj;
            if (vol && i < N0) vol[((size_t)b * N1 + j) * N0 + i] = d * inv;
            const float s = d * sc;
            if (s > m) { const float r = ex2f(m - s); l *= r; sx *= r; sy *= r; m = s; }
            const float pj = ex2f(s - m);
            l += pj;
            sx = fmaf(pj, gm_linspace(W1, j % W1), sx);
            sy = fmaf(pj, gm_linspace(H1, j / W1), sy);
        }
    }
    if (i < N0) {
        flow[((size_t)b * 2 + 0) * N0 + i] = sx / l;
        flow[((size_t)b * 2 + 1) * N0 + i] = sy / l;
    }
}

// pos_embed on an already materialised volume [B,N1,N0] (reference: model/network.py:430-440).
__global__ void __launch_bounds__(128) pos_embed_kernel(const float* __restrict__ vol, float* __restrict__ flow,
                                                        int N0, int N1, int H1, int W1) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= N0) return;
    const float* v = vol + (size_t)b * N1 * N0 + i;
    float m = -INFINITY, l = 0.f, sx = 0.f, sy = 0.f;
    for (int j = 0; j < N1; ++j) {
        const float s = __ldg(v + (size_t)j * N0) * 1.4426950408889634f;
        if (s > m) { const float r = ex2f(m - s); l *= r; sx *= r; sy *= r; m = s; }
        const float pj = ex2f(s - m);
        l += pj;
        sx = fmaf(pj, gm_linspace(W1, j % W1), sx);
        sy = fmaf(pj, gm_linspace(H1, j / W1), sy);
    }
    flow[((size_t)b * 2 + 0) * N0 + i] = sx / l;
    flow[((size_t)b * 2 + 1) * N0 + i] = sy / l;
}

// ------------------------------------------------------------------------------------------------
// tcgen05 version.
namespace tc {

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, 128B swizzle, 8-row groups 1024 B apart (sm_100 format)
__device__ __forceinline__ uint64_t smem_desc_k128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address
    d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset
    d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M x N
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

constexpr int TM = 128;              // rows per tile (source positions / target positions)
constexpr int ATOM_BYTES = TM * 128; // one K atom: 128 rows x 32 fp32

// Load rows [n0, n0+128) x channels [0, KA*32) of src[C][N] (zero-padded), split hi/lo, store swizzled.
template <int KA>
__device__ __forceinline__ void load_split_tile(unsigned char* hi, unsigned char* lo, const float* __restrict__ src,
                                                int C, int N, int n0, bool want_lo, const int r) {
    // one tile row per loader thread: coalesced across the 128 loaders for every channel
    const int n = n0 + r;
    const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
    // 32 loads in flight per thread before anything is converted: with one 128-thread CTA per SM the tile load is pure
    // latency (ncu: 40 % of the samples waited on these loads when only 8 were in flight)
#pragma unroll 1
    for (int cb = 0; cb < KA; ++cb) {
        float v[8][4];
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = (cb * 8 + q) * 4 + e;
                v[q][e] = (n < N && c < C) ? __ldg(src + (size_t)c * N + n) : 0.f;
            }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int cq = cb * 8 + q;
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(v[q][0]) & 0xFFFFE000u); l.x = v[q][0] - h.x;
            h.y = __uint_as_float(__float_as_uint(v[q][1]) & 0xFFFFE000u); l.y = v[q][1] - h.y;
            h.z = __uint_as_float(__float_as_uint(v[q][2]) & 0xFFFFE000u); l.z = v[q][2] - h.z;
            h.w = __uint_as_float(__float_as_uint(v[q][3]) & 0xFFFFE000u); l.w = v[q][3] - h.w;
            const uint32_t off = (uint32_t)(cq >> 3) * ATOM_BYTES + row_off + (uint32_t)(((cq & 7) ^ (r & 7)) << 4);
            *reinterpret_cast<float4*>(hi + off) = h;
            if (want_lo) *reinterpret_cast<float4*>(lo + off) = l;
        }
    }
}

// Roles: warps 0-3 = softmax epilogue (thread = TMEM lane = source position), warps 4-7 = loaders (thread = tile row: NCHW
// -> TF32 hi/lo, K-major, swizzled), warp 8 = MMA issuer.  They meet only through mbarriers, so the conversion of target tile
// j+1 runs while tile j is in the tensor core and tile j-1 in the softmax (round 1 ran the three phases back to back in one
// 4-warp CTA: tensor pipe 10 %).
constexpr int GM_THREADS = 9 * 32;

template <int KA>
__global__ void __launch_bounds__(GM_THREADS, 1) gm_tc_kernel(const float* __restrict__ f0, const float* __restrict__ f1,
                                                              float* __restrict__ flow, float* __restrict__ vol,
                                                              int C, int N0, int N1, int H1, int W1, int precision) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int TILE_BYTES = KA * ATOM_BYTES;
    unsigned char* a_hi = smem;
    unsigned char* a_lo = smem + TILE_BYTES;
    unsigned char* b_hi[2] = {smem + 2 * TILE_BYTES, smem + 4 * TILE_BYTES};
    unsigned char* b_lo[2] = {smem + 3 * TILE_BYTES, smem + 5 * TILE_BYTES};
    __shared__ uint64_t mma_bar[2];      // MMA into accumulator / from B buffer `buf` retired
    __shared__ uint64_t b_full[2];       // B buffer written by the 128 loader threads
    __shared__ uint64_t d_free[2];       // accumulator drained by the 4 epilogue warps
    __shared__ uint64_t a_full;
    __shared__ uint32_t tmem_base_s;
    __shared__ float gxs[64], gys[64];  // not used when W1/H1 > 64 (falls back to gm_linspace)

    const int b = blockIdx.y, i0 = blockIdx.x * TM;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool want_lo = precision == 0;
    const float* f0b = f0 + (size_t)b * C * N0;
    const float* f1b = f1 + (size_t)b * C * N1;

    if (threadIdx.x == 0) {
        for (int k = 0; k < 2; ++k) { mbar_init(&mma_bar[k], 1); mbar_init(&b_full[k], 128); mbar_init(&d_free[k], 4); }
        mbar_init(&a_full, 128);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (threadIdx.x < 64) {
        gxs[threadIdx.x] = gm_linspace(W1, min((int)threadIdx.x, W1 - 1));
        gys[threadIdx.x] = gm_linspace(H1, min((int)threadIdx.x, H1 - 1));
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;
    constexpr uint32_t IDESC = idesc_tf32(TM, TM);
    const int ntiles = (N1 + TM - 1) / TM;

    if (warp >= 4 && warp < 8) {
        // ================= loaders =================
        const int r = threadIdx.x - 128;
        load_split_tile<KA>(a_hi, a_lo, f0b, C, N0, i0, want_lo, r);
        fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(&a_full);
        for (int jt = 0; jt < ntiles; ++jt) {
            const int buf = jt & 1;
            if (jt >= 2) mbar_wait(&mma_bar[buf], ((jt >> 1) - 1) & 1);      // MMA(jt - 2) has read B[buf]
            load_split_tile<KA>(b_hi[buf], b_lo[buf], f1b, C, N1, jt * TM, want_lo, r);
            fence_proxy_async();
            mbar_arrive(&b_full[buf]);
        }
    } else if (warp == 8) {
        // ================= MMA issuer =================
        if (lane == 0) {
            mbar_wait(&a_full, 0);
            for (int jt = 0; jt < ntiles; ++jt) {
                const int buf = jt & 1;
                mbar_wait(&b_full[buf], (jt >> 1) & 1);
                if (jt >= 2) mbar_wait(&d_free[buf], ((jt >> 1) - 1) & 1);   // epilogue(jt - 2) has drained TMEM[buf]
                fence_after_sync();
                const uint32_t d = tmem_base + (uint32_t)buf * TM;           // D[buf] = A * B[buf]^T over K = KA*32 in steps of 8
                uint32_t acc = 0;
#pragma unroll
                for (int ka = 0; ka < KA; ++ka) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t off = (uint32_t)ka * ATOM_BYTES + (uint32_t)ks * 32u;
                        const uint64_t ah = smem_desc_k128(smem_u32(a_hi) + off), bh = smem_desc_k128(smem_u32(b_hi[buf]) + off);
                        if (want_lo) {
                            const uint64_t al = smem_desc_k128(smem_u32(a_lo) + off), bl = smem_desc_k128(smem_u32(b_lo[buf]) + off);
                            mma_tf32(d, al, bh, IDESC, acc); acc = 1;
                            mma_tf32(d, ah, bl, IDESC, acc);
                        }
                        mma_tf32(d, ah, bh, IDESC, acc); acc = 1;
                    }
                }
                mma_commit(&mma_bar[buf]);
            }
        }
    } else {
        // ================= epilogue: online softmax x grid expectation =================
        const int i = i0 + threadIdx.x;          // this thread's source position = TMEM lane
        const float inv = 1.f / sqrtf((float)C), sc = 1.4426950408889634f * inv;
        const bool small_grid = W1 <= 64 && H1 <= 64;
        // Fast path (W1 >= 32, whole chunks): the grid is linear in the column / row index, so a chunk of 32 target positions
        // only needs P = sum p, T = sum p * e (e = position inside the chunk, an immediate) and Q = sum of the p past the one
        // row wrap a chunk can contain:  sum p * jx = jx0 P + T - W1 Q,  sum p * jy = jy0 P + Q.  Six instructions per
        // element instead of ~38 (the per-element row / column bookkeeping made the softmax warps the bottleneck).
        const bool fast = W1 >= 32 && !vol;
        float m = -INFINITY, l = 0.f, sx = 0.f, sy = 0.f;      // fast path: sx / sy accumulate p * jx / p * jy
        for (int jt = 0; jt < ntiles; ++jt) {
            const int buf = jt & 1;
            mbar_wait(&mma_bar[buf], (jt >> 1) & 1);
            fence_after_sync();
            const int j0 = jt * TM;
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * TM + ch * 32), r);
                tmem_ld_wait();
                if (ch == 3) {                   // the accumulator is in registers: hand it back before the arithmetic
                    fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&d_free[buf]);
                }
                const int jb = j0 + ch * 32;
                if (jb >= N1) continue;
                int jy = jb / W1, jx = jb - jy * W1;
                if (fast && jb + 32 <= N1) {
                    float dmax = __uint_as_float(r[0]);
#pragma unroll
                    for (int e = 1; e < 32; ++e) dmax = fmaxf(dmax, __uint_as_float(r[e]));
                    const float cmax = dmax * sc;
                    if (cmax > m) { const float rs = ex2f(m - cmax); l *= rs; sx *= rs; sy *= rs; m = cmax; }
                    const float nm = -m;
                    const int ewrap = W1 - jx;                      // elements e >= ewrap sit on the next grid row
                    float P = 0.f, T = 0.f, Q = 0.f;
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const float pj = ex2f(fmaf(__uint_as_float(r[e]), sc, nm));
                        P += pj;
                        T = fmaf(pj, (float)e, T);
                        if (e >= ewrap) Q += pj;
                    }
                    l += P;
                    sx += fmaf((float)jx, P, T) - (float)W1 * Q;
                    sy += fmaf((float)jy, P, Q);
                    continue;
                }
                float cmax = -INFINITY;
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float d = __uint_as_float(r[e]);
                    if (vol && i < N0 && jb + e < N1) vol[((size_t)b * N1 + jb + e) * N0 + i] = d * inv;
                    cmax = fmaxf(cmax, (jb + e < N1) ? d * sc : -INFINITY);
                }
                if (cmax > m) { const float rs = ex2f(m - cmax); l *= rs; sx *= rs; sy *= rs; m = cmax; }
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float s = (jb + e < N1) ? __uint_as_float(r[e]) * sc : -INFINITY;
                    const float pj = ex2f(s - m);
                    // fast mode accumulates indices (the partial last chunk of a fast run lands here)
                    const float gx = fast ? (float)jx : small_grid ? gxs[jx] : gm_linspace(W1, jx);
                    const float gy = fast ? (float)jy : small_grid ? gys[min(jy, 63)] : gm_linspace(H1, jy);
                    l += pj;
                    sx = fmaf(pj, gx, sx);
                    sy = fmaf(pj, gy, sy);
                    if (++jx == W1) { jx = 0; ++jy; }
                }
            }
        }
        if (i < N0) {
            float fx = sx / l, fy = sy / l;
            if (fast) {          // expectation of the index -> expectation of torch.linspace(-1 + 1/n, 1 - 1/n, n)
                const float x0 = (float)(-1.0 + 1.0 / (double)W1), x1 = (float)(1.0 - 1.0 / (double)W1);
                const float y0 = (float)(-1.0 + 1.0 / (double)H1), y1 = (float)(1.0 - 1.0 / (double)H1);
                fx = fmaf((x1 - x0) / (float)(W1 - 1), fx, x0);
                fy = H1 > 1 ? fmaf((y1 - y0) / (float)(H1 - 1), fy, y0) : y0;
            }
            flow[((size_t)b * 2 + 0) * N0 + i] = fx;
            flow[((size_t)b * 2 + 1) * N0 + i] = fy;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace tc
}  // namespace gfb

using namespace gfb;

extern "C" int gfb_pos_embed_f32(const float* vol, float* flow_out, int B, int H0, int W0, int H1, int W1,
                                 gfb_stream_t stream) {
    GFB_CHECK_ARG(vol && flow_out && B > 0 && B <= 65535 && H0 > 0 && W0 > 0 && H1 > 0 && W1 > 0);
    dim3 grid((H0 * W0 + 127) / 128, B);
    pos_embed_kernel<<<grid, 128, 0, gfb_cu(stream)>>>(vol, flow_out, H0 * W0, H1 * W1, H1, W1);
    GFB_LAUNCH_RESULT();
}

extern "C" int gfb_global_match_f32(const float* f0, const float* f1, float* flow_out, float* vol_out,
                                    int B, int C, int H0, int W0, int H1, int W1,
                                    int precision, int algo, gfb_stream_t stream) {
    GFB_CHECK_ARG(f0 && f1 && flow_out && B > 0 && C > 0 && H0 > 0 && W0 > 0 && H1 > 0 && W1 > 0);
    GFB_CHECK_ARG(precision == 0 || precision == 1);
    GFB_CHECK_ARG(algo >= 0 && algo <= 2 && B <= 65535);
    const int N0 = H0 * W0, N1 = H1 * W1;
    cudaStream_t st = gfb_cu(stream);
    const bool tc_ok = C <= 64;
    if (algo == 2 && !tc_ok) return GFB_EUNSUPPORTED;
    if (algo != 1 && tc_ok) {
        dim3 grid((N0 + tc::TM - 1) / tc::TM, B);
        cudaError_t e;
        if (C <= 32) {
            const int smem = 6 * 1 * tc::ATOM_BYTES;
            e = cudaFuncSetAttribute(tc::gm_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
            tc::gm_tc_kernel<1><<<grid, tc::GM_THREADS, smem, st>>>(f0, f1, flow_out, vol_out, C, N0, N1, H1, W1, precision);
        } else {
            const int smem = 6 * 2 * tc::ATOM_BYTES;
            e = cudaFuncSetAttribute(tc::gm_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return (int)e;
            tc::gm_tc_kernel<2><<<grid, tc::GM_THREADS, smem, st>>>(f0, f1, flow_out, vol_out, C, N0, N1, H1, W1, precision);
        }
        GFB_LAUNCH_RESULT();
    }
    const size_t smem = ((size_t)C * GM_SIMT_T + (size_t)GM_SIMT_TJ * C) * sizeof(float);
    if (smem > 200 * 1024) return GFB_EUNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(gm_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((N0 + GM_SIMT_T - 1) / GM_SIMT_T, B);
    gm_simt_kernel<<<grid, GM_SIMT_T, smem, st>>>(f0, f1, flow_out, vol_out, C, N0, N1, H1, W1);
    GFB_LAUNCH_RESULT();
}
