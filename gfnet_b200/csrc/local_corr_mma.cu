// K1, third generation (reference: utils/local_correlation.py:4-72, call site model/network.py:553-554).
//
//   corr[b,k,gy,gx] = (1/sqrt C) sum_c f0[b,c,gy,gx] * bilinear(f1[b,c], flow[b,:,gy,gx] + off_k)
//                   = bilerp(D),  D[j,i] = sum_c f0[c] * f1[c, y0-r+j, x0-r+i]  over the (2r+2)^2 integer patch
//
// lc_mma_kernel: D on the warp-level tensor-core path (mma.sync m16n8k16, bf16 hi/lo split, fp32 accumulate) with every
// operand taken from the fp32 NCHW tensors as they are -- no pre-pass, no workspace, no TMEM hand-off.
//
//  * A warp owns 8 x 4 lattice points = two M-blocks of 8 x 2 points (rows of the A fragments); the bf16 hi / lo parts of
//    their f0 vectors stay in registers for the whole tile.
//  * The CTA (NWX x NWY warps) streams the image rows of its bounding box through a two-stage shared-memory ring:
//    one row of all C channels per stage ([c][PL] fp32, cp.async 16 B with zero fill = padding_mode "zeros").
//  * Per image row a warp multiplies its A fragments with N-tiles of 8 consecutive positions starting at the warp's own
//    leftmost window origin (B fragments: four LDS.32 per tile and K step, split into bf16 hi / lo on the fly; the plane
//    pitch PL = 4 (mod 8) words makes them bank-conflict free).  hi*hi + hi*lo + lo*hi, relative error ~1e-5.
//  * The accumulators (point x position) go through a warp-private staging tile; lane = lattice point then reads its own
//    W columns at its own offset (transposed layout: conflict free), x-lerps, y-lerps against the previous image row kept
//    in registers and stores: for a fixed k the 32 lanes write 4 x 32 contiguous bytes.
//  * Work per 16 points and image row is proportional to the span of THEIR windows (3-4 N-tiles), not to a 128-point tile:
//    ~40 % of the multiplied (point, position) pairs are used, against ~10 % in the tcgen05 formulation (lc_tc2_kernel),
//    which is what pays for the 4x lower issue rate of mma.sync.
//
// Points whose windows do not fit the staged box (wild flows) take the exact per-sample gather (lc_generic_point).
#include "common.cuh"
#include "lc_common.cuh"
#include <limits.h>

namespace gfb {
namespace lcm {

constexpr int DPITCH = 36;            // staging tile: [column][DPITCH] floats, lane = point (4 * column + point: conflict-free stores)

__device__ __forceinline__ uint32_t bf16x2_rn(float upper, float lower) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
    return d;
}
// (v0, v1) -> packed bf16 hi parts (v0 in the low half) and packed bf16 residuals
__device__ __forceinline__ void split2(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    hi = bf16x2_rn(v1, v0);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    lo = bf16x2_rn(v1 - h1, v0 - h0);
}
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void out_put(float* ptr, float v, int accumulate) {
    if (accumulate) v += __ldcs(ptr);
    __stcs(ptr, v);
}

__device__ unsigned long long g_mma_stats[4];   // [0] points on the gather path

// One image row of one warp: D[point, position] for NTW N-tiles of 8 positions and the active M-blocks, accumulators to
// the staging tile.  st = this lane's first B element (channel 2t, position xoff + g) in the ring stage.
template <int C, int NTW, bool M0, bool M1>
__device__ __forceinline__ void row_mma(const float* __restrict__ st, const int PL,
                                        const uint32_t (&ahi)[2][C / 16][4], const uint32_t (&alo)[2][C / 16][4],
                                        float* __restrict__ dst, const int g, const int t) {
    constexpr int NKS = C / 16;
    float acc[2][NTW][4];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int n = 0; n < NTW; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mb][n][e] = 0.f;
    const float* s0 = st;
    const float* s1 = st + PL;
    const float* s8 = st + 8 * PL;
    const float* s9 = st + 9 * PL;
    const int ksp = 16 * PL;
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
        uint32_t bh[NTW][2], bl[NTW][2];
#pragma unroll
        for (int n = 0; n < NTW; ++n) {
            split2(s0[ks * ksp + 8 * n], s1[ks * ksp + 8 * n], bh[n][0], bl[n][0]);
            split2(s8[ks * ksp + 8 * n], s9[ks * ksp + 8 * n], bh[n][1], bl[n][1]);
        }
#pragma unroll
        for (int n = 0; n < NTW; ++n) {
            if (M0) hmma(acc[0][n], alo[0][ks], bh[n][0], bh[n][1]);
            if (M1) hmma(acc[1][n], alo[1][ks], bh[n][0], bh[n][1]);
        }
#pragma unroll
        for (int n = 0; n < NTW; ++n) {
            if (M0) hmma(acc[0][n], ahi[0][ks], bl[n][0], bl[n][1]);
            if (M1) hmma(acc[1][n], ahi[1][ks], bl[n][0], bl[n][1]);
        }
#pragma unroll
        for (int n = 0; n < NTW; ++n) {
            if (M0) hmma(acc[0][n], ahi[0][ks], bh[n][0], bh[n][1]);
            if (M1) hmma(acc[1][n], ahi[1][ks], bh[n][0], bh[n][1]);
        }
    }
    // accumulators -> staging tile [column][point]: c0/c1 = (row g, columns 2t, 2t+1), c2/c3 = (row g + 8, ...)
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
        if (mb == 0 ? M0 : M1) {
#pragma unroll
            for (int n = 0; n < NTW; ++n) {
                float* d = dst + (8 * n + 2 * t) * DPITCH + mb * 16 + g;
                d[0] = acc[mb][n][0]; d[DPITCH] = acc[mb][n][1];
                d[8] = acc[mb][n][2]; d[DPITCH + 8] = acc[mb][n][3];
            }
        }
}
template <int C, int NTW>
__device__ __forceinline__ void row_dispatch(const bool m0, const bool m1, const float* __restrict__ st, const int PL,
                                             const uint32_t (&ahi)[2][C / 16][4], const uint32_t (&alo)[2][C / 16][4],
                                             float* __restrict__ dst, const int g, const int t) {
    if (m0 && m1) row_mma<C, NTW, true, true>(st, PL, ahi, alo, dst, g, t);
    else if (m0) row_mma<C, NTW, true, false>(st, PL, ahi, alo, dst, g, t);
    else row_mma<C, NTW, false, true>(st, PL, ahi, alo, dst, g, t);
}

struct MmaCfg {
    int PL;                 // plane pitch of a ring stage in floats (multiple of 4, = 4 mod 8)
    int tiles_x, tiles_y;
    int rmax;               // image rows a CTA streams at most; points whose windows end later take the gather
    int lq;                 // loader lanes per channel row: 16 or 32, >= PL / 4
};

template <int R, int C, int NT, int NWX, int NWY, int MINB>
__global__ void __launch_bounds__(NWX * NWY * 32, MINB)
lc_mma_kernel(const LcParams p, const MmaCfg cfg) {
    constexpr int W = 2 * R + 2, KW = 2 * R + 1, KK = KW * KW;
    constexpr int NKS = C / 16, NW = NWX * NWY, NTHREADS = NW * 32;
    static_assert(C % 16 == 0 && W <= 8 * NT, "shape");
    extern __shared__ __align__(16) float smem[];
    const int PL = cfg.PL;
    float* ring = smem;                                            // [2][C][PL]
    float* dst_all = ring + 2 * C * PL;                            // [NW][8 NT][DPITCH]
    int* sbox = reinterpret_cast<int*>(dst_all + NW * 8 * NT * DPITCH);   // [0] min x, [1] min y, [2] max y + W

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int G = p.G;
    const size_t gg = (size_t)G * G;
    int tile = blockIdx.x;
    const int tx = tile % cfg.tiles_x; tile /= cfg.tiles_x;
    const int ty = tile % cfg.tiles_y;
    const int b = tile / cfg.tiles_y;
    const int wx = warp % NWX, wy = warp / NWX;
    const int gx0 = (tx * NWX + wx) * 8, gy0 = (ty * NWY + wy) * 4;          // the warp's 8 x 4 points
    const int gx = gx0 + (lane & 7), gy = gy0 + (lane >> 3);                 // this lane's point (epilogue role)
    const bool valid = gx < G && gy < G;

    if (tid == 0) { sbox[0] = INT_MAX; sbox[1] = INT_MAX; sbox[2] = INT_MIN; }
    const PointGeom pg = point_geom(p, b, gy, gx, valid, R);
    __syncthreads();
    {
        const int mx = warp_min(pg.live ? pg.xb : INT_MAX), my = warp_min(pg.live ? pg.yb : INT_MAX);
        if (lane == 0 && mx != INT_MAX) { atomicMin(&sbox[0], mx); atomicMin(&sbox[1], my); }
    }
    __syncthreads();
    const int bx = sbox[0];
    const bool any_live = bx != INT_MAX;
    const int X0 = any_live ? (bx & ~3) : 0, Y0 = any_live ? sbox[1] : 0;     // 16-byte aligned column origin
    // fit: the window lies inside the CTA's staged columns / row budget and inside the warp's N-tiles
    const bool fit1 = pg.live && pg.xb - X0 + W <= PL && pg.yb - Y0 + W <= cfg.rmax;
    const int amin = warp_min(fit1 ? pg.xb : INT_MAX);
    const bool fit = fit1 && pg.xb - amin + W <= 8 * NT;
    const int amax = warp_max(fit ? pg.xb : INT_MIN);
    const int ntw = amax == INT_MIN ? 0 : (amax - amin + W + 7) >> 3;         // N-tiles this warp multiplies per row
    {
        const int my = warp_max(fit ? pg.yb + W : INT_MIN);
        if (lane == 0 && my != INT_MIN) atomicMax(&sbox[2], my);
    }

    // ---- A fragments: f0 of the warp's points, bf16 hi / lo, in registers for the whole tile -------------------------
    // M-block mb = lattice rows 2 mb, 2 mb + 1 of the warp; fragment row m <-> point (gx0 + (m & 7), gy0 + 2 mb + (m >> 3))
    uint32_t ahi[2][NKS][4], alo[2][NKS][4];
    {
        // 32-bit element indices (the launcher checks that f0 / f1 have fewer than 2^31 elements)
        const unsigned g32 = (unsigned)gg;
        const unsigned f0base = ((unsigned)b * (unsigned)p.f0_ctot + (unsigned)p.c0) * g32;
        const int ax = gx0 + g;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
            for (int ks = 0; ks < NKS; ++ks) {
                float v[2][4];                                       // [h: point row g / g + 8][channel 2t, 2t+1, 2t+8, 2t+9]
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int ay = gy0 + 2 * mb + h;
                    const bool ok = ax < G && ay < G;
                    const unsigned i0 = f0base + (unsigned)(16 * ks + 2 * t) * g32 + (unsigned)(ay * G + ax);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        v[h][e] = ok ? __ldg(p.f0 + (i0 + (unsigned)((e & 1) + 8 * (e >> 1)) * g32)) : 0.f;
                }
                split2(v[0][0], v[0][1], ahi[mb][ks][0], alo[mb][ks][0]);
                split2(v[1][0], v[1][1], ahi[mb][ks][1], alo[mb][ks][1]);
                split2(v[0][2], v[0][3], ahi[mb][ks][2], alo[mb][ks][2]);
                split2(v[1][2], v[1][3], ahi[mb][ks][3], alo[mb][ks][3]);
            }
    }
    __syncthreads();
    const int nrows = sbox[2] == INT_MIN ? 0 : sbox[2] - Y0;

    // ---- row streaming ------------------------------------------------------------------------------------------------
    // Loader: LQ lanes per channel row (one 16-byte group each), NTHREADS / LQ channels per pass.  Everything but the image
    // row is fixed per tile, so a row costs one cp.async and two adds per group (32-bit element indices).
    const int Q = PL >> 2;                                           // 16-byte groups per plane
    const bool lq32 = cfg.lq == 32;                                  // lq = 16 or 32 >= Q (launcher)
    const int lxg = lq32 ? (tid & 31) : (tid & 15), lcg = lq32 ? (tid >> 5) : (tid >> 4);
    const unsigned plane = (unsigned)(p.Hs * p.pitch);
    const int lx = X0 + 4 * lxg;
    const int nvx = (lxg < Q && lx >= 0) ? min(max(p.Ws - lx, 0), 4) : 0;     // valid floats of this thread's group
    const unsigned f1base = ((unsigned)b * (unsigned)p.Ctot + (unsigned)p.c0) * plane;
    const unsigned lsrc = f1base + (unsigned)lcg * plane + (unsigned)(nvx ? lx : 0);
    const uint32_t ldst = smem_u32(ring) + (uint32_t)(lcg * PL + 4 * lxg) * 4u;
    auto load_row = [&](int q) {
        if (lxg < Q) {
            const int y = Y0 + q;
            const bool yin = (unsigned)y < (unsigned)p.Hs;
            const uint32_t nb = yin ? (uint32_t)nvx * 4u : 0u;
            unsigned si = nb ? lsrc + (unsigned)(y * p.pitch) : 0u;
            uint32_t d = ldst + (uint32_t)((q & 1) * C * PL) * 4u;
            if (lq32) {
                constexpr int NCG = NTHREADS / 32, NIT = (C + NCG - 1) / NCG;
                const unsigned sstep = nb ? NCG * plane : 0u;
                const uint32_t dstep = (uint32_t)(NCG * PL) * 4u;
#pragma unroll
                for (int k = 0; k < NIT; ++k) {
                    if (C % NCG == 0 || lcg + k * NCG < C) cp_async16(d, p.f1 + si, nb);
                    si += sstep; d += dstep;
                }
            } else {
                constexpr int NCG = NTHREADS / 16, NIT = (C + NCG - 1) / NCG;
                const unsigned sstep = nb ? NCG * plane : 0u;
                const uint32_t dstep = (uint32_t)(NCG * PL) * 4u;
#pragma unroll
                for (int k = 0; k < NIT; ++k) {
                    if (C % NCG == 0 || lcg + k * NCG < C) cp_async16(d, p.f1 + si, nb);
                    si += sstep; d += dstep;
                }
            }
        }
        cp_async_commit();
    };

    float* dst = dst_all + (size_t)warp * 8 * NT * DPITCH;
    const int xoff = ntw ? amin - X0 : 0;
    const int col0 = fit ? pg.xb - amin : 0;
    const float a1 = pg.fx, a0 = 1.f - pg.fx;
    const float wy1 = pg.fy * p.inv_sqrt_c, wy0 = (1.f - pg.fy) * p.inv_sqrt_c;
    float hprev[KW];
#pragma unroll
    for (int i = 0; i < KW; ++i) hprev[i] = 0.f;
    const unsigned gg32 = (unsigned)gg;
    // element index of this lane's first output; the launcher guarantees B * k_total * G * G < 2^31
    const unsigned lane_idx = (((unsigned)b * (unsigned)p.k_total + (unsigned)p.k_offset) * (unsigned)G + (unsigned)gy) * (unsigned)G + (unsigned)gx;

    if (nrows > 0) load_row(0);
    cp_async_wait_all();
    __syncthreads();
    for (int q = 0; q < nrows; ++q) {
        if (q + 1 < nrows) load_row(q + 1);
        const int j = Y0 + q - pg.yb;
        const bool act = fit && (unsigned)j < (unsigned)W;
        const unsigned ball = __ballot_sync(0xffffffffu, act);
        if (ball) {
            const bool m0 = (ball & 0xffffu) != 0, m1 = (ball >> 16) != 0;
            const float* st = ring + (size_t)(q & 1) * C * PL + 2 * t * PL + xoff + g;
            // straight-line code per (N-tiles, active M-blocks): the conditions are warp-uniform, which the compiler cannot see
            switch (ntw) {
                case 1: row_dispatch<C, 1>(m0, m1, st, PL, ahi, alo, dst, g, t); break;
                case 2: row_dispatch<C, 2>(m0, m1, st, PL, ahi, alo, dst, g, t); break;
                case 3: row_dispatch<C, 3>(m0, m1, st, PL, ahi, alo, dst, g, t); break;
                case 4: row_dispatch<C, (NT >= 4 ? 4 : NT)>(m0, m1, st, PL, ahi, alo, dst, g, t); break;
                default: row_dispatch<C, NT>(m0, m1, st, PL, ahi, alo, dst, g, t); break;
            }
            __syncwarp();
            if (act) {
                const float* rowp = dst + col0 * DPITCH + lane;
                float hv[KW];
                float d0 = rowp[0];
#pragma unroll
                for (int i = 0; i < KW; ++i) {
                    const float d1 = rowp[(i + 1) * DPITCH];
                    hv[i] = a0 * d0 + a1 * d1;
                    d0 = d1;
                }
                if (j >= 1) {
                    float* o = p.out + (lane_idx + (unsigned)(j - 1) * (KW * gg32));
                    if (!p.accumulate) {
#pragma unroll
                        for (int i = 0; i < KW; ++i) __stcs(o + (size_t)i * gg, wy0 * hprev[i] + wy1 * hv[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < KW; ++i) __stcs(o + (size_t)i * gg, __ldcs(o + (size_t)i * gg) + wy0 * hprev[i] + wy1 * hv[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < KW; ++i) hprev[i] = hv[i];
            }
            __syncwarp();
        }
        cp_async_wait_all();
        __syncthreads();
    }

    // ---- points outside the streamed path: zeros for dead windows, the exact gather (whole warp per point) for the rest ----
    if (valid && !pg.live && !p.accumulate) {
        float* outp = p.out + ((size_t)b * p.k_total + p.k_offset) * gg + (size_t)gy * G + gx;
        for (int k = 0; k < KK; ++k) st_stream(outp + (size_t)k * gg, 0.f);
    }
    unsigned todo = __ballot_sync(0xffffffffu, valid && pg.live && !fit);
    if (todo && lane == 0) atomicAdd(&g_mma_stats[0], (unsigned long long)__popc(todo));
    while (todo) {
        const int l = __ffs(todo) - 1;
        todo &= todo - 1;
        const int px = gx0 + (l & 7), py = gy0 + (l >> 3);
        float* outp = p.out + ((size_t)b * p.k_total + p.k_offset) * gg + (size_t)py * G + px;
        for (int k = lane; k < KK; k += 32) out_put(outp + (size_t)k * gg, lc_generic_point(p, b, k, py, px), p.accumulate);
    }
}

template <int R, int C, int NT, int NWX, int NWY, int MINB>
static int launch_mma(const LcParams& p, cudaStream_t st) {
    constexpr int W = 2 * R + 2, NW = NWX * NWY;
    const int G = p.G;
    if ((size_t)p.B * p.k_total * G * G >= (1ull << 31)) return GFB_EUNSUPPORTED;     // 32-bit output indices
    if ((size_t)p.B * p.f0_ctot * G * G >= (1ull << 31) || (size_t)p.B * p.Ctot * p.Hs * p.pitch >= (1ull << 31)) return GFB_EUNSUPPORTED;
    if (p.pitch % 4 != 0 || !gfb_aligned(p.f1, 16)) return GFB_EALIGN;                 // 16-byte cp.async sources
    MmaCfg c;
    const float s = (float)p.Ws / (float)G;
    // staged columns: span of the CTA's window origins (1.35 = magnification head-room) + window + alignment + shear
    int need = (int)ceilf((8 * NWX - 1) * s * 1.4f) + W + 3 + 8;
    need = max(need, 8 * NT);
    c.PL = ((need + 3) / 8) * 8 + 4;                       // smallest 8 m + 4 >= need
    if (c.PL < need) c.PL += 8;
    if (c.PL > 128) return GFB_EUNSUPPORTED;                 // one loader lane per 16-byte group, at most 32 per channel row
    c.lq = c.PL <= 64 ? 16 : 32;
    c.tiles_x = (G + 8 * NWX - 1) / (8 * NWX);
    c.tiles_y = (G + 4 * NWY - 1) / (4 * NWY);
    c.rmax = (int)ceilf((4 * NWY - 1) * s * 1.4f + (8 * NWX - 1) * s * 0.45f) + W + 8;   // + shear of a rotated lattice row
    const long long tiles = (long long)p.B * c.tiles_x * c.tiles_y;
    if (tiles > 0x7fffffffLL) return GFB_EUNSUPPORTED;
    const size_t smem = ((size_t)2 * C * c.PL + (size_t)NW * 8 * NT * DPITCH) * sizeof(float) + 16;
    if (smem > 200 * 1024) return GFB_EUNSUPPORTED;
    auto kern = lc_mma_kernel<R, C, NT, NWX, NWY, MINB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<(unsigned)tiles, NW * 32, smem, st>>>(p, c);
    GFB_LAUNCH_RESULT();
}

}  // namespace lcm
}  // namespace gfb

using namespace gfb;

// (r, C) the kernel is instantiated for: the GFNet scales (2,16) (4,32) (6,64) (7,64) and the microbench sweep's radii
#define GFB_MMA_ALL \
    GFB_MMA_CASE(2, 16, 4) GFB_MMA_CASE(1, 16, 3) GFB_MMA_CASE(2, 32, 4) GFB_MMA_CASE(3, 32, 4) GFB_MMA_CASE(4, 32, 4) \
    GFB_MMA_CASE(2, 64, 4) GFB_MMA_CASE(3, 64, 4) GFB_MMA_CASE(4, 64, 4) GFB_MMA_CASE(5, 64, 5) GFB_MMA_CASE(6, 64, 5) \
    GFB_MMA_CASE(7, 64, 5) GFB_MMA_CASE(8, 64, 5)

// One channel slice [c0, c0 + C) of tensors with Ctot channels, scaled by 1/sqrt(Ctot), stored (accumulate = 0) or added to
// out (accumulate = 1); Ctot = C, c0 = 0, accumulate = 0 is the plain operator.  shape: 0 = auto, else NWX | NWY << 4.
extern "C" int gfb_local_corr_mma_f32(const float* f0, const float* f1, const float* flow, float* out,
                                      int B, int C, int Ctot, int c0, int accumulate, int Hs, int Ws, int f1_pitch, int G, int r,
                                      int k_total, int k_offset, int shape, gfb_stream_t stream) {
    GFB_CHECK_ARG(f0 && f1 && flow && out);
    GFB_CHECK_ARG(B > 0 && C > 0 && Hs > 0 && Ws > 0 && G > 0 && r >= 0);
    GFB_CHECK_ARG(f1_pitch == 0 || f1_pitch >= Ws);
    const int kk = (2 * r + 1) * (2 * r + 1);
    GFB_CHECK_ARG(k_offset >= 0 && k_offset + kk <= k_total);
    LcParams p;
    p.f0 = f0; p.f1 = f1; p.flow = flow; p.out = out;
    p.B = B; p.C = C; p.Hs = Hs; p.Ws = Ws; p.G = G; p.r = r;
    p.Ctot = Ctot > 0 ? Ctot : C; p.c0 = c0; p.accumulate = accumulate ? 1 : 0; p.f0_ctot = p.Ctot;
    GFB_CHECK_ARG(c0 >= 0 && c0 + C <= p.Ctot);
    p.pitch = f1_pitch ? f1_pitch : Ws;
    p.k_total = k_total; p.k_offset = k_offset;
    p.sample_mode = 0; p.padding_mode = 0;
    p.ox0 = (float)(-2.0 * r / Ws); p.ox1 = (float)(2.0 * r / Ws);
    p.oy0 = (float)(-2.0 * r / Hs); p.oy1 = (float)(2.0 * r / Hs);
    p.inv_sqrt_c = (float)(1.0 / sqrt((double)p.Ctot));
    p.debug = 0;
    return gfb::lc_mma_launch(p, shape, gfb_cu(stream));
}

int gfb::lc_mma_launch(const LcParams& p, int shape, cudaStream_t st) {
    const int r = p.r, C = p.C;
    // CTA shapes: 4 x 2 warps (32 x 8 points, fewest staged bytes per point), 4 x 1, 2 x 2, 2 x 1
    int nwx = shape & 15, nwy = (shape >> 4) & 15;
    if (shape == 0) {
        // 2 x 1 warps (16 x 4 points) measured fastest on every GFNet shape: the row barrier couples fewer warps and the
        // bounding box of a tile is small; wider CTAs stage fewer bytes per point and win only when L2 traffic binds
        nwx = 2; nwy = 1;
    }
#define GFB_MMA_CASE(RR, CC, NTT) \
    if (r == RR && C == CC) { \
        if constexpr (CC <= 32) { if (nwx == 4 && nwy == 2) return lcm::launch_mma<RR, CC, NTT, 4, 2, 2>(p, st); } \
        if (nwx == 4 && nwy == 1) return lcm::launch_mma<RR, CC, NTT, 4, 1, (CC <= 32 ? 4 : 3)>(p, st); \
        if (nwx == 2 && nwy == 1) return lcm::launch_mma<RR, CC, NTT, 2, 1, (CC <= 32 ? 8 : 6)>(p, st); \
        if (nwx == 2 && nwy == 2) return lcm::launch_mma<RR, CC, NTT, 2, 2, (CC <= 32 ? 4 : 3)>(p, st); \
        return GFB_EUNSUPPORTED; \
    }
    GFB_MMA_ALL
#undef GFB_MMA_CASE
    return GFB_EUNSUPPORTED;
}

extern "C" int gfb_debug_local_corr_mma_counters(unsigned long long* host_out4, int reset) {
    cudaError_t e = cudaSuccess;
    if (host_out4) e = cudaMemcpyFromSymbol(host_out4, lcm::g_mma_stats, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess && reset) {
        unsigned long long z[4] = {0, 0, 0, 0};
        e = cudaMemcpyToSymbol(lcm::g_mma_stats, z, sizeof(z));
    }
    return e == cudaSuccess ? GFB_OK : (int)e;
}
