// K1 on tensor cores -- local correlation as a banded GEMM on tcgen05 (reference: utils/local_correlation.py:4-72,
// call site model/network.py:553-554).
//
//   corr[b,k,gy,gx] = (1/sqrt C) sum_c f0[b,c,gy,gx] * bilinear(f1[b,c], flow[b,:,gy,gx] + off_k)
//
// The window offsets are whole pixels, so every sample of a lattice point shares one fractional part and
// corr = bilerp(D), D[j,i] = sum_c f0[c] * f1[c, y0-r+j, x0-r+i] over a (2r+2)^2 integer patch.  D is a band of the
// GEMM  f0_tile[128 points x C] * f1_region[C x pixels]: a tile of 128 lattice points (TRy x TCx) is the M side,
// the pixels of the bounding box of the tile's windows, taken a few image rows at a time, are the N side.
//
//   plan kernel     one warp per tile: bounding box of the tile's windows -> TileDesc (global workspace)
//   converter warps read f0 / f1 (fp32, NCHW) from global memory, split every value into bf16 hi + lo
//                   and store K-major rows [hi(C) | lo(C)] with the 128-byte swizzle the tensor core expects
//   MMA thread      tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM), three products per K block:
//                   hi*hi + hi*lo + lo*hi  (relative error ~2^-17 of |f0||f1|: inside the 1e-4 fp32 tolerance)
//   epilogue warps  one TMEM lane = one lattice point: pull the image row out of TMEM (tcgen05.ld), stage it in a
//                   lane-private shared-memory row, read the lane's own W columns back at its data-dependent offset,
//                   interpolate in x, combine with the previous row in y, store the 2r+1 outputs of that row.
// The CTA is persistent (one per SM, tiles round-robin); rings of A stages, B stages and TMEM accumulators decouple
// the three roles.  Tiles whose windows do not fit the staged box (wild flow) fall back to exact per-sample gathers.
#include "common.cuh"
#include "lc_common.cuh"

namespace gfb {
namespace lctc {

struct TileDesc {
    int x0, y0;      // first column / row of the staged region (clipped to the image)
    int bw, nrows;   // region row width (multiple of 16) and number of image rows streamed
    int ylo, yhi;    // unclipped row range of the union of the windows
    int flags, pad;
};
enum { TF_EMPTY = 1, TF_GATHER = 2 };

struct TcCfg {
    int TCx, TRy, tiles_x, tiles_y, ntiles;
    int nmax;        // rows per B stage = columns per TMEM accumulator (128 or 256)
    int nstb, nsta, nacc;
    int bwmax;       // widest region row a tile may stream
    int pitch;       // floats per lane in the epilogue staging rows (= 4 mod 32)
    int ncw;         // converter warps
};

constexpr int EPI_WARPS = 4;
constexpr int MMA_WARP = 4;
constexpr int FIRST_CONV_WARP = 5;
constexpr int MAX_WARPS = 13;
constexpr int MARGIN = 16;          // zero columns either side of a staged row (>= 2r+2)

// ---- tcgen05 plumbing --------------------------------------------------------------------------------------
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
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N runtime
__device__ __forceinline__ uint32_t idesc_bf16(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t bf16x2_rn(float upper, float lower) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
    return d;
}

// ---- plan: bounding box of every tile's windows --------------------------------------------------------------
__global__ void __launch_bounds__(128) lc_plan_kernel(const LcParams p, const TcCfg c, TileDesc* __restrict__ plan) {
    const int tile = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (tile >= c.ntiles) return;
    const int R = p.r, W = 2 * R + 2, G = p.G;
    int t = tile;
    const int tx = t % c.tiles_x; t /= c.tiles_x;
    const int ty = t % c.tiles_y;
    const int b = t / c.tiles_y;
    const size_t gg = (size_t)G * G;
    int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN;
    for (int m = lane; m < 128; m += 32) {
        const int gy = ty * c.TRy + m / c.TCx, gx = tx * c.TCx + m % c.TCx;
        if (gy < G && gx < G) {
            const float* fl = p.flow + (size_t)b * 2 * gg + (size_t)gy * G + gx;
            const float sx = unnormalize(__ldg(fl), p.Ws), sy = unnormalize(__ldg(fl + gg), p.Hs);
            if (fabsf(sx) < 1e6f && fabsf(sy) < 1e6f) {
                const int xb = (int)floorf(sx) - R, yb = (int)floorf(sy) - R;
                if (!(xb >= p.Ws || xb + W <= 0 || yb >= p.Hs || yb + W <= 0)) {
                    xmin = min(xmin, xb); xmax = max(xmax, xb + W);
                    ymin = min(ymin, yb); ymax = max(ymax, yb + W);
                }
            }
        }
    }
    xmin = warp_min(xmin); xmax = warp_max(xmax); ymin = warp_min(ymin); ymax = warp_max(ymax);
    if (lane == 0) {
        TileDesc d;
        d.x0 = 0; d.y0 = 0; d.bw = 16; d.nrows = 0; d.ylo = 0; d.yhi = 0; d.flags = 0; d.pad = 0;
        if (xmin == INT_MAX) {
            d.flags = TF_EMPTY;
        } else {
            d.x0 = max(xmin, 0);
            d.y0 = max(ymin, 0);
            d.bw = (min(xmax, p.Ws) - d.x0 + 15) & ~15;
            d.nrows = min(ymax, p.Hs) - d.y0;
            d.ylo = ymin; d.yhi = ymax;
            if (d.bw > c.bwmax) d.flags = TF_GATHER;
        }
        plan[tile] = d;
    }
}

// ---- operand conversion ----------------------------------------------------------------------------------------
// 16 channels of one K-major row: load fp32 (stride cstride), split into bf16 hi / lo, store the two 32-byte pieces
// at logical byte offsets 32 g (hi part) and 2 C + 32 g (lo part) of row n, 128-byte swizzled.
template <int C>
__device__ __forceinline__ void convert_unit16(unsigned char* stage, uint32_t atom_bytes, int n, int g,
                                               const float* __restrict__ src, size_t cstride, bool inb) {
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = inb ? __ldg(src + (size_t)e * cstride) : 0.f;
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const uint32_t h = bf16x2_rn(v[2 * e + 1], v[2 * e]);
        const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xffff0000u);
        hi[e] = h;
        lo[e] = bf16x2_rn(v[2 * e + 1] - h1, v[2 * e] - h0);
    }
    const uint32_t row_off = (uint32_t)(n >> 3) * 1024u + (uint32_t)(n & 7) * 128u;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t oh = (uint32_t)g * 32u + half * 16u, ol = 2u * C + oh;
        unsigned char* ph = stage + (oh >> 7) * atom_bytes + row_off + ((((oh & 127u) >> 4) ^ (uint32_t)(n & 7)) << 4);
        unsigned char* pl = stage + (ol >> 7) * atom_bytes + row_off + ((((ol & 127u) >> 4) ^ (uint32_t)(n & 7)) << 4);
        *reinterpret_cast<uint4*>(ph) = make_uint4(hi[4 * half], hi[4 * half + 1], hi[4 * half + 2], hi[4 * half + 3]);
        *reinterpret_cast<uint4*>(pl) = make_uint4(lo[4 * half], lo[4 * half + 1], lo[4 * half + 2], lo[4 * half + 3]);
    }
}

// ---- the kernel -------------------------------------------------------------------------------------------------
template <int R, int C>
__global__ void __launch_bounds__(MAX_WARPS * 32, 1)
lc_tc_kernel(const LcParams p, const TcCfg c, const TileDesc* __restrict__ plan) {
    constexpr int W = 2 * R + 2, KW = 2 * R + 1, KK = KW * KW;
    constexpr int ATOMS = (2 * C * 2 + 127) / 128;          // 128-byte atoms per K-major row [hi(C) | lo(C)] of bf16
    constexpr int NKS = C / 16;                              // K = 16 steps per part
    constexpr uint32_t A_ATOM = 128 * 128;
    static_assert(W <= MARGIN, "margin too small");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t b_atom = (uint32_t)c.nmax * 128u;
    const uint32_t a_stage = ATOMS * A_ATOM, b_stage = ATOMS * b_atom;
    unsigned char* a_base = smem;
    unsigned char* b_base = smem + (size_t)c.nsta * a_stage;
    float* ebuf = reinterpret_cast<float*>(b_base + (size_t)c.nstb * b_stage);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ebuf + (size_t)EPI_WARPS * 32 * c.pitch);
    uint64_t* a_full = bars;               // [nsta]
    uint64_t* a_empty = a_full + 2;        // [nsta]
    uint64_t* b_full = a_empty + 2;        // [nstb]
    uint64_t* b_empty = b_full + 4;        // [nstb]
    uint64_t* d_full = b_empty + 4;        // [nacc]
    uint64_t* d_empty = d_full + 4;        // [nacc]
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = p.G;
    const size_t gg = (size_t)G * G;

    if (threadIdx.x == 0) {
        for (int s = 0; s < c.nsta; ++s) { mbar_init(&a_full[s], c.ncw); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < c.nstb; ++s) { mbar_init(&b_full[s], c.ncw); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < c.nacc; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (warp >= FIRST_CONV_WARP) {
        // ================= converter warps =================
        const int ct = threadIdx.x - FIRST_CONV_WARP * 32, nct = c.ncw * 32;
        uint32_t q = 0, tt = 0;
        for (int tile = blockIdx.x; tile < c.ntiles; tile += gridDim.x) {
            const TileDesc d = plan[tile];
            if (d.flags) continue;
            int t = tile;
            const int tx = t % c.tiles_x; t /= c.tiles_x;
            const int ty = t % c.tiles_y;
            const int b = t / c.tiles_y;
            {   // A: the tile's 128 lattice points x C channels of f0
                const uint32_t as = tt % c.nsta;
                mbar_wait(&a_empty[as], ((tt / c.nsta) & 1) ^ 1);
                unsigned char* stage = a_base + (size_t)as * a_stage;
                for (int u = ct; u < 128 * NKS; u += nct) {
                    const int m = u & 127, g = u >> 7;
                    const int gy = ty * c.TRy + m / c.TCx, gx = tx * c.TCx + m % c.TCx;
                    const bool inb = gy < G && gx < G;
                    const float* src = p.f0 + ((size_t)b * C + g * 16) * gg + (size_t)min(gy, G - 1) * G + min(gx, G - 1);
                    convert_unit16<C>(stage, A_ATOM, m, g, src, gg, inb);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[as]);
            }
            const int rpc = c.nmax / d.bw;
            const int nchunks = (d.nrows + rpc - 1) / rpc;
            const size_t plane = (size_t)p.Hs * p.pitch;
            for (int ch = 0; ch < nchunks; ++ch, ++q) {
                const uint32_t s = q % c.nstb;
                mbar_wait(&b_empty[s], ((q / c.nstb) & 1) ^ 1);
                unsigned char* stage = b_base + (size_t)s * b_stage;
                const int rows_here = min(rpc, d.nrows - ch * rpc);
                const int N = rows_here * d.bw;
                const int ybase = d.y0 + ch * rpc;
                for (int u = ct; u < N * NKS; u += nct) {
                    const int g = u / N, n = u - g * N;
                    const int ry = n / d.bw, xx = n - ry * d.bw;
                    const int y = ybase + ry, x = d.x0 + xx;
                    const bool inb = x < p.Ws;
                    const float* src = p.f1 + ((size_t)b * C + g * 16) * plane + (size_t)y * p.pitch + min(x, p.Ws - 1);
                    convert_unit16<C>(stage, b_atom, n, g, src, plane, inb);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&b_full[s]);
            }
            ++tt;
        }
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t q = 0, tt = 0;
            for (int tile = blockIdx.x; tile < c.ntiles; tile += gridDim.x) {
                const TileDesc d = plan[tile];
                if (d.flags) continue;
                const uint32_t as = tt % c.nsta;
                mbar_wait(&a_full[as], (tt / c.nsta) & 1);
                const uint32_t a_addr = smem_u32(a_base + (size_t)as * a_stage);
                const int rpc = c.nmax / d.bw;
                const int nchunks = (d.nrows + rpc - 1) / rpc;
                const uint32_t idesc = idesc_bf16(rpc * d.bw);
                for (int ch = 0; ch < nchunks; ++ch, ++q) {
                    const uint32_t s = q % c.nstb, acc = q % c.nacc;
                    mbar_wait(&b_full[s], (q / c.nstb) & 1);
                    mbar_wait(&d_empty[acc], ((q / c.nacc) & 1) ^ 1);
                    fence_after_sync();
                    const uint32_t b_addr = smem_u32(b_base + (size_t)s * b_stage);
                    const uint32_t dt = tmem_base + acc * (uint32_t)c.nmax;
                    uint32_t accum = 0;
#pragma unroll
                    for (int combo = 0; combo < 3; ++combo) {
                        const int pa = combo == 2 ? 1 : 0, pb = combo == 1 ? 1 : 0;   // lo*hi, hi*lo, then hi*hi last
#pragma unroll
                        for (int ks = 0; ks < NKS; ++ks) {
                            const uint32_t oa = (uint32_t)(pa * C * 2 + ks * 32), ob = (uint32_t)(pb * C * 2 + ks * 32);
                            const uint64_t ad = smem_desc_k128(a_addr + (oa >> 7) * A_ATOM + (oa & 127u));
                            const uint64_t bd = smem_desc_k128(b_addr + (ob >> 7) * b_atom + (ob & 127u));
                            mma_bf16(dt, ad, bd, idesc, accum);
                            accum = 1;
                        }
                    }
                    mma_commit(&b_empty[s]);
                    mma_commit(&d_full[acc]);
                    if (ch + 1 == nchunks) mma_commit(&a_empty[as]);
                }
                ++tt;
            }
        }
    } else {
        // ================= epilogue warps: TMEM lane = lattice point =================
        float* buf = ebuf + (size_t)(warp * 32 + lane) * c.pitch;
#pragma unroll
        for (int i = 0; i < MARGIN / 4; ++i) *reinterpret_cast<float4*>(buf + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t q = 0;
        for (int tile = blockIdx.x; tile < c.ntiles; tile += gridDim.x) {
            const TileDesc d = plan[tile];
            int t = tile;
            const int tx = t % c.tiles_x; t /= c.tiles_x;
            const int ty = t % c.tiles_y;
            const int b = t / c.tiles_y;
            const int m = warp * 32 + lane;
            const int gy = ty * c.TRy + m / c.TCx, gx = tx * c.TCx + m % c.TCx;
            const bool valid = gy < G && gx < G;
            float* outp = p.out + ((size_t)b * p.k_total + p.k_offset) * gg + (size_t)gy * G + gx;
            if (d.flags & TF_GATHER) {
                if (valid)
                    for (int k = 0; k < KK; ++k) st_stream(outp + (size_t)k * gg, lc_generic_point(p, b, k, gy, gx));
                continue;
            }
            bool live = false;
            int xb = 0, yb = 0;
            float wx1 = 0.f, wy0 = 0.f, wy1 = 0.f;
            if (valid) {
                const float* fl = p.flow + (size_t)b * 2 * gg + (size_t)gy * G + gx;
                const float sx = unnormalize(__ldg(fl), p.Ws), sy = unnormalize(__ldg(fl + gg), p.Hs);
                if (fabsf(sx) < 1e6f && fabsf(sy) < 1e6f) {
                    const float x0f = floorf(sx), y0f = floorf(sy);
                    xb = (int)x0f - R; yb = (int)y0f - R;
                    const float ty_ = sy - y0f;
                    wx1 = sx - x0f;
                    wy0 = (1.f - ty_) * p.inv_sqrt_c;
                    wy1 = ty_ * p.inv_sqrt_c;
                    live = !(xb >= p.Ws || xb + W <= 0 || yb >= p.Hs || yb + W <= 0);
                }
                if (!live)
                    for (int k = 0; k < KK; ++k) st_stream(outp + (size_t)k * gg, 0.f);
            }
            if (d.flags & TF_EMPTY) continue;
#pragma unroll
            for (int i = 0; i < MARGIN / 4; ++i)
                *reinterpret_cast<float4*>(buf + MARGIN + d.bw + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
            const int rpc = c.nmax / d.bw;
            const float* rowp = buf + MARGIN + (live ? xb - d.x0 : 0);
            const float a1 = wx1, a0 = 1.f - wx1;
            float hprev[KW];
#pragma unroll
            for (int i = 0; i < KW; ++i) hprev[i] = 0.f;
            int ri = 0, ry = 0;
            uint32_t acc = 0;
            for (int y = d.ylo; y < d.yhi; ++y) {
                const int j = y - yb;
                const bool act = live && (unsigned)j < (unsigned)W;
                const bool in_img = (unsigned)y < (unsigned)p.Hs;
                if (in_img && ry == 0) {
                    acc = q % c.nacc;
                    mbar_wait(&d_full[acc], (q / c.nacc) & 1);
                    fence_after_sync();
                }
                if (__any_sync(0xffffffffu, act)) {
                    if (in_img) {
                        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * (uint32_t)c.nmax + (uint32_t)(ry * d.bw);
                        float* dst = buf + MARGIN;
                        int c0 = 0;
                        for (; c0 + 32 <= d.bw; c0 += 32) {
                            uint32_t r[32];
                            tmem_ld32(taddr + c0, r);
                            tmem_ld_wait();
#pragma unroll
                            for (int v = 0; v < 8; ++v)
                                *reinterpret_cast<uint4*>(dst + c0 + 4 * v) = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
                        }
                        if (c0 < d.bw) {
                            uint32_t r[16];
                            tmem_ld16(taddr + c0, r);
                            tmem_ld_wait();
#pragma unroll
                            for (int v = 0; v < 4; ++v)
                                *reinterpret_cast<uint4*>(dst + c0 + 4 * v) = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
                        }
                    }
                    if (act) {
                        float D[W];
#pragma unroll
                        for (int i = 0; i < W; ++i) D[i] = in_img ? rowp[i] : 0.f;
                        float* op = outp + (size_t)(j - 1) * KW * gg;
#pragma unroll
                        for (int i = 0; i < KW; ++i) {
                            const float h = a0 * D[i] + a1 * D[i + 1];
                            if (j >= 1) st_stream(op + (size_t)i * gg, wy0 * hprev[i] + wy1 * h);
                            hprev[i] = h;
                        }
                    }
                }
                if (in_img) {
                    ++ri;
                    if (++ry == rpc || ri == d.nrows) {
                        ry = 0;
                        fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&d_empty[acc]);
                        ++q;
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int R, int C>
static int launch_tc(const LcParams& p, cudaStream_t st, void* workspace, size_t ws_bytes, int tune) {
    constexpr int W = 2 * R + 2;
    constexpr int ATOMS = (2 * C * 2 + 127) / 128;
    TcCfg c;
    const int G = p.G;
    // tune: bits 0-7 TCx override, bits 8-11 converter warps, bits 12-15 nmax / 64, bits 16-19 grid CTAs per SM x 148 divisor (unused)
    c.TCx = (G % 32 == 0) ? 32 : (G % 16 == 0 ? 16 : (G % 8 == 0 ? 8 : 32));
    if (tune & 0xff) c.TCx = tune & 0xff;
    if (c.TCx != 8 && c.TCx != 16 && c.TCx != 32 && c.TCx != 64 && c.TCx != 128) return GFB_EINVAL;
    c.TRy = 128 / c.TCx;
    c.tiles_x = (G + c.TCx - 1) / c.TCx;
    c.tiles_y = (G + c.TRy - 1) / c.TRy;
    c.ntiles = p.B * c.tiles_x * c.tiles_y;
    c.ncw = (tune >> 8) & 15 ? (tune >> 8) & 15 : 8;
    if (c.ncw > MAX_WARPS - FIRST_CONV_WARP) return GFB_EINVAL;
    c.nmax = (tune >> 12) & 15 ? ((tune >> 12) & 15) * 64 : (C == 64 ? 128 : 256);
    if (c.nmax != 128 && c.nmax != 256) return GFB_EINVAL;
    c.nacc = 512 / c.nmax;
    const float s = (float)p.Ws / (float)G;
    int bwmax = (((int)ceilf((float)c.TCx * s * 1.6f) + W + 16) + 15) & ~15;
    bwmax = min(bwmax, (p.Ws + 15) & ~15);
    c.bwmax = min(bwmax, c.nmax);
    c.pitch = MARGIN + c.bwmax + MARGIN;
    c.pitch += (4 - c.pitch % 32 + 32) % 32;
    const size_t a_stage = (size_t)ATOMS * 128 * 128, b_stage = (size_t)ATOMS * c.nmax * 128;
    const size_t epi = (size_t)EPI_WARPS * 32 * c.pitch * sizeof(float);
    const size_t fixed = 1024 + epi + 24 * sizeof(uint64_t);
    const size_t budget = 227 * 1024;
    c.nsta = 2;
    if (fixed + 2 * a_stage + 2 * b_stage > budget) c.nsta = 1;
    if (fixed + c.nsta * a_stage + 2 * b_stage > budget) return GFB_EUNSUPPORTED;
    c.nstb = (int)min((size_t)4, (budget - fixed - c.nsta * a_stage) / b_stage);
    const size_t smem = fixed + c.nsta * a_stage + c.nstb * b_stage;

    if (ws_bytes < (size_t)c.ntiles * sizeof(TileDesc) || !workspace) return GFB_EWORKSPACE;
    TileDesc* plan = reinterpret_cast<TileDesc*>(workspace);
    lc_plan_kernel<<<(c.ntiles + 3) / 4, 128, 0, st>>>(p, c, plan);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;

    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto kern = lc_tc_kernel<R, C>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<min(c.ntiles, sms), (FIRST_CONV_WARP + c.ncw) * 32, smem, st>>>(p, c, plan);
    GFB_LAUNCH_RESULT();
}

}  // namespace lctc
}  // namespace gfb

using namespace gfb;

extern "C" size_t gfb_local_corr_tc_workspace_bytes(int B, int G) {
    if (B <= 0 || G <= 0) return 0;
    size_t tiles = 0;
    for (int tcx = 8; tcx <= 128; tcx *= 2) {
        const size_t t = (size_t)((G + tcx - 1) / tcx) * ((G + 128 / tcx - 1) / (128 / tcx));
        tiles = t > tiles ? t : tiles;
    }
    return (size_t)B * tiles * sizeof(lctc::TileDesc);
}

extern "C" int gfb_local_corr_tc_f32(const float* f0, const float* f1, const float* flow, float* out,
                                     int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                                     int k_total, int k_offset, int tune,
                                     void* workspace, size_t workspace_bytes, gfb_stream_t stream) {
    GFB_CHECK_ARG(f0 && f1 && flow && out);
    GFB_CHECK_ARG(B > 0 && C > 0 && Hs > 0 && Ws > 0 && G > 0 && r >= 0);
    GFB_CHECK_ARG(f1_pitch == 0 || f1_pitch >= Ws);
    const int kk = (2 * r + 1) * (2 * r + 1);
    GFB_CHECK_ARG(k_offset >= 0 && k_offset + kk <= k_total);
    GFB_CHECK_ARG(tune >= 0);
    LcParams p;
    p.f0 = f0; p.f1 = f1; p.flow = flow; p.out = out;
    p.B = B; p.C = C; p.Hs = Hs; p.Ws = Ws; p.G = G; p.r = r;
    p.pitch = f1_pitch ? f1_pitch : Ws;
    p.k_total = k_total; p.k_offset = k_offset;
    p.sample_mode = 0; p.padding_mode = 0;
    p.ox0 = (float)(-2.0 * r / Ws); p.ox1 = (float)(2.0 * r / Ws);
    p.oy0 = (float)(-2.0 * r / Hs); p.oy1 = (float)(2.0 * r / Hs);
    p.inv_sqrt_c = (float)(1.0 / sqrt((double)C));
    p.debug = 0;
    cudaStream_t st = gfb_cu(stream);
#define GFB_TC_CASE(RR, CC) if (r == RR && C == CC) return lctc::launch_tc<RR, CC>(p, st, workspace, workspace_bytes, tune);
    GFB_TC_CASE(2, 16) GFB_TC_CASE(4, 32) GFB_TC_CASE(6, 64) GFB_TC_CASE(7, 64)
#undef GFB_TC_CASE
    return GFB_EUNSUPPORTED;
}
