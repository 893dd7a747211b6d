// K1, second generation (reference: utils/local_correlation.py:4-72, call site model/network.py:553-554).
//
//   corr[b,k,gy,gx] = (1/sqrt C) sum_c f0[b,c,gy,gx] * bilinear(f1[b,c], flow[b,:,gy,gx] + off_k)
//
// The window offsets are whole pixels, so corr = bilerp(D) with D[j,i] = sum_c f0[c] * f1[c, y0-r+j, x0-r+i] over a
// (2r+2)^2 integer patch.  Two kernels, both fed by TMA and both with zero per-sample address arithmetic:
//
//  lc_pt_kernel   (small C, small r: the 1/2-resolution scales)  one lattice point per thread, the whole patch of D in
//                 registers.  A CTA owns TX x TY points; the bounding box of their windows is one TMA box per channel
//                 group (out-of-image pixels zero-filled by the TMA unit = padding_mode "zeros"); the inner loop is
//                 LDS.64 + FFMA2 with immediate offsets only.
//
//  lc_tc2_kernel  (C >= 32) D as a banded GEMM on tcgen05.  A pre-pass (lc_prep_kernel) rewrites f0 and f1 ONCE as
//                 position-major rows [bf16 hi(C) | bf16 lo(C)] (a workspace that stays L2-resident per batch group),
//                 so that the main kernel needs no converter warps: one thread issues TMA loads (128-byte swizzle,
//                 zero fill), one thread issues tcgen05.mma (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM), four
//                 warps run the epilogue (TMEM lane = lattice point; one tcgen05.ld.x32 per image row and warp, a
//                 lane-private transposed staging column, bilerp, streaming stores).  Two CTAs per SM (256 TMEM columns
//                 each) overlap each other's load / MMA / epilogue phases.
//
// Points whose windows do not fit the staged box (wild flows) take the exact per-sample gather (lc_generic_point).
#include "common.cuh"
#include "lc_common.cuh"
#include <limits.h>

namespace gfb {
namespace lcv2 {

__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// debug counters (gfb_debug_local_corr_v2_counters): [0] lc_pt points on the global-memory path, [1] lc_tc2 points on
// the gather path, [2] lc_tc2 gather tiles
__device__ unsigned long long g_v2_stats[8];   // [4..7] (debug bit 0): epilogue warp 0 clocks waiting / TMEM pull / rows, chunks

// =====================================================================================================================
// lc_pt_kernel: one point per thread, the (2r+2)^2 patch in registers
// =====================================================================================================================
// A point whose window leaves the staged box: the same factored sum straight from global memory (rare; exact zero padding).
template <int R>
__device__ __noinline__ void lc_point_global(const LcParams& p, int b, int gy, int gx, const PointGeom pg, float* outp) {
    constexpr int W = 2 * R + 2, KW = 2 * R + 1;
    const size_t gg = (size_t)p.G * p.G, plane = (size_t)p.Hs * p.pitch;
    const float* f0 = p.f0 + ((size_t)b * p.f0_ctot + p.c0) * gg + (size_t)gy * p.G + gx;
    const float* f1 = p.f1 + ((size_t)b * p.Ctot + p.c0) * plane;
    const float a1 = pg.fx, a0 = 1.f - pg.fx;
    const float wy1 = pg.fy * p.inv_sqrt_c, wy0 = (1.f - pg.fy) * p.inv_sqrt_c;
    float hprev[KW];
#pragma unroll
    for (int i = 0; i < KW; ++i) hprev[i] = 0.f;
    for (int j = 0; j < W; ++j) {
        float D[W];
#pragma unroll
        for (int i = 0; i < W; ++i) D[i] = 0.f;
        const int y = pg.yb + j;
        if ((unsigned)y < (unsigned)p.Hs) {
            for (int c = 0; c < p.C; ++c) {
                const float f = __ldg(f0 + (size_t)c * gg);
                const float* row = f1 + (size_t)c * plane + (size_t)y * p.pitch;
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    const int x = pg.xb + i;
                    if ((unsigned)x < (unsigned)p.Ws) D[i] = fmaf(__ldg(row + x), f, D[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < KW; ++i) {
            const float h = a0 * D[i] + a1 * D[i + 1];
            if (j >= 1) st_stream(outp + (size_t)((j - 1) * KW + i) * gg, wy0 * hprev[i] + wy1 * h);
            hprev[i] = h;
        }
    }
}

// Persistent CTA of TX x TY points.  f1 streams through a two-stage ring of TMA boxes (CG channels x BH x BW pixels
// around the bounding box of the tile's windows; pixels outside the image are zero-filled by the TMA unit); the ring
// runs across tile boundaries, so the next tile's boxes are in flight while this tile finishes.  f0 of a tile arrives
// by TMA as well.  Every thread keeps the W x (W+2) patch of its point in registers (LDS.64 + FFMA2, immediate offsets).
template <int R, int C, int TX, int TY, int BW, int BH, int CG, int NBUF>
__global__ void __launch_bounds__(TX * TY, 3)
lc_pt_kernel(const LcParams p, const int ntiles, const __grid_constant__ CUtensorMap tmap1,
             const __grid_constant__ CUtensorMap tmap0) {
    constexpr int W = 2 * R + 2, KW = 2 * R + 1, KK = KW * KW, NT = TX * TY;
    constexpr int WA = W + 2;                 // even-aligned accumulator width (LDS.64)
    constexpr int NG = C / CG;
    constexpr int PLANE = BH * BW, STAGE = CG * PLANE, F0SZ = C * NT;
    static_assert(BW % 4 == 0 && C % CG == 0 && NT % 32 == 0 && NG >= NBUF, "shape");
    static_assert((STAGE * 4) % 128 == 0 && (F0SZ * 4) % 128 == 0, "TMA destinations");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);                          // [NBUF][CG][BH][BW]
    float* f0s = ring + NBUF * STAGE;                                          // [2][C][NT]
    uint64_t* full = reinterpret_cast<uint64_t*>(f0s + 2 * F0SZ);              // [NBUF]
    int* s_min = reinterpret_cast<int*>(full + NBUF);                          // [2][2]

    const int tid = threadIdx.x, lane = tid & 31;
    const int G = p.G;
    const size_t gg = (size_t)G * G;
    const int tiles_x = (G + TX - 1) / TX, tiles_y = (G + TY - 1) / TY;
    const int lx = tid % TX, ly = tid / TX;

    auto geom_of = [&](int tile) {
        if (tile >= ntiles) { PointGeom z; z.xb = 0; z.yb = 0; z.fx = 0.f; z.fy = 0.f; z.live = false; return z; }
        int t = tile;
        const int tx = t % tiles_x; t /= tiles_x;
        const int ty = t % tiles_y;
        const int gx = tx * TX + lx, gy = ty * TY + ly;
        return point_geom(p, t / tiles_y, gy, gx, gx < G && gy < G, R);
    };
    auto reduce_bbox = [&](const PointGeom& g, int slot) {
        const int mx = warp_min(g.live ? g.xb : INT_MAX), my = warp_min(g.live ? g.yb : INT_MAX);
        if (lane == 0 && mx != INT_MAX) { atomicMin(&s_min[2 * slot], mx); atomicMin(&s_min[2 * slot + 1], my); }
    };
    uint32_t issued = 0;                      // thread 0: boxes issued so far
    auto issue = [&](int tile, int g, int slot) {
        int t = tile;
        const int tx = t % tiles_x; t /= tiles_x;
        const int ty = t % tiles_y;
        const int b = t / tiles_y;
        const int mx = s_min[2 * slot], my = s_min[2 * slot + 1];
        const int X0 = mx == INT_MAX ? 0 : (mx & ~3), Y0 = mx == INT_MAX ? 0 : my;   // 16-byte aligned innermost coordinate
        const uint32_t buf = issued % NBUF;
        mbar_expect_tx(&full[buf], (uint32_t)((STAGE + (g == 0 ? F0SZ : 0)) * sizeof(float)));
        tma_load_3d(ring + buf * STAGE, &tmap1, &full[buf], X0, Y0, b * C + g * CG);
        if (g == 0) tma_load_3d(f0s + slot * F0SZ, &tmap0, &full[buf], tx * TX, ty * TY, b * p.f0_ctot);
        ++issued;
    };

    int tile = blockIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) mbar_init(&full[i], 1);
        for (int i = 0; i < 4; ++i) s_min[i] = INT_MAX;
        mbar_fence_init();
    }
    __syncthreads();
    PointGeom g_cur = geom_of(tile);
    reduce_bbox(g_cur, 0);
    __syncthreads();
    if (tid == 0)
        for (int g = 0; g < NBUF; ++g) issue(tile, g, 0);
    PointGeom g_nxt = geom_of(tile + gridDim.x);
    uint32_t gl = 0;                          // boxes consumed so far

    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int nxt = tile + gridDim.x;
        const int slot = it & 1;
        if (tid == 0) { s_min[2 * (slot ^ 1)] = INT_MAX; s_min[2 * (slot ^ 1) + 1] = INT_MAX; }
        __syncthreads();
        reduce_bbox(g_nxt, slot ^ 1);         // read by thread 0 after the next __syncthreads at the earliest
        const PointGeom g_nn = geom_of(nxt + gridDim.x);

        int t = tile;
        const int tx = t % tiles_x; t /= tiles_x;
        const int ty = t % tiles_y;
        const int b = t / tiles_y;
        const int gx = tx * TX + lx, gy = ty * TY + ly;
        const bool valid = gx < G && gy < G;
        const int mx = s_min[2 * slot], my = s_min[2 * slot + 1];
        const int X0 = mx == INT_MAX ? 0 : (mx & ~3), Y0 = mx == INT_MAX ? 0 : my;
        const int ox = g_cur.xb - X0, oy = g_cur.yb - Y0;
        const int oxa = ox & ~1;
        const bool fit = g_cur.live && oxa + WA <= BW && oy + W <= BH;

        unsigned long long acc[W][WA / 2];
#pragma unroll
        for (int j = 0; j < W; ++j)
#pragma unroll
            for (int h = 0; h < WA / 2; ++h) acc[j][h] = 0ull;
        const float* f0t = f0s + slot * F0SZ + tid;
#pragma unroll 1
        for (int g = 0; g < NG; ++g, ++gl) {
            const uint32_t buf = gl % NBUF;
            mbar_wait(&full[buf], (gl / NBUF) & 1);
            if (fit) {
                const float* base = ring + buf * STAGE + oy * BW + oxa;
#pragma unroll
                for (int cc = 0; cc < CG; ++cc) {
                    const float fv = f0t[(g * CG + cc) * NT];
                    const unsigned long long f = pack2(fv, fv);
#pragma unroll
                    for (int j = 0; j < W; ++j)
#pragma unroll
                        for (int h = 0; h < WA / 2; ++h) {
                            const unsigned long long v = *reinterpret_cast<const unsigned long long*>(base + cc * PLANE + j * BW + 2 * h);
                            ffma2(acc[j][h], v, f);
                        }
                }
            }
            __syncthreads();                  // every thread is done with ring[buf] (and, at g = NG-1, with f0s[slot])
            if (tid == 0) {
                if (g + NBUF < NG) issue(tile, g + NBUF, slot);
                else if (nxt < ntiles) issue(nxt, g + NBUF - NG, slot ^ 1);
            }
        }

        float* outp = p.out + ((size_t)b * p.k_total + p.k_offset) * gg + (size_t)gy * G + gx;
        if (fit) {
            const bool odd = ox & 1;
            const float a1 = g_cur.fx, a0 = 1.f - g_cur.fx;
            const float wy1 = g_cur.fy * p.inv_sqrt_c, wy0 = (1.f - g_cur.fy) * p.inv_sqrt_c;
            float hprev[KW];
#pragma unroll
            for (int j = 0; j < W; ++j) {
                float a[WA];
#pragma unroll
                for (int h = 0; h < WA / 2; ++h) {
                    const float2 v = unpack2(acc[j][h]);
                    a[2 * h] = v.x; a[2 * h + 1] = v.y;
                }
                float D[W];
#pragma unroll
                for (int i = 0; i < W; ++i) D[i] = odd ? a[i + 1] : a[i];
#pragma unroll
                for (int i = 0; i < KW; ++i) {
                    const float h = a0 * D[i] + a1 * D[i + 1];
                    if (j >= 1) st_stream(outp + (size_t)((j - 1) * KW + i) * gg, wy0 * hprev[i] + wy1 * h);
                    hprev[i] = h;
                }
            }
        } else if (valid) {
            if (!g_cur.live) {
                for (int k = 0; k < KK; ++k) st_stream(outp + (size_t)k * gg, 0.f);
            } else {
                atomicAdd(&g_v2_stats[0], 1ull);
                lc_point_global<R>(p, b, gy, gx, g_cur, outp);
            }
        }
        g_cur = g_nxt;
        g_nxt = g_nn;
    }
}

template <int R, int C, int TX, int TY, int BW, int BH, int CG, int NBUF>
static int launch_pt(const LcParams& p, cudaStream_t st) {
    constexpr size_t smem = (size_t)(NBUF * CG * BH * BW + 2 * C * TX * TY) * sizeof(float) + NBUF * sizeof(uint64_t) + 16;
    const int G = p.G;
    if (G % 4 != 0) return GFB_EUNSUPPORTED;              // 16-byte global strides of the f0 tensor map
    CUtensorMap tmap1, tmap0;
    {
        uint64_t dims[3] = {(uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)p.B * p.C};
        uint64_t strides[2] = {(uint64_t)p.pitch * 4, (uint64_t)p.Hs * p.pitch * 4};
        uint32_t box[3] = {(uint32_t)BW, (uint32_t)BH, (uint32_t)CG};
        int rc = gfb_encode_tmap_f32(&tmap1, p.f1, 3, dims, strides, box, 0);
        if (rc != GFB_OK) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)G, (uint64_t)G, (uint64_t)p.B * p.f0_ctot};
        uint64_t strides[2] = {(uint64_t)G * 4, (uint64_t)G * G * 4};
        uint32_t box[3] = {(uint32_t)TX, (uint32_t)TY, (uint32_t)C};
        int rc = gfb_encode_tmap_f32(&tmap0, p.f0, 3, dims, strides, box, 0);
        if (rc != GFB_OK) return rc;
    }
    auto kern = lc_pt_kernel<R, C, TX, TY, BW, BH, CG, NBUF>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long tiles = (long long)p.B * ((G + TY - 1) / TY) * ((G + TX - 1) / TX);
    if (tiles > 0x7fffffffLL) return GFB_EUNSUPPORTED;
    int dev = 0, sms = 148, per_sm = 1;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TX * TY, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int grid = (int)min(tiles, (long long)sms * per_sm);
    kern<<<grid, TX * TY, smem, st>>>(p, (int)tiles, tmap1, tmap0);
    GFB_LAUNCH_RESULT();
}

// =====================================================================================================================
// lc_rot_kernel: one point per thread, conflict-free window reads by rotating the visiting order (r = 2)
// =====================================================================================================================
// lc_pt_kernel sits on the shared-memory roof (ncu: 3.6 wavefronts per LDS.64 where 2 are ideal, plus two padding columns
// per window row): neighbouring lattice points read windows that start ~1.75 px apart and on different image rows, so the
// lanes of one load collide in the banks.  Here every lane visits the W x W cells of its window in a rotated order:
// instruction (jj, ii) reads the one cell of the lane's window whose box row is == jj and whose box column is == ii
// (mod W).  All lanes of a load then touch cells of ONE residue class; two different cells of a class are multiples of W
// apart in x and / or y, and with a box pitch of 8 or 24 (mod 32) words they fall into different banks across the
// footprint of a warp of 16 x 2 lattice points (equal cells broadcast).  Measured on the bench flows by simulation
// (tools/sim/sim5.py): 1.003 wavefronts per LDS.32, against 1.88 per half-warp LDS.64 before -- exactly W*W loads per
// point and channel, no padding columns.  The accumulators are un-rotated once per tile by a select network.
//
// The rest of the kernel is built around the two other limits the first kernel hit: (1) the L2 -> shared-memory fill:
// the TMA box of a tile is the narrowest / shortest of 3 x 5 shapes that holds the tile's windows (16 x 16 points per
// tile: 8.8 staged pixels per point in the median instead of 16); (2) barriers: a producer warp (geometry of the next
// tile, tile descriptor, TMA issue) and eight consumer warps meet only through full / empty mbarriers, there is no
// CTA-wide barrier in the main loop.
#ifndef ROT_MINB
#define ROT_MINB 2
#endif
namespace rot {
constexpr int R = 2, W = 6, KW = 5, KK = 25;
constexpr int TX = 16, TY = 16, NT = TX * TY, NCW = NT / 32;     // tile, consumer warps
constexpr int NBW = 3, NBH = 5;
__host__ __device__ constexpr int box_w(int i) { return 40 + 16 * i; }       // 40, 56, 72: all 8 or 24 (mod 32)
__host__ __device__ constexpr int box_h(int i) { return 24 + 8 * i; }        // 24 .. 56
constexpr int NTS = 3;                                                         // tile slots (bounding box + descriptor)
struct Maps { CUtensorMap m[NBW][NBH]; };
struct TileInfo { int P, bh, X0, Y0; };                                       // P = 0: no window of the tile meets the image
struct BBox { int mnx, mny, mxx, mxy; };
}  // namespace rot

// rotate a 6-vector: out[j] = v[(j + n) % 6], n in [0, 6)
__device__ __forceinline__ void rot6(float (&v)[6], int n) {
    float t[6];
#pragma unroll
    for (int s = 1; s <= 4; s <<= 1) {
        const bool on = (n & s) != 0;
#pragma unroll
        for (int j = 0; j < 6; ++j) t[j] = on ? v[(j + s) % 6] : v[j];
#pragma unroll
        for (int j = 0; j < 6; ++j) v[j] = t[j];
    }
}

// Roles.  Consumer warps (thread = lattice point, warp = 16 x 2 points): at the start of tile k they turn the flow of
// their point of tile k+1 (loaded one tile earlier) into window origin + fractions, reduce the tile's bounding box into
// shared memory and signal `gready`; then they load the flow of tile k+2, wait for the descriptor of tile k and run its
// channels.  Producer warp (one lane): waits for `gready` of a tile, picks the TMA box, publishes the descriptor
// (`tfull`) and streams the channels through the ring (each stage = one channel of the f1 box + the 16 x 16 f0 tile).
template <int C, int NBUF, int MBW, int MBH>
__global__ void __launch_bounds__(rot::NT + 32, ROT_MINB)
lc_rot_kernel(const LcParams p, const int ntiles, const __grid_constant__ rot::Maps maps,
              const __grid_constant__ CUtensorMap tmap0) {
    using namespace rot;
    constexpr int BOXMAX = box_w(MBW) * box_h(MBH);                            // floats of the largest box
    constexpr int SLOT = BOXMAX + NT;                                          // ring slot: one channel of f1 box + f0 tile
    static_assert((SLOT * 4) % 128 == 0 && (BOXMAX * 4) % 128 == 0 && C % 2 == 0 && NBUF % 2 == 0, "ring layout");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);                          // [NBUF][SLOT]: f1 box, then f0[NT]
    TileInfo* info = reinterpret_cast<TileInfo*>(ring + NBUF * SLOT);          // [NTS]
    BBox* bbox = reinterpret_cast<BBox*>(info + NTS);                          // [NTS]
    uint64_t* full = reinterpret_cast<uint64_t*>(bbox + NTS);                  // [NBUF]
    uint64_t* empty = full + NBUF;                                             // [NBUF]
    uint64_t* gready = empty + NBUF;                                           // [NTS] bounding box of the tile complete
    uint64_t* tfull = gready + NTS;                                            // [NTS] descriptor published

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = p.G;
    const size_t gg = (size_t)G * G;
    const int tiles_x = (G + TX - 1) / TX, tiles_y = (G + TY - 1) / TY;

    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NCW); }
        for (int i = 0; i < NTS; ++i) {
            mbar_init(&gready[i], NCW); mbar_init(&tfull[i], 1);
            bbox[i].mnx = INT_MAX; bbox[i].mny = INT_MAX; bbox[i].mxx = INT_MIN; bbox[i].mxy = INT_MIN;
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NCW) {
        // ================= producer =================
        if (lane == 0) {
            uint32_t q = 0;                                   // ring stages issued
            int n = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n) {
                const int slot = n % NTS;
                mbar_wait(&gready[slot], (n / NTS) & 1);
                const BBox bb = bbox[slot];
                bbox[slot].mnx = INT_MAX; bbox[slot].mny = INT_MAX; bbox[slot].mxx = INT_MIN; bbox[slot].mxy = INT_MIN;
                const bool none = bb.mnx == INT_MAX;
                const int X0 = none ? 0 : (bb.mnx & ~3), Y0 = none ? 0 : bb.mny;    // 16-byte aligned innermost coordinate
                const int ex = none ? 0 : bb.mxx + W - X0, ey = none ? 0 : bb.mxy + W - Y0;
                const int bwi = min(max((ex - box_w(0) + 15) / 16, 0), MBW);
                const int bhi = min(max((ey - box_h(0) + 7) / 8, 0), MBH);
                TileInfo ti;
                ti.P = none ? 0 : box_w(bwi); ti.bh = box_h(bhi); ti.X0 = X0; ti.Y0 = Y0;
                info[slot] = ti;
                mbar_arrive(&tfull[slot]);
                if (none) continue;
                int t = tile;
                const int tx = t % tiles_x; t /= tiles_x;
                const int ty = t % tiles_y;
                const int b = t / tiles_y;
                const CUtensorMap* tm = &maps.m[bwi][bhi];
                const uint32_t bytes = (uint32_t)((box_w(bwi) * box_h(bhi) + NT) * sizeof(float));
                for (int c = 0; c < C; c += 2, q += 2) {         // one full / empty pair per two channels (slots s, s + 1)
                    const uint32_t s = q % NBUF;
                    mbar_wait(&empty[s], ((q / NBUF) & 1) ^ 1);
                    if (p.debug & 2) { mbar_arrive(&full[s]); continue; }
                    mbar_expect_tx(&full[s], 2 * bytes);
                    tma_load_3d(ring + s * SLOT, tm, &full[s], X0, Y0, b * C + c);
                    tma_load_3d(ring + s * SLOT + BOXMAX, &tmap0, &full[s], tx * TX, ty * TY, b * p.f0_ctot + c);
                    tma_load_3d(ring + (s + 1) * SLOT, tm, &full[s], X0, Y0, b * C + c + 1);
                    tma_load_3d(ring + (s + 1) * SLOT + BOXMAX, &tmap0, &full[s], tx * TX, ty * TY, b * p.f0_ctot + c + 1);
                }
            }
        }
        return;
    }

    // ================= consumer warps =================
    // tile coordinates (b, ty, tx) advance by gridDim.x tiles per step: mixed-radix increments instead of divisions
    struct TileCoord { int tx, ty, b; };
    auto coord_of = [&](int tile) {
        TileCoord c;
        int t = tile;
        c.tx = t % tiles_x; t /= tiles_x;
        c.ty = t % tiles_y;
        c.b = t / tiles_y;
        return c;
    };
    const TileCoord step = coord_of((int)gridDim.x);
    auto advance = [&](TileCoord& c) {
        c.tx += step.tx; c.ty += step.ty; c.b += step.b;
        if (c.tx >= tiles_x) { c.tx -= tiles_x; ++c.ty; }
        if (c.ty >= tiles_y) { c.ty -= tiles_y; ++c.b; }
    };
    const int px = tid % TX, py = tid / TX;
    // raw flow of this thread's point of a tile (NaN = no point); volatile so that the loads stay where they are issued
    auto load_flow = [&](int tile, const TileCoord& c, float& rx, float& ry) {
        rx = ry = __int_as_float(0x7fc00000);
        const int gx = c.tx * TX + px, gy = c.ty * TY + py;
        if (tile < ntiles && gx < G && gy < G) {
            const float* fl = p.flow + (size_t)c.b * 2 * gg + (size_t)gy * G + gx;
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(rx) : "l"(fl));
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(ry) : "l"(fl + gg));
        }
    };
    // flow -> geometry; contributes the window origin to the bounding box of tile slot `slot`
    auto make_geom = [&](float rx, float ry, int slot) {
        PointGeom g;
        g.xb = 0; g.yb = 0; g.fx = 0.f; g.fy = 0.f; g.live = false;
        const float sx = unnormalize(rx, p.Ws), sy = unnormalize(ry, p.Hs);
        if (fabsf(sx) < 1e6f && fabsf(sy) < 1e6f) {           // false for NaN (no point) too
            const float x0f = floorf(sx), y0f = floorf(sy);
            g.xb = (int)x0f - R; g.yb = (int)y0f - R;
            g.fx = sx - x0f; g.fy = sy - y0f;
            g.live = !(g.xb >= p.Ws || g.xb + W <= 0 || g.yb >= p.Hs || g.yb + W <= 0);
        }
        const int mnx = warp_min(g.live ? g.xb : INT_MAX), mny = warp_min(g.live ? g.yb : INT_MAX);
        const int mxx = warp_max(g.live ? g.xb : INT_MIN), mxy = warp_max(g.live ? g.yb : INT_MIN);
        if (lane == 0) {
            if (mnx != INT_MAX) {
                atomicMin(&bbox[slot].mnx, mnx); atomicMin(&bbox[slot].mny, mny);
                atomicMax(&bbox[slot].mxx, mxx); atomicMax(&bbox[slot].mxy, mxy);
            }
            mbar_arrive(&gready[slot]);                       // release: the atomics above are visible to the producer
        }
        return g;
    };

    const unsigned char* ring_b = reinterpret_cast<const unsigned char*>(ring);
    uint32_t q = 0;
    float rx, ry;
    TileCoord tc = coord_of((int)blockIdx.x), tc2 = tc;   // this tile, the tile whose flow is loaded next
    load_flow(blockIdx.x, tc2, rx, ry);
    PointGeom g_nxt = make_geom(rx, ry, 0);
    advance(tc2);
    load_flow(blockIdx.x + gridDim.x, tc2, rx, ry);
    int n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++n, advance(tc)) {
        const int slot = n % NTS;
        const PointGeom gm = g_nxt;
        if (tile + gridDim.x < ntiles) g_nxt = make_geom(rx, ry, (n + 1) % NTS);     // bounding box of the next tile
        advance(tc2);
        load_flow(tile + 2 * gridDim.x, tc2, rx, ry);
        mbar_wait(&tfull[slot], (n / NTS) & 1);
        const TileInfo ti = info[slot];
        const int b = tc.b, gx = tc.tx * TX + px, gy = tc.ty * TY + py;
        const bool valid = gx < G && gy < G;
        const bool live = gm.live;
        const int gu = gm.xb - ti.X0, goy = gm.yb - ti.Y0;
        const bool fit = live && gu + W <= ti.P && goy + W <= ti.bh;
        const int u = fit ? gu : 0, oy = fit ? goy : 0;
        const int m = u % W, nn = oy % W;
        // instruction (jj, ii) reads box row oy - nn + jj (+ W if jj < nn), box column u - m + ii (+ W if ii < m)
        int rowb[W], colb[W];                                 // byte offsets
#pragma unroll
        for (int k = 0; k < W; ++k) {
            rowb[k] = (oy - nn + k + (k < nn ? W : 0)) * ti.P * 4;
            colb[k] = (u - m + k + (k < m ? W : 0)) * 4;
        }
        unsigned long long acc[W][W / 2];
#pragma unroll
        for (int j = 0; j < W; ++j)
#pragma unroll
            for (int h = 0; h < W / 2; ++h) acc[j][h] = 0ull;
        if (ti.P != 0) {
#pragma unroll 1
            for (int c = 0; c < C; c += 2, q += 2) {          // two channels per step: ring slots s, s + 1 (NBUF even)
                const uint32_t s = q % NBUF;
                const uint32_t par = (q / NBUF) & 1;
                mbar_wait(&full[s], par);
                if (fit && !(p.debug & 1)) {
                    const unsigned char* sb = ring_b + s * (SLOT * 4);
                    const float fa = *reinterpret_cast<const float*>(sb + (BOXMAX + tid) * 4);
                    const float fb = *reinterpret_cast<const float*>(sb + (BOXMAX + tid) * 4 + SLOT * 4);
                    const unsigned long long f0a = pack2(fa, fa), f0b = pack2(fb, fb);
#pragma unroll
                    for (int jj = 0; jj < W; ++jj) {
                        const unsigned char* rb = sb + rowb[jj];
#pragma unroll
                        for (int h = 0; h < W / 2; ++h) {
                            const unsigned char* a0 = rb + colb[2 * h];
                            const unsigned char* a1 = rb + colb[2 * h + 1];
                            const float v0 = *reinterpret_cast<const float*>(a0), v1 = *reinterpret_cast<const float*>(a1);
                            const float w0 = *reinterpret_cast<const float*>(a0 + SLOT * 4);
                            const float w1 = *reinterpret_cast<const float*>(a1 + SLOT * 4);
                            ffma2(acc[jj][h], pack2(v0, v1), f0a);
                            ffma2(acc[jj][h], pack2(w0, w1), f0b);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
        }

        float* outp = p.out + ((size_t)b * p.k_total + p.k_offset) * gg + (size_t)gy * G + gx;
        if (fit) {
            // un-rotate: D[j][i] = acc[(j + nn) % W][(i + m) % W]
            float D[W][W];
#pragma unroll
            for (int jj = 0; jj < W; ++jj)
#pragma unroll
                for (int h = 0; h < W / 2; ++h) {
                    const float2 v = unpack2(acc[jj][h]);
                    D[jj][2 * h] = v.x; D[jj][2 * h + 1] = v.y;
                }
#pragma unroll
            for (int jj = 0; jj < W; ++jj) rot6(D[jj], m);
#pragma unroll
            for (int i = 0; i < W; ++i) {
                float col[W];
#pragma unroll
                for (int jj = 0; jj < W; ++jj) col[jj] = D[jj][i];
                rot6(col, nn);
#pragma unroll
                for (int jj = 0; jj < W; ++jj) D[jj][i] = col[jj];
            }
            const float a1 = gm.fx, a0 = 1.f - gm.fx;
            const float wy1 = gm.fy * p.inv_sqrt_c, wy0 = (1.f - gm.fy) * p.inv_sqrt_c;
            float hprev[KW];
#pragma unroll
            for (int j = 0; j < W; ++j)
#pragma unroll
                for (int i = 0; i < KW; ++i) {
                    const float h = a0 * D[j][i] + a1 * D[j][i + 1];
                    if (j >= 1) st_stream(outp + (size_t)((j - 1) * KW + i) * gg, wy0 * hprev[i] + wy1 * h);
                    hprev[i] = h;
                }
        } else if (valid) {
            if (!live) {
                for (int k = 0; k < KK; ++k) st_stream(outp + (size_t)k * gg, 0.f);
            } else {
                atomicAdd(&g_v2_stats[0], 1ull);
                lc_point_global<R>(p, b, gy, gx, gm, outp);
            }
        }
    }
}

template <int C, int NBUF, int MBW = rot::NBW - 1, int MBH = rot::NBH - 1>
static int launch_rot(const LcParams& p, cudaStream_t st) {
    using namespace rot;
    constexpr int SLOT = box_w(MBW) * box_h(MBH) + NT;
    constexpr size_t smem = (size_t)NBUF * SLOT * sizeof(float) + NTS * (sizeof(TileInfo) + sizeof(BBox)) +
                            (2 * NBUF + 2 * NTS) * sizeof(uint64_t) + 128;
    const int G = p.G;
    if (G % 4 != 0) return GFB_EUNSUPPORTED;              // 16-byte global strides of the f0 tensor map
    Maps maps;
    CUtensorMap tmap0;
    for (int i = 0; i < NBW; ++i)
        for (int j = 0; j < NBH; ++j) {
            uint64_t dims[3] = {(uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)p.B * p.C};
            uint64_t strides[2] = {(uint64_t)p.pitch * 4, (uint64_t)p.Hs * p.pitch * 4};
            uint32_t box[3] = {(uint32_t)box_w(i), (uint32_t)box_h(j), 1u};
            int rc = gfb_encode_tmap_f32(&maps.m[i][j], p.f1, 3, dims, strides, box, 0);
            if (rc != GFB_OK) return rc;
        }
    {
        uint64_t dims[3] = {(uint64_t)G, (uint64_t)G, (uint64_t)p.B * p.f0_ctot};
        uint64_t strides[2] = {(uint64_t)G * 4, (uint64_t)G * G * 4};
        uint32_t box[3] = {(uint32_t)TX, (uint32_t)TY, 1u};
        int rc = gfb_encode_tmap_f32(&tmap0, p.f0, 3, dims, strides, box, 0);
        if (rc != GFB_OK) return rc;
    }
    auto kern = lc_rot_kernel<C, NBUF, MBW, MBH>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long tiles = (long long)p.B * ((G + TY - 1) / TY) * ((G + TX - 1) / TX);
    if (tiles > 0x7fffffffLL) return GFB_EUNSUPPORTED;
    int dev = 0, sms = 148, per_sm = 1;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT + 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int grid = (int)min(tiles, (long long)sms * per_sm);
    kern<<<grid, NT + 32, smem, st>>>(p, (int)tiles, maps, tmap0);
    GFB_LAUNCH_RESULT();
}

// =====================================================================================================================
// tcgen05 path
// =====================================================================================================================
struct TileDesc {
    int x0, y0;      // first column (may be negative: TMA zero-fills) / first image row (clipped) of the staged region
    int nrows;       // image rows streamed
    int ylo, yhi;    // unclipped row range of the union of the windows
    int flags;
    int wx[4];       // per epilogue warp: smallest window origin of its 32 points
    int bwi;         // staged row width of this tile = NBW_FIRST + 8 * bwi positions (index of the TMA box)
    int pad;
};
enum { TF_EMPTY = 1, TF_GATHER = 2 };

// The main kernel is bound by L2 -> shared-memory traffic of the haloed B tiles, so every tile streams the narrowest box
// that holds its windows: one tensor map per width 32, 40, ..., 64.
constexpr int NBW = 5, NBW_FIRST = 32;
struct TmapSet { CUtensorMap m[NBW]; };

struct TcCfg {
    int blx, bly;            // lattice points per warp block (blx * bly = 32)
    int nbx, nby;            // warp blocks per tile (nbx * nby = 4)
    int tiles_x, tiles_y, ntiles;
    int bw, rps;             // widest staged row (64), image rows per B stage (N = bw * rps <= 128)
    int nstb;
};

constexpr int NMAX = 128;            // positions per B stage = TMEM columns per accumulator
constexpr int NACC = 2;              // accumulators per CTA (256 TMEM columns; two CTAs share an SM)
constexpr int EPI_WARPS = 4;         // warp w reads TMEM lanes 32 w .. 32 w + 31
constexpr int TMA_WARP = 4, MMA_WARP = 5;
constexpr int TC_THREADS = 6 * 32;
constexpr int LDW = 32;              // TMEM columns an epilogue warp pulls per image row
constexpr int RPS = 2;               // image rows per B stage / accumulator (N = RPS * bw <= 128)

__device__ __forceinline__ void tile_point(const TcCfg& c, int tx, int ty, int m, int& gy, int& gx) {
    const int w = m >> 5, l = m & 31;
    const int bx = w % c.nbx, by = w / c.nbx;
    gx = tx * (c.nbx * c.blx) + bx * c.blx + l % c.blx;
    gy = ty * (c.nby * c.bly) + by * c.bly + l / c.blx;
}

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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t bf16x2_rn(float upper, float lower) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
    return d;
}

__device__ __forceinline__ void st_global_256(uint32_t* ptr, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// ---- pre-pass + plan (one launch) ------------------------------------------------------------------------------
// 16 channels of one position: fp32 NCHW -> [bf16 hi(C) | bf16 lo(C)] position-major
template <int C>
__device__ __forceinline__ void prep_unit(const float* __restrict__ x, uint32_t* __restrict__ ws, size_t u,
                                          int H, int W, int pitch, int Ctot, int c0) {
    constexpr int NG = C / 16;
    const size_t npos = (size_t)H * W, plane = (size_t)H * pitch;
    const size_t pos = u % npos;
    const size_t t = u / npos;
    const int g = (int)(t % NG);
    const size_t b = t / NG;
    const int y = (int)(pos / W), xx = (int)(pos - (size_t)y * W);
    const float* src = x + (b * Ctot + c0 + (size_t)g * 16) * plane + (size_t)y * pitch + xx;
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = __ldg(src + (size_t)e * plane);
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const uint32_t h = bf16x2_rn(v[2 * e + 1], v[2 * e]);
        const float h0 = __uint_as_float(h << 16), h1 = __uint_as_float(h & 0xffff0000u);
        hi[e] = h;
        lo[e] = bf16x2_rn(v[2 * e + 1] - h1, v[2 * e] - h0);
    }
    uint32_t* row = ws + (b * npos + pos) * C;          // 2C bf16 = C words per position
    // one 256-bit store per 32-byte piece: a 16-byte store would still move a whole 32-byte sector to L2
    st_global_256(row + g * 8, hi);
    st_global_256(row + C / 2 + g * 8, lo);
}

// blocks [0, nplan): bounding box of the windows of two tiles each (128 threads per tile); the other blocks convert
// f0 and f1 (grid-stride)
template <int C>
__global__ void __launch_bounds__(256) lc_prep_plan_kernel(const LcParams p, const TcCfg c, TileDesc* __restrict__ plan,
                                                           uint32_t* __restrict__ ws0, uint32_t* __restrict__ ws1,
                                                           int nplan) {
    if ((int)blockIdx.x >= nplan) {
        const size_t u0 = (size_t)p.B * (C / 16) * p.G * p.G, u1 = (size_t)p.B * (C / 16) * p.Hs * p.Ws;
        const size_t stride = (size_t)(gridDim.x - nplan) * blockDim.x;
        for (size_t u = (size_t)(blockIdx.x - nplan) * blockDim.x + threadIdx.x; u < u0 + u1; u += stride) {
            if (u < u0) prep_unit<C>(p.f0, ws0, u, p.G, p.G, p.G, p.f0_ctot, p.c0);
            else prep_unit<C>(p.f1, ws1, u - u0, p.Hs, p.Ws, p.pitch, p.Ctot, p.c0);
        }
        return;
    }
    __shared__ int sm[2][4][4];
    const int half = threadIdx.x >> 7, m = threadIdx.x & 127, warp = m >> 5, lane = m & 31;
    const int tile = blockIdx.x * 2 + half;
    const int R = p.r, W = 2 * R + 2, G = p.G;
    int t = min(tile, c.ntiles - 1);
    const int tx = t % c.tiles_x; t /= c.tiles_x;
    const int ty = t % c.tiles_y;
    const int b = t / c.tiles_y;
    int gy, gx;
    tile_point(c, tx, ty, m, gy, gx);
    const PointGeom pg = point_geom(p, b, gy, gx, gy < G && gx < G, R);
    const int xmin = warp_min(pg.live ? pg.xb : INT_MAX), xmax = warp_max(pg.live ? pg.xb + W : INT_MIN);
    const int ymin = warp_min(pg.live ? pg.yb : INT_MAX), ymax = warp_max(pg.live ? pg.yb + W : INT_MIN);
    if (lane == 0) { sm[half][warp][0] = xmin; sm[half][warp][1] = xmax; sm[half][warp][2] = ymin; sm[half][warp][3] = ymax; }
    __syncthreads();
    if (m == 0 && tile < c.ntiles) {
        TileDesc d;
        int X0 = INT_MAX, X1 = INT_MIN, Y0 = INT_MAX, Y1 = INT_MIN;
        for (int w = 0; w < 4; ++w) {
            X0 = min(X0, sm[half][w][0]); X1 = max(X1, sm[half][w][1]);
            Y0 = min(Y0, sm[half][w][2]); Y1 = max(Y1, sm[half][w][3]);
        }
        d.x0 = 0; d.y0 = 0; d.nrows = 0; d.ylo = 0; d.yhi = 0; d.flags = 0; d.bwi = 0; d.pad = 0;
        for (int w = 0; w < 4; ++w) d.wx[w] = 0;
        if (X0 == INT_MAX) {
            d.flags = TF_EMPTY;
        } else {
            d.x0 = X0;
            d.y0 = max(Y0, 0);
            d.nrows = min(Y1, p.Hs) - d.y0;
            d.ylo = Y0; d.yhi = Y1;
            for (int w = 0; w < 4; ++w) d.wx[w] = sm[half][w][0] == INT_MAX ? X0 : sm[half][w][0];
            d.bwi = min(max((X1 - X0 - NBW_FIRST + 7) / 8, 0), NBW - 1);
            if (X1 - X0 > c.bw) { d.flags = TF_GATHER; atomicAdd(&g_v2_stats[2], 1ull); }
        }
        plan[tile] = d;
    }
}

// single-thread roles poll with a back-off so that they do not steal issue slots from the epilogue warps
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, int mode = 0) {
    if (mode & 128) { while (!mbar_test_wait(bar, parity)) { } return; }     // debug: non-blocking test in a tight loop
    while (!mbar_try_wait(bar, parity))
        if (!(mode & 32)) __nanosleep(32);
}
__device__ __forceinline__ void st_stream_pred(float* ptr, float v, bool pred) {
    // no "memory" clobber: the outputs are never read back, and the compiler must stay free to hoist the staging reads
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.cs.f32 [%0], %1;\n\t}"
                 ::"l"(ptr), "f"(v), "r"((int)pred));
}

// corr of a channel slice: the first slice stores, the following ones add (correlation is linear in the channels)
__device__ __forceinline__ void out_put(float* ptr, float v, int accumulate) {
    if (accumulate) v += __ldcs(ptr);
    __stcs(ptr, v);
}

// ---- main kernel ---------------------------------------------------------------------------------------------------
template <int R, int C, bool ACC>      // ACC: out += (second and later channel slices); compile-time so that the store path stays branch-free
__global__ void __launch_bounds__(TC_THREADS, 2)
lc_tc2_kernel(const LcParams p, const TcCfg c, const TileDesc* __restrict__ plan,
              const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ TmapSet tmapB) {
    constexpr int W = 2 * R + 2, KW = 2 * R + 1, KK = KW * KW;
    constexpr int ATOMS = (2 * C * 2 + 127) / 128;          // 128-byte atoms per K-major row [hi(C) | lo(C)] of bf16
    constexpr int NKS = C / 16;                              // K = 16 steps per part
    constexpr uint32_t A_ATOM = 128 * 128, B_ATOM = NMAX * 128;
    constexpr uint32_t A_STAGE = ATOMS * A_ATOM, B_STAGE = ATOMS * B_ATOM;
    constexpr int NSTA = C <= 32 ? 2 : 1;                    // A tiles in shared memory (two where they fit: the next tile's A
                                                             // arrives while this tile still runs)
    static_assert(W <= LDW, "window wider than the TMEM pull");
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_base = smem;
    unsigned char* b_base = smem + NSTA * A_STAGE;
    float* ebuf = reinterpret_cast<float*>(b_base + (size_t)c.nstb * B_STAGE);     // [EPI_WARPS][LDW][32 lanes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ebuf + EPI_WARPS * LDW * 32);
    uint64_t* a_full = bars;               // [NSTA <= 2]
    uint64_t* a_empty = a_full + 2;        // [NSTA]
    uint64_t* b_full = a_empty + 2;        // [nstb <= 6]
    uint64_t* b_empty = b_full + 6;        // [nstb]
    uint64_t* d_full = b_empty + 6;        // [NACC]
    uint64_t* d_empty = d_full + NACC;     // [NACC]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + NACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = p.G;
    const size_t gg = (size_t)G * G;

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();          // the swizzled operand stages need 1024-byte alignment
        for (int s = 0; s < NSTA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < c.nstb; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < NACC; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, NACC * NMAX);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == TMA_WARP) {
        // ================= TMA producer =================
        if (lane == 0) {
            tma_prefetch_desc(&tmapA);
            for (int k = 0; k < NBW; ++k) tma_prefetch_desc(&tmapB.m[k]);
            uint32_t q = 0, tt = 0;
            for (int tile = blockIdx.x; tile < c.ntiles; tile += gridDim.x) {
                const TileDesc d = plan[tile];
                if (d.flags) continue;
                int t = tile;
                const int tx = t % c.tiles_x; t /= c.tiles_x;
                const int ty = t % c.tiles_y;
                const int b = t / c.tiles_y;
                const int nchunks = (d.nrows + RPS - 1) / RPS;
                const CUtensorMap* tmB = &tmapB.m[d.bwi];
                const uint32_t b_bytes = (uint32_t)(ATOMS * RPS * (NBW_FIRST + 8 * d.bwi) * 128);
                // the first B stages of this tile go out before its A tile: the A buffer is free only when the previous
                // tile's last MMA has retired, the B ring usually has room earlier
                // (with two A buffers the A tile simply goes first)
                const int pre = NSTA > 1 ? 0 : min(nchunks, c.nstb);
                const uint32_t as = tt % NSTA, aph = (tt / NSTA) & 1;
                unsigned char* a_dst = a_base + as * A_STAGE;
                for (int ch = 0; ch < nchunks; ++ch, ++q) {
                    if (ch == pre) {
                        mbar_wait_sleep(&a_empty[as], aph ^ 1, p.debug);
                        mbar_expect_tx(&a_full[as], A_STAGE);
#pragma unroll
                        for (int at = 0; at < ATOMS; ++at)
                            for (int w = 0; w < 4; ++w) {
                                const int bx = w % c.nbx, by = w / c.nbx;
                                tma_load_4d(a_dst + at * A_ATOM + w * 4096, &tmapA, &a_full[as], at * 32,
                                            tx * (c.nbx * c.blx) + bx * c.blx, ty * (c.nby * c.bly) + by * c.bly, b);
                            }
                    }
                    const uint32_t s = q % c.nstb;
                    mbar_wait_sleep(&b_empty[s], ((q / c.nstb) & 1) ^ 1, p.debug);
                    if (p.debug & 8) { mbar_arrive(&b_full[s]); continue; }
                    mbar_expect_tx(&b_full[s], b_bytes);
#pragma unroll
                    for (int at = 0; at < ATOMS; ++at)
                        tma_load_4d(b_base + (size_t)s * B_STAGE + at * B_ATOM, tmB, &b_full[s], at * 32, d.x0,
                                    d.y0 + ch * RPS, b);
                }
                if (pre == nchunks) {                  // short tile: every B stage went out first
                    mbar_wait_sleep(&a_empty[as], aph ^ 1, p.debug);
                    mbar_expect_tx(&a_full[as], A_STAGE);
#pragma unroll
                    for (int at = 0; at < ATOMS; ++at)
                        for (int w = 0; w < 4; ++w) {
                            const int bx = w % c.nbx, by = w / c.nbx;
                            tma_load_4d(a_dst + at * A_ATOM + w * 4096, &tmapA, &a_full[as], at * 32,
                                        tx * (c.nbx * c.blx) + bx * c.blx, ty * (c.nby * c.bly) + by * c.bly, b);
                        }
                }
                ++tt;
            }
        }
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t q = 0, tt = 0;
            for (int tile = blockIdx.x; tile < c.ntiles; tile += gridDim.x) {
                const TileDesc d = plan[tile];
                if (d.flags) continue;
                const uint32_t as = tt % NSTA;
                const uint32_t a_addr = smem_u32(a_base + as * A_STAGE);
                mbar_wait_sleep(&a_full[as], (tt / NSTA) & 1, p.debug);
                const int nchunks = (d.nrows + RPS - 1) / RPS;
                const uint32_t idesc = idesc_bf16(RPS * (NBW_FIRST + 8 * d.bwi));
                for (int ch = 0; ch < nchunks; ++ch, ++q) {
                    const uint32_t s = q % c.nstb, acc = q % NACC;
                    mbar_wait_sleep(&b_full[s], (q / c.nstb) & 1, p.debug);
                    mbar_wait_sleep(&d_empty[acc], ((q / NACC) & 1) ^ 1, p.debug);
                    fence_after_sync();
                    const uint32_t b_addr = smem_u32(b_base + (size_t)s * B_STAGE);
                    const uint32_t dt = tmem_base + acc * (uint32_t)NMAX;
                    uint32_t accum = 0;
#pragma unroll
                    for (int combo = (p.debug & 16) ? 3 : 0; combo < 3; ++combo) {
                        const int pa = combo == 2 ? 1 : 0, pb = combo == 1 ? 1 : 0;   // lo*hi, hi*lo, then hi*hi last
#pragma unroll
                        for (int ks = 0; ks < NKS; ++ks) {
                            const uint32_t oa = (uint32_t)(pa * C * 2 + ks * 32), ob = (uint32_t)(pb * C * 2 + ks * 32);
                            const uint64_t ad = smem_desc_k128(a_addr + (oa >> 7) * A_ATOM + (oa & 127u));
                            const uint64_t bd = smem_desc_k128(b_addr + (ob >> 7) * B_ATOM + (ob & 127u));
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
        // staging: column v of the pull goes to stg[v * 32] -- a lane-private column of a [LDW][32] block, so both the
        // 32 stores and the reads at the lane's own offset are bank-conflict free
        const int quad = warp;
        float* stg = ebuf + (size_t)warp * LDW * 32 + lane;
        auto stage = [&](const uint32_t (&r)[32]) {
            if (p.debug & 4) { stg[0] = __uint_as_float(r[0] ^ r[31]); return; }
#pragma unroll
            for (int v = 0; v < 32; ++v) stg[v * 32] = __uint_as_float(r[v]);
        };
        uint32_t q = 0;
        long long dbg_wait = 0, dbg_pull = 0, dbg_rows = 0, dbg_n = 0;
        const bool dbg = p.debug & 1;
        for (int tile = blockIdx.x; tile < c.ntiles; tile += gridDim.x) {
            const TileDesc d = plan[tile];
            int t = tile;
            const int tx = t % c.tiles_x; t /= c.tiles_x;
            const int ty = t % c.tiles_y;
            const int b = t / c.tiles_y;
            int gy, gx;
            tile_point(c, tx, ty, quad * 32 + lane, gy, gx);
            const bool valid = gy < G && gx < G;
            float* outp = p.out + ((size_t)b * p.k_total + p.k_offset) * gg + (size_t)gy * G + gx;
            if (d.flags & TF_GATHER) {
                if (valid)
                    for (int k = 0; k < KK; ++k) out_put(outp + (size_t)k * gg, lc_generic_point(p, b, k, gy, gx), p.accumulate);
                continue;
            }
            const PointGeom pg = point_geom(p, b, gy, gx, valid, R);
            bool live = pg.live;
            if (valid && !live && !p.accumulate)
                for (int k = 0; k < KK; ++k) st_stream(outp + (size_t)k * gg, 0.f);
            if (d.flags & TF_EMPTY) continue;
            const int tbw = NBW_FIRST + 8 * d.bwi;                 // this tile's staged row width
            const int start = min(max(d.wx[quad] - d.x0, 0), tbw - LDW);
            const int off = pg.xb - d.x0 - start;
            if (live && (off < 0 || off + W > LDW)) {      // window outside the warp's TMEM pull: exact gather
                atomicAdd(&g_v2_stats[1], 1ull);
                for (int k = 0; k < KK; ++k) out_put(outp + (size_t)k * gg, lc_generic_point(p, b, k, gy, gx), p.accumulate);
                live = false;
            }
            const int col0 = live ? off : 0;                       // first staged column this lane reads (col0 + KW <= LDW - 1)
            const float a1 = pg.fx, a0 = 1.f - pg.fx;
            const float wy1 = pg.fy * p.inv_sqrt_c, wy0 = (1.f - pg.fy) * p.inv_sqrt_c;
            const int yb = pg.yb;
            float hprev[KW];
#pragma unroll
            for (int i = 0; i < KW; ++i) hprev[i] = 0.f;
            const unsigned gg32 = (unsigned)gg;
            // element index of this lane's first output; the launcher guarantees B * k_total * G * G < 2^31
            const unsigned lane_idx = (((unsigned)b * (unsigned)p.k_total + (unsigned)p.k_offset) * (unsigned)G + (unsigned)gy) * (unsigned)G + (unsigned)gx;
            // one image row of D, read back from the staging column at the lane's own offset (immediate offsets from
            // rowp): x-lerp, y-lerp with the previous row, predicated streaming stores addressed as out[32-bit index]
            // (a uniform base + one IMAD.WIDE per store instead of 64-bit pointer arithmetic per output)
            const float* rowp = stg + col0 * 32;
            auto emit_row = [&](int j, bool act) {
                const bool st = act && j >= 1 && !(p.debug & 2);
                unsigned idx = lane_idx + (unsigned)(j - 1) * (KW * gg32);
                float d0 = rowp[0];
#pragma unroll
                for (int i = 0; i < KW; ++i) {
                    const float d1 = rowp[(i + 1) * 32];
                    const float h = a0 * d0 + a1 * d1;
                    if (st) out_put(p.out + idx, wy0 * hprev[i] + wy1 * h, ACC);
                    idx += gg32;
                    hprev[i] = h;
                    d0 = d1;
                }
            };
            // a row outside the image: D = 0
            auto emit_zero = [&](int j, bool act) {
                const bool st = act && j >= 1;
                unsigned idx = lane_idx + (unsigned)(j - 1) * (KW * gg32);
#pragma unroll
                for (int i = 0; i < KW; ++i) {
                    if (st) out_put(p.out + idx, wy0 * hprev[i], ACC);
                    idx += gg32;
                    hprev[i] = 0.f;
                }
            };
            for (int y = d.ylo; y < min(d.y0, d.yhi); ++y) {            // rows above the image
                const int j = y - yb;
                const bool act = live && (unsigned)j < (unsigned)W;
                if (__any_sync(0xffffffffu, act)) emit_zero(j, act);
            }
            // Software-pipelined pull: while the rows of chunk ch are interpolated and stored, the two TMEM loads of chunk
            // ch + 1 are already in flight (their registers were freed by staging the current rows first), and the
            // accumulator goes back to the MMA issuer as soon as they have landed.
            const int nchunks = (d.nrows + RPS - 1) / RPS;
            const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)start;
            uint32_t rA[32], rB[32];
            bool actA, actB, anyA, anyB;
            int jA;
            auto activity = [&](int ch) {
                jA = d.y0 + ch * RPS - yb;
                actA = live && (unsigned)jA < (unsigned)W;
                actB = live && (ch * RPS + 1 < d.nrows) && (unsigned)(jA + 1) < (unsigned)W;
                anyA = __any_sync(0xffffffffu, actA);
                anyB = __any_sync(0xffffffffu, actB);
            };
            auto release = [&](uint32_t acc) {
                tmem_ld_wait();
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[acc]);
            };
            {   // first chunk of the tile: nothing to overlap with
                const long long c0 = dbg ? clock64() : 0;
                const uint32_t acc = q % NACC;
                mbar_wait(&d_full[acc], (q / NACC) & 1);
                fence_after_sync();
                activity(0);
                if (anyA) tmem_ld32(tlane + acc * (uint32_t)NMAX, rA);
                if (anyB) tmem_ld32(tlane + acc * (uint32_t)NMAX + (uint32_t)tbw, rB);
                release(acc);
                if (dbg) dbg_pull += clock64() - c0;
            }
            for (int ch = 0; ch < nchunks; ++ch, ++q) {
                const long long c1 = dbg ? clock64() : 0;
                const bool cA = actA, cB = actB, hA = anyA, hB = anyB;
                const int cj = jA;
                const bool more = ch + 1 < nchunks;
                const uint32_t nacc = (q + 1) % NACC;
                if (more) {
                    if (p.debug & 128) { while (!mbar_test_wait(&d_full[nacc], ((q + 1) / NACC) & 1)) { } }
                    else mbar_wait(&d_full[nacc], ((q + 1) / NACC) & 1);
                    fence_after_sync();
                    activity(ch + 1);
                }
                const long long c2 = dbg ? clock64() : 0;
                const bool emit = !(p.debug & 64);     // bit 6: pulls and hand-offs only
                if (hA) stage(rA);
                if (more && anyA) tmem_ld32(tlane + nacc * (uint32_t)NMAX, rA);
                if (hA && emit) emit_row(cj, cA);
                if (hB) stage(rB);
                if (more && anyB) tmem_ld32(tlane + nacc * (uint32_t)NMAX + (uint32_t)tbw, rB);
                if (hB && emit) emit_row(cj + 1, cB);
                if (more) release(nacc);
                if (dbg) { const long long c3 = clock64(); dbg_wait += c2 - c1; dbg_rows += c3 - c2; ++dbg_n; }
            }
            for (int y = max(d.y0 + d.nrows, d.ylo); y < d.yhi; ++y) {   // rows below the image
                const int j = y - yb;
                const bool act = live && (unsigned)j < (unsigned)W;
                if (__any_sync(0xffffffffu, act)) emit_zero(j, act);
            }
        }
        if (dbg && warp == 0 && lane == 0) {
            atomicAdd(&g_v2_stats[4], (unsigned long long)dbg_wait);
            atomicAdd(&g_v2_stats[5], (unsigned long long)dbg_pull);
            atomicAdd(&g_v2_stats[6], (unsigned long long)dbg_rows);
            atomicAdd(&g_v2_stats[7], (unsigned long long)dbg_n);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, NACC * NMAX);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void tc2_config(int G, int Ws, int r, TcCfg& c) {
    const int W = 2 * r + 2;
    const float s = (float)Ws / (float)G;
    // warp block 8 x 4 points when its windows fit one 32-column TMEM pull, else 4 x 8
    c.blx = ((int)ceilf(8.f * s * 1.35f) + W + 1 <= LDW) ? 8 : 4;
    c.bly = 32 / c.blx;
    c.nbx = 16 / c.blx; c.nby = 4 / c.nbx;                 // tile = 16 x 8 points
    const int tw = c.nbx * c.blx, th = c.nby * c.bly;
    c.tiles_x = (G + tw - 1) / tw;
    c.tiles_y = (G + th - 1) / th;
    (void)s;
    c.bw = NMAX / RPS;                                     // widest box; tiles whose windows spread further take the gather path
    c.rps = RPS;
}

static int tc2_group(int B, size_t per_elem, int group) {
    // elements per group (equal sizes): one launch pair for the whole batch measured fastest (fewer, longer persistent
    // kernels; the workspace need not stay L2-resident), so the cap only bounds the workspace at 512 MB
    int gb = group > 0 ? group : (int)max((size_t)1, (size_t)(512u << 20) / per_elem);
    gb = min(gb, B);
    const int ngroups = (B + gb - 1) / gb;
    return (B + ngroups - 1) / ngroups;
}

template <int R, int C>
// phase 0: pre-pass + plan + main; 1: pre-pass only (feature0 / feature1 -> workspace); 2: plan + main on a workspace that
// phase 1 filled (the features of a refiner scale do not change between its iterations, only the flow does)
static int launch_tc2(const LcParams& p0, cudaStream_t st, void* workspace, size_t ws_bytes, int group, int phase = 0) {
    constexpr int ATOMS = (2 * C * 2 + 127) / 128;
    TcCfg c;
    tc2_config(p0.G, p0.Ws, R, c);
    const int G = p0.G;
    const size_t a_stage = (size_t)ATOMS * 128 * 128, b_stage = (size_t)ATOMS * NMAX * 128;
    const size_t fixed = (C <= 32 ? 2 : 1) * a_stage + (size_t)EPI_WARPS * LDW * 32 * sizeof(float) + 24 * sizeof(uint64_t);
    const size_t budget = 112 * 1024 + 256;                // two CTAs per SM
    c.nstb = (int)min((size_t)6, (budget - fixed) / b_stage);
    if (c.nstb < 2) return GFB_EUNSUPPORTED;
    if ((size_t)p0.B * p0.k_total * G * G >= (1ull << 31)) return GFB_EUNSUPPORTED;     // 32-bit output indices
    const size_t smem = fixed + c.nstb * b_stage;

    const size_t ws0_per = (size_t)G * G * C * 4, ws1_per = (size_t)p0.Hs * p0.Ws * C * 4;
    const int gb = tc2_group(p0.B, ws0_per + ws1_per, group);
    if (phase != 0 && gb < p0.B) return GFB_EUNSUPPORTED;  // a prepared workspace holds the whole batch
    const size_t plan_bytes = align_up((size_t)gb * c.tiles_x * c.tiles_y * sizeof(TileDesc), 1024);
    const size_t need = plan_bytes + align_up(gb * ws0_per, 1024) + gb * ws1_per;
    if (!workspace || ws_bytes < need) return GFB_EWORKSPACE;
    if (!gfb_aligned(workspace, 128)) return GFB_EALIGN;
    TileDesc* plan = reinterpret_cast<TileDesc*>(workspace);
    uint32_t* ws0 = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(workspace) + plan_bytes);
    uint32_t* ws1 = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(ws0) + align_up(gb * ws0_per, 1024));

    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (p0.accumulate && C != 64) return GFB_EUNSUPPORTED;                     // channel slices are 64 wide
    auto kern = (C == 64 && p0.accumulate) ? lc_tc2_kernel<R, C, (C == 64)> : lc_tc2_kernel<R, C, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;

    const size_t gg = (size_t)G * G, plane = (size_t)p0.Hs * p0.pitch;
    for (int b0 = 0; b0 < p0.B; b0 += gb) {
        LcParams p = p0;
        p.B = min(gb, p0.B - b0);
        p.f0 = p0.f0 + (size_t)b0 * p0.f0_ctot * gg;
        p.f1 = p0.f1 + (size_t)b0 * p0.Ctot * plane;
        p.flow = p0.flow ? p0.flow + (size_t)b0 * 2 * gg : nullptr;
        p.out = p0.out ? p0.out + (size_t)b0 * p0.k_total * gg : nullptr;
        c.ntiles = p.B * c.tiles_x * c.tiles_y;

        const int nplan = phase == 1 ? 0 : (c.ntiles + 1) / 2;
        const size_t units = (size_t)p.B * (C / 16) * (gg + (size_t)p.Hs * p.Ws);
        const int nprep = phase == 2 ? 0 : (int)min((size_t)sms * 5, (units + 255) / 256);    // one resident wave next to the short plan blocks
        lc_prep_plan_kernel<C><<<nplan + nprep, 256, 0, st>>>(p, c, plan, ws0, ws1, nplan);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        if (phase == 1) continue;

        CUtensorMap tmapA;
        TmapSet tmapB;
        {
            uint64_t dims[4] = {(uint64_t)C, (uint64_t)G, (uint64_t)G, (uint64_t)p.B};
            uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)G * C * 4, (uint64_t)gg * C * 4};
            uint32_t box[4] = {32u, (uint32_t)c.blx, (uint32_t)c.bly, 1u};
            int rc = gfb_encode_tmap_f32(&tmapA, ws0, 4, dims, strides, box, 3);
            if (rc != GFB_OK) return rc;
        }
        for (int k = 0; k < NBW; ++k) {
            uint64_t dims[4] = {(uint64_t)C, (uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)p.B};
            uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)p.Ws * C * 4, (uint64_t)p.Hs * p.Ws * C * 4};
            uint32_t box[4] = {32u, (uint32_t)(NBW_FIRST + 8 * k), (uint32_t)RPS, 1u};
            int rc = gfb_encode_tmap_f32(&tmapB.m[k], ws1, 4, dims, strides, box, 3);
            if (rc != GFB_OK) return rc;
        }
        kern<<<min(c.ntiles, 2 * sms), TC_THREADS, smem, st>>>(p, c, plan, tmapA, tmapB);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return GFB_OK;
}

}  // namespace lcv2
}  // namespace gfb

using namespace gfb;

static int fill_params(LcParams& p, const float* f0, const float* f1, const float* flow, float* out,
                       int B, int C, int Hs, int Ws, int f1_pitch, int G, int r, int k_total, int k_offset,
                       int Ctot = 0, int c0 = 0, int accumulate = 0) {
    GFB_CHECK_ARG(f0 && f1 && flow && out);
    GFB_CHECK_ARG(B > 0 && C > 0 && Hs > 0 && Ws > 0 && G > 0 && r >= 0);
    GFB_CHECK_ARG(f1_pitch == 0 || f1_pitch >= Ws);
    const int kk = (2 * r + 1) * (2 * r + 1);
    GFB_CHECK_ARG(k_offset >= 0 && k_offset + kk <= k_total);
    p.f0 = f0; p.f1 = f1; p.flow = flow; p.out = out;
    p.B = B; p.C = C; p.Hs = Hs; p.Ws = Ws; p.G = G; p.r = r;
    p.Ctot = Ctot > 0 ? Ctot : C; p.c0 = c0; p.accumulate = accumulate; p.f0_ctot = p.Ctot;
    GFB_CHECK_ARG(c0 >= 0 && c0 + C <= p.Ctot);
    p.pitch = f1_pitch ? f1_pitch : Ws;
    p.k_total = k_total; p.k_offset = k_offset;
    p.sample_mode = 0; p.padding_mode = 0;
    p.ox0 = (float)(-2.0 * r / Ws); p.ox1 = (float)(2.0 * r / Ws);
    p.oy0 = (float)(-2.0 * r / Hs); p.oy1 = (float)(2.0 * r / Hs);
    p.inv_sqrt_c = (float)(1.0 / sqrt((double)p.Ctot));
    p.debug = 0;
    return GFB_OK;
}

// One lattice point per thread (bilinear, zero padding, window offsets in pixels of f1).  f1 rows must be 16-byte
// multiples apart (f1_pitch % 4 == 0) for the TMA descriptor.
static int local_corr_pt(const float* f0, const float* f1, const float* flow, float* out,
                         int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                         int k_total, int k_offset, int tune, int debug, gfb_stream_t stream, int f0_ctot = 0) {
    LcParams p;
    int rc = fill_params(p, f0, f1, flow, out, B, C, Hs, Ws, f1_pitch, G, r, k_total, k_offset);
    if (rc != GFB_OK) return rc;
    p.debug = debug;
    if (f0_ctot) p.f0_ctot = f0_ctot;
    if (p.pitch % 4 != 0 || !gfb_aligned(f1, 16)) return GFB_EALIGN;
    if ((size_t)B * p.f0_ctot >= (1ull << 31)) return GFB_EUNSUPPORTED;
    cudaStream_t st = gfb_cu(stream);
    const float s = (float)Ws / (float)G;
    if (!gfb_aligned(f0, 16)) return GFB_EALIGN;
    // box = tile span x (1.3 magnification) + rotation shear + window + alignment, rounded up to a row pitch of 0 mod 32
    // floats: the 16 lanes of a half-warp then hit distinct banks whatever rows their windows start on
    if (r == 2 && C == 16) {
        if (tune == 0 || tune == 33) return lcv2::launch_rot<16, 6>(p, st);      // default: rotated-order kernel
        if (tune == 32) return lcv2::launch_rot<16, 4>(p, st);
        if (tune == 1 || (tune == 7 && s <= 1.2f)) return lcv2::launch_pt<2, 16, 16, 8, 64, 24, 2, 3>(p, st);
        if (tune == 2 || (tune == 7 && s <= 2.0f)) return lcv2::launch_pt<2, 16, 16, 8, 64, 32, 2, 3>(p, st);
        if (tune == 4) return lcv2::launch_pt<2, 16, 16, 8, 56, 32, 4, 2>(p, st);
        if (tune == 5) return lcv2::launch_pt<2, 16, 8, 16, 40, 48, 2, 3>(p, st);    // warp = 4 lattice rows x 8 points
        if (tune == 6) return lcv2::launch_pt<2, 16, 8, 16, 64, 48, 2, 3>(p, st);
        return lcv2::launch_pt<2, 16, 8, 8, 64, 40, 2, 3>(p, st);
    }
    if (r == 4 && C == 32) {
        if (tune == 1 || (tune == 0 && s <= 1.2f)) return lcv2::launch_pt<4, 32, 16, 8, 64, 28, 2, 2>(p, st);
        return lcv2::launch_pt<4, 32, 16, 8, 64, 36, 2, 2>(p, st);
    }
    if (r == 1 && C == 16) return lcv2::launch_pt<1, 16, 16, 8, 64, 32, 2, 3>(p, st);
    if (r == 1 && C == 8) return lcv2::launch_pt<1, 8, 16, 8, 64, 32, 2, 3>(p, st);
    if (r == 2 && C == 8) return lcv2::launch_pt<2, 8, 16, 8, 64, 32, 2, 3>(p, st);
    return GFB_EUNSUPPORTED;
}

extern "C" int gfb_local_corr_pt_f32(const float* f0, const float* f1, const float* flow, float* out,
                                     int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                                     int k_total, int k_offset, int tune, gfb_stream_t stream) {
    return local_corr_pt(f0, f1, flow, out, B, C, Hs, Ws, f1_pitch, G, r, k_total, k_offset, tune, 0, stream);
}
extern "C" int gfb_debug_local_corr_pt_f32(const float* f0, const float* f1, const float* flow, float* out,
                                           int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                                           int k_total, int k_offset, int tune, int debug, gfb_stream_t stream) {
    return local_corr_pt(f0, f1, flow, out, B, C, Hs, Ws, f1_pitch, G, r, k_total, k_offset, tune, debug, stream);
}

extern "C" int gfb_debug_local_corr_v2_counters(unsigned long long* host_out4, int reset) {
    cudaError_t e = cudaSuccess;
    if (host_out4) e = cudaMemcpyFromSymbol(host_out4, lcv2::g_v2_stats, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess && reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        e = cudaMemcpyToSymbol(lcv2::g_v2_stats, z, sizeof(z));
    }
    return e == cudaSuccess ? GFB_OK : (int)e;
}

extern "C" size_t gfb_local_corr_tc2_workspace_bytes(int B, int C, int Hs, int Ws, int G, int r, int group) {
    if (B <= 0 || C <= 0 || Hs <= 0 || Ws <= 0 || G <= 0 || r < 0) return 0;
    lcv2::TcCfg c;
    lcv2::tc2_config(G, Ws, r, c);
    const size_t ws0_per = (size_t)G * G * C * 4, ws1_per = (size_t)Hs * Ws * C * 4;
    const int gb = lcv2::tc2_group(B, ws0_per + ws1_per, group);
    return lcv2::align_up((size_t)gb * c.tiles_x * c.tiles_y * sizeof(lcv2::TileDesc), 1024) +
           lcv2::align_up(gb * ws0_per, 1024) + gb * ws1_per;
}

// (r, C) pairs the tcgen05 kernel is instantiated for: the GFNet scales (4,32) (6,64) (7,64) and the microbench sweep's radii
#define GFB_TC2_ALL GFB_TC2_CASE(4, 32) GFB_TC2_CASE(6, 64) GFB_TC2_CASE(7, 64) GFB_TC2_CASE(3, 32) GFB_TC2_CASE(5, 64) \
                    GFB_TC2_CASE(4, 64) GFB_TC2_CASE(2, 64) GFB_TC2_CASE(3, 64) GFB_TC2_CASE(8, 64) GFB_TC2_CASE(2, 32)

// Split form of gfb_local_corr_tc2_f32 for callers that correlate the same feature0 / feature1 against several flows
// (the iterations of one refiner scale, model/network.py:230-281): prepare once, run per flow.
extern "C" int gfb_local_corr_tc2_prepare_f32(const float* f0, const float* f1, int B, int C, int Hs, int Ws, int f1_pitch,
                                              int G, int r, void* workspace, size_t workspace_bytes, gfb_stream_t stream) {
    LcParams p;
    float dummy;
    const int kk = (2 * r + 1) * (2 * r + 1);
    int rc = fill_params(p, f0, f1, &dummy, &dummy, B, C, Hs, Ws, f1_pitch, G, r, kk, 0);
    if (rc != GFB_OK) return rc;
    p.flow = nullptr; p.out = nullptr;
    cudaStream_t st = gfb_cu(stream);
#define GFB_TC2_CASE(RR, CC) if (r == RR && C == CC) return lcv2::launch_tc2<RR, CC>(p, st, workspace, workspace_bytes, 0, 1);
    GFB_TC2_ALL
#undef GFB_TC2_CASE
    return GFB_EUNSUPPORTED;
}

extern "C" int gfb_local_corr_tc2_run_f32(const float* f0, const float* f1, const float* flow, float* out,
                                          int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                                          int k_total, int k_offset,
                                          void* workspace, size_t workspace_bytes, gfb_stream_t stream) {
    LcParams p;
    int rc = fill_params(p, f0, f1, flow, out, B, C, Hs, Ws, f1_pitch, G, r, k_total, k_offset);
    if (rc != GFB_OK) return rc;
    cudaStream_t st = gfb_cu(stream);
#define GFB_TC2_CASE(RR, CC) if (r == RR && C == CC) return lcv2::launch_tc2<RR, CC>(p, st, workspace, workspace_bytes, 0, 2);
    GFB_TC2_ALL
#undef GFB_TC2_CASE
    return GFB_EUNSUPPORTED;
}

// number of (pre-pass, main) launch pairs gfb_local_corr_tc2_f32 issues for these shapes (bench.py's launch count)
extern "C" int gfb_local_corr_tc2_groups(int B, int C, int Hs, int Ws, int G, int group) {
    if (B <= 0 || C <= 0 || Hs <= 0 || Ws <= 0 || G <= 0) return 0;
    const size_t per = (size_t)G * G * C * 4 + (size_t)Hs * Ws * C * 4;
    const int gb = lcv2::tc2_group(B, per, group);
    return (B + gb - 1) / gb;
}

// One channel slice of a correlation over Ctot channels: the channels c0 .. c0 + C of feature0 / feature1 ([B,Ctot,...]) are
// correlated (scaled by 1/sqrt(Ctot)) and stored (accumulate = 0) or added to out (accumulate = 1).  Correlation is linear
// in the channels, so C-channel kernels cover any Ctot that is a multiple of C (the microbench sweep's C = 128 .. 512).
extern "C" int gfb_local_corr_tc2_slice_f32(const float* f0, const float* f1, const float* flow, float* out,
                                            int B, int C, int Ctot, int c0, int accumulate, int Hs, int Ws, int f1_pitch, int G, int r,
                                            int k_total, int k_offset, void* workspace, size_t workspace_bytes, gfb_stream_t stream) {
    LcParams p;
    int rc = fill_params(p, f0, f1, flow, out, B, C, Hs, Ws, f1_pitch, G, r, k_total, k_offset, Ctot, c0, accumulate ? 1 : 0);
    if (rc != GFB_OK) return rc;
    cudaStream_t st = gfb_cu(stream);
#define GFB_TC2_CASE(RR, CC) if (r == RR && C == CC) return lcv2::launch_tc2<RR, CC>(p, st, workspace, workspace_bytes, 0);
    GFB_TC2_ALL
#undef GFB_TC2_CASE
    return GFB_EUNSUPPORTED;
}

static int local_corr_tc2(const float* f0, const float* f1, const float* flow, float* out,
                          int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                          int k_total, int k_offset, int group, int debug,
                          void* workspace, size_t workspace_bytes, gfb_stream_t stream) {
    LcParams p;
    int rc = fill_params(p, f0, f1, flow, out, B, C, Hs, Ws, f1_pitch, G, r, k_total, k_offset);
    if (rc != GFB_OK) return rc;
    GFB_CHECK_ARG(group >= 0 && group <= 255);
    p.debug = debug;
    cudaStream_t st = gfb_cu(stream);
#define GFB_TC2_CASE(RR, CC) if (r == RR && C == CC) return lcv2::launch_tc2<RR, CC>(p, st, workspace, workspace_bytes, group);
    GFB_TC2_ALL
#undef GFB_TC2_CASE
    return GFB_EUNSUPPORTED;
}

extern "C" int gfb_local_corr_tc2_f32(const float* f0, const float* f1, const float* flow, float* out,
                                      int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                                      int k_total, int k_offset, int group,
                                      void* workspace, size_t workspace_bytes, gfb_stream_t stream) {
    return local_corr_tc2(f0, f1, flow, out, B, C, Hs, Ws, f1_pitch, G, r, k_total, k_offset, group, 0, workspace, workspace_bytes, stream);
}
extern "C" int gfb_debug_local_corr_tc2_f32(const float* f0, const float* f1, const float* flow, float* out,
                                            int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                                            int k_total, int k_offset, int group, int debug,
                                            void* workspace, size_t workspace_bytes, gfb_stream_t stream) {
    return local_corr_tc2(f0, f1, flow, out, B, C, Hs, Ws, f1_pitch, G, r, k_total, k_offset, group, debug & 255, workspace, workspace_bytes, stream);
}

// ---- local correlation inside the refiner-input buffer (model/network.py:553-555) --------------------------
// d [B,Dtot,G,G]: feature0 = d[:, 0:C] (written by gfb_refiner_assemble_f32), corr -> d[:, k_offset : k_offset + (2r+1)^2].
// Dispatch: (r, C) of the point kernels -> gfb_local_corr_pt_f32's kernels; C = 32 -> mma.sync kernel; C = 64 -> tcgen05
// kernel (phase 0 = pre-pass + plan + main, 1 = pre-pass only, 2 = plan + main on a prepared workspace, as
// gfb_local_corr_tc2_prepare_f32 / _run_f32; workspace sized by gfb_local_corr_tc2_workspace_bytes) or, without a
// workspace, the mma.sync kernel.  phase 1 on a shape that needs no pre-pass is a no-op.
extern "C" int gfb_local_corr_cat_f32(float* d, int Dtot, const float* f1, const float* flow,
                                      int B, int C, int Hs, int Ws, int f1_pitch, int G, int r, int k_offset,
                                      int phase, void* workspace, size_t workspace_bytes, gfb_stream_t stream) {
    GFB_CHECK_ARG(d && Dtot >= C && phase >= 0 && phase <= 2);
    const bool pt_shape = (r == 2 && C == 16) || (r == 1 && C == 16) || (r == 1 && C == 8) || (r == 2 && C == 8);
    if (pt_shape && G % 4 == 0) {
        if (phase == 1) return GFB_OK;
        return local_corr_pt(d, f1, flow, d, B, C, Hs, Ws, f1_pitch, G, r, Dtot, k_offset, 0, 0, stream, Dtot);
    }
    LcParams p;
    float dummy;
    int rc = fill_params(p, d, f1, phase == 1 ? &dummy : flow, d, B, C, Hs, Ws, f1_pitch, G, r, Dtot, k_offset);
    if (rc != GFB_OK) return rc;
    p.f0_ctot = Dtot;
    cudaStream_t st = gfb_cu(stream);
    if (C == 64 && workspace) {
        if (phase == 1) { p.flow = nullptr; p.out = nullptr; }
#define GFB_TC2_CASE(RR, CC) if (r == RR && C == CC) return lcv2::launch_tc2<RR, CC>(p, st, workspace, workspace_bytes, 0, phase);
        GFB_TC2_ALL
#undef GFB_TC2_CASE
        return GFB_EUNSUPPORTED;
    }
    if (phase == 1) return GFB_OK;
    return lc_mma_launch(p, 0, st);
}
