// K1 -- local correlation (reference: utils/local_correlation.py:4-72, call site model/network.py:553).
//
// corr[b,k,gy,gx] = (1/sqrt C) sum_c f0[b,c,gy,gx] * bilinear(f1[b,c], flow[b,:,gy,gx] + off_k)
//
// Two kernels:
//  * lc_generic_kernel  -- one thread per output element, direct gathers.  Mirrors the reference's
//    coordinate arithmetic exactly (per-k fp32 `flow + offset`, unnormalise, floor) and covers every
//    mode the reference can be called with that we support (bilinear/nearest, zeros/border,
//    grid-based windows).  Slow path and in-kernel fallback.
//  * lc_stream_kernel   -- the hot kernel.  Window offsets are whole pixels (linspace step 2/w is
//    exactly one pixel under align_corners=False), so all K samples of a lattice point share one
//    fractional part and corr = bilerp(D) with D[j,i] = sum_c f0[c] * f1[c, y0-r+j, x0-r+i] over a
//    (2r+2)^2 integer patch.  A CTA owns a tile of lattice points; a producer warp streams the
//    rows of f1 the tile touches through a shared-memory ring with TMA (box = BW x 1 row x CCH
//    channels, out-of-image coordinates zero-filled by the TMA unit = padding_mode "zeros"); each
//    consumer thread owns P vertically adjacent lattice points, keeps one D row per point in
//    registers, reads a 128-bit-aligned superset segment of the f1 row once for its P points and
//    emits the bilinear-combined outputs row by row with coalesced streaming stores.
#include "common.cuh"
#include "lc_common.cuh"

namespace gfb {


__global__ void __launch_bounds__(256) lc_generic_kernel(LcParams p) {
    const int kk = (2 * p.r + 1) * (2 * p.r + 1);
    const size_t gg = (size_t)p.G * p.G;
    const size_t total = (size_t)p.B * kk * gg;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        int gx = (int)(t % p.G);
        int gy = (int)((t / p.G) % p.G);
        int k = (int)((t / gg) % kk);
        int b = (int)(t / (gg * kk));
        p.out[((size_t)b * p.k_total + p.k_offset + k) * gg + (size_t)gy * p.G + gx] = lc_generic_point(p, b, k, gy, gx);
    }
}

// ---------------------------------------------------------------------------------------------
// Hot kernel.
//   CCH  channels per shared-memory stage (C must be a multiple; C == CCH lets f0 live in registers)
//   R    window radius; W = 2R+2 columns/rows of the integer patch D, KW = 2R+1 outputs per row
//   P    vertically adjacent lattice points per thread (share one f1 segment load)
//   WP   padded segment width in floats (multiple of 4): W + alignment + shear slack
//   F0REG keep the thread's f0 vectors in registers (needs C == CCH and CCH*P <= 64)
struct LcTile {
    int TR, TC;    // lattice rows / cols per CTA tile (TR = P * NBR)
    int BW;        // smem row width in floats (multiple of 4, <= 256)
    int NST;       // ring stages
    int NCW;       // consumer warps
    int RPS;       // f1 rows per ring stage = per TMA operation (quad kernel)
};

// debug counters: [0] tiles, [1] tiles without any streamed point, [2] points sent to the gather path,
// [3] tiles whose box had to be centred (spread wider than BW)
// [4..7] (quad kernel, thread 0 of every CTA, SM clocks): set-up, wait for the first row, row loop, tail
__device__ unsigned long long g_lc_stats[8];

template <int CCH, int R, int P, int WP, bool F0REG>
__global__ void __launch_bounds__(P >= 4 ? 224 : 352, 1)
lc_stream_kernel(const LcParams p, const LcTile t, const __grid_constant__ CUtensorMap tmap_f1) {
    constexpr int W = 2 * R + 2;
    constexpr int KW = 2 * R + 1;
    constexpr int SHMAX = WP - W;
    static_assert(WP % 4 == 0 && SHMAX >= 3 && SHMAX <= 15, "segment must cover any 16B misalignment");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);
    const int stage_floats = CCH * t.BW;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)t.NST * stage_floats * sizeof(float));
    uint64_t* empty_bar = full_bar + t.NST;
    __shared__ int s_red[8];  // xmin, xmax, sum(base), n threads | ymin, ymax, n fast points, n slow points

    const int G = p.G;
    const int tiles_x = (G + t.TC - 1) / t.TC, tiles_y = (G + t.TR - 1) / t.TR;
    int tile = blockIdx.x;
    const int tcx = tile % tiles_x; tile /= tiles_x;
    const int tcy = tile % tiles_y;
    const int b = tile / tiles_y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_consumer = warp < t.NCW;
    const int NBR = t.TR / P;
    const size_t gg = (size_t)G * G;

    if (threadIdx.x == 0) {
        s_red[0] = INT_MAX; s_red[1] = INT_MIN; s_red[2] = 0; s_red[3] = 0;
        s_red[4] = INT_MAX; s_red[5] = INT_MIN; s_red[6] = 0; s_red[7] = 0;
        for (int s = 0; s < t.NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], t.NCW); }
        mbar_fence_init();
    }
    if (warp == t.NCW && lane == 0) tma_prefetch_desc(&tmap_f1);

    // ---- per-thread point setup -------------------------------------------------------------
    const int tid = threadIdx.x;
    const bool has_pts = is_consumer && tid < NBR * t.TC;
    const int rg = has_pts ? tid / t.TC : 0;
    const int gx = tcx * t.TC + (has_pts ? tid % t.TC : 0);
    int yb[P], sh[P];
    float wx1[P], wy0[P], wy1[P];
    bool valid[P], fast[P], slow[P];
    int basex = INT_MAX;
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int gy = tcy * t.TR + rg * P + q;
        valid[q] = has_pts && gy < G && gx < G;
        fast[q] = false; slow[q] = false; yb[q] = 0; sh[q] = 0; wx1[q] = 0.f; wy0[q] = 0.f; wy1[q] = 0.f;
        if (valid[q]) {
            const float* fl = p.flow + (size_t)b * 2 * gg + (size_t)gy * G + gx;
            float sx = unnormalize(__ldg(fl), p.Ws), sy = unnormalize(__ldg(fl + gg), p.Hs);
            if (fabsf(sx) < 1e6f && fabsf(sy) < 1e6f) {
                float x0f = floorf(sx), y0f = floorf(sy);
                int xb = (int)x0f - R;
                yb[q] = (int)y0f - R;
                float tx = sx - x0f, ty = sy - y0f;
                wx1[q] = tx;
                wy0[q] = (1.f - ty) * p.inv_sqrt_c;
                wy1[q] = ty * p.inv_sqrt_c;
                // a window entirely outside the image is all zeros (padding_mode "zeros"): neither fast nor slow
                fast[q] = !((xb >= p.Ws) || (xb + W <= 0) || (yb[q] >= p.Hs) || (yb[q] + W <= 0));
                sh[q] = xb;  // absolute for now
                if (fast[q]) basex = min(basex, xb);
            }
        }
    }
    bool any_live = basex != INT_MAX;
    if (any_live) {
        basex &= ~3;  // floor to a 16-byte boundary (two's complement: works for negatives)
#pragma unroll
        for (int q = 0; q < P; ++q) {
            if (fast[q]) {
                sh[q] -= basex;
                if (sh[q] > SHMAX) { fast[q] = false; slow[q] = true; }   // cannot share the thread's segment
            }
        }
    } else {
        basex = 0;
    }
    __syncthreads();  // s_red + barriers initialised
    {   // phase 1: where to put the staged box in x
        int xmn = warp_min(any_live ? basex : INT_MAX), xmx = warp_max(any_live ? basex + WP : INT_MIN);
        int sm = any_live ? basex : 0, cn = any_live ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sm += __shfl_xor_sync(0xffffffffu, sm, o); cn += __shfl_xor_sync(0xffffffffu, cn, o); }
        if (lane == 0 && cn) {
            atomicMin(&s_red[0], xmn); atomicMax(&s_red[1], xmx);
            atomicAdd(&s_red[2], sm); atomicAdd(&s_red[3], cn);
        }
    }
    __syncthreads();
    int xbox0 = 0;
    bool centred = false;
    if (s_red[3] > 0) {
        xbox0 = s_red[0] & ~3;
        if (s_red[1] - xbox0 > t.BW) {     // spread wider than the box: centre it on the mean segment
            const int mean = s_red[2] / s_red[3];   // (sum of ~hundreds of |base| < 1e6 fits in int)
            xbox0 = (mean + WP / 2 - t.BW / 2) & ~3;
            centred = true;
        }
    }
    if (any_live && (basex < xbox0 || basex + WP > xbox0 + t.BW)) {
#pragma unroll
        for (int q = 0; q < P; ++q) if (fast[q]) { fast[q] = false; slow[q] = true; }
    }
    {   // phase 2: rows to stream = union of the fast points' windows
        int ymin = INT_MAX, ymax = INT_MIN, nf = 0, ns = 0;
#pragma unroll
        for (int q = 0; q < P; ++q) {
            if (fast[q]) { ymin = min(ymin, yb[q]); ymax = max(ymax, yb[q] + W); ++nf; }
            ns += slow[q];
        }
        ymin = warp_min(ymin); ymax = warp_max(ymax);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { nf += __shfl_xor_sync(0xffffffffu, nf, o); ns += __shfl_xor_sync(0xffffffffu, ns, o); }
        if (lane == 0 && (nf | ns)) {
            if (nf) { atomicMin(&s_red[4], ymin); atomicMax(&s_red[5], ymax); }
            atomicAdd(&s_red[6], nf); atomicAdd(&s_red[7], ns);
        }
    }
    __syncthreads();
    const bool stream_any = s_red[6] > 0;
    const int ylo = s_red[4], yhi = s_red[5];
    if (threadIdx.x == 0) {
        atomicAdd(&g_lc_stats[0], 1ull);
        if (!stream_any) atomicAdd(&g_lc_stats[1], 1ull);
        if (s_red[7]) atomicAdd(&g_lc_stats[2], (unsigned long long)s_red[7]);
        if (centred) atomicAdd(&g_lc_stats[3], 1ull);
    }

    float* outb = p.out + ((size_t)b * p.k_total + p.k_offset) * gg;
    const int nchunk = p.C / CCH;

    // ---- producer warp: stream rows [ylo, yhi) ∩ [0, Hs) x channel chunks through the ring -----
    if (!is_consumer) {
        if (stream_any && warp == t.NCW && lane == 0) {
            const uint32_t bytes = (uint32_t)stage_floats * sizeof(float);
            int seq = 0;
            for (int y = max(ylo, 0); y < min(yhi, p.Hs); ++y) {
                for (int ch = 0; ch < nchunk; ++ch, ++seq) {
                    const int s = seq % t.NST;
                    if (seq >= t.NST) mbar_wait(&empty_bar[s], ((seq / t.NST) - 1) & 1);
                    mbar_expect_tx(&full_bar[s], bytes);
                    tma_load_3d(ring + (size_t)s * stage_floats, &tmap_f1, &full_bar[s], xbox0, y, b * p.C + ch * CCH);
                }
            }
        }
        return;
    }

    // ---- consumer threads ----------------------------------------------------------------------
    if (stream_any) {
        float acc[P][WP], hprev[P][KW];
#pragma unroll
        for (int q = 0; q < P; ++q) {
#pragma unroll
            for (int i = 0; i < WP; ++i) acc[q][i] = 0.f;
#pragma unroll
            for (int i = 0; i < KW; ++i) hprev[q][i] = 0.f;
        }
        const float* f0p[P];
        float f0r[F0REG ? P : 1][F0REG ? CCH : 1];
#pragma unroll
        for (int q = 0; q < P; ++q) {
            const int gy = min(tcy * t.TR + rg * P + q, G - 1);
            f0p[q] = p.f0 + (size_t)b * p.C * gg + (size_t)gy * G + min(gx, G - 1);
            if (F0REG) {
#pragma unroll
                for (int c = 0; c < CCH; ++c) f0r[F0REG ? q : 0][F0REG ? c : 0] = fast[q] ? __ldg(f0p[q] + (size_t)c * gg) : 0.f;
            }
        }
        const int xoff = max(basex - xbox0, 0);  // multiple of 4 floats (only used by fast points)
        int seq = 0;
        for (int y = ylo; y < yhi; ++y) {
            bool act[P];
            bool any_act = false;
#pragma unroll
            for (int q = 0; q < P; ++q) { act[q] = fast[q] && (unsigned)(y - yb[q]) < (unsigned)W; any_act |= act[q]; }
            if ((unsigned)y < (unsigned)p.Hs) {
                for (int ch = 0; ch < nchunk; ++ch, ++seq) {
                    const int s = seq % t.NST;
                    // every consumer warp waits on every stage (even rows it skips): that bounds how far
                    // a warp can run ahead and keeps its empty-barrier arrivals in the right phase
                    mbar_wait(&full_bar[s], (seq / t.NST) & 1);
                    if (any_act) {
                        const float* srow = ring + (size_t)s * stage_floats + xoff;
#pragma unroll (F0REG ? CCH : 4)
                        for (int c = 0; c < CCH; ++c) {
                            float seg[WP];
#pragma unroll
                            for (int v = 0; v < WP / 4; ++v) {
                                float4 u = *reinterpret_cast<const float4*>(srow + c * t.BW + 4 * v);
                                seg[4 * v] = u.x; seg[4 * v + 1] = u.y; seg[4 * v + 2] = u.z; seg[4 * v + 3] = u.w;
                            }
#pragma unroll
                            for (int q = 0; q < P; ++q) {
                                if (act[q]) {
                                    const float f = F0REG ? f0r[F0REG ? q : 0][F0REG ? c : 0]
                                                          : __ldg(f0p[q] + (size_t)(ch * CCH + c) * gg);
#pragma unroll
                                    for (int i = 0; i < WP; ++i) acc[q][i] = fmaf(f, seg[i], acc[q][i]);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[s]);
                }
            }
            // ---- finish D row j = y - yb for every active point: shift, lerp x, lerp y, store -----
#pragma unroll
            for (int q = 0; q < P; ++q) {
                if (!act[q]) continue;
                const int j = y - yb[q];
                if (SHMAX >= 8 && (sh[q] & 8)) {
#pragma unroll
                    for (int i = 0; i + 8 < WP; ++i) acc[q][i] = acc[q][i + 8];
                }
                if (SHMAX >= 4 && (sh[q] & 4)) {
#pragma unroll
                    for (int i = 0; i + 4 < WP; ++i) acc[q][i] = acc[q][i + 4];
                }
                if (sh[q] & 2) {
#pragma unroll
                    for (int i = 0; i + 2 < WP; ++i) acc[q][i] = acc[q][i + 2];
                }
                if (sh[q] & 1) {
#pragma unroll
                    for (int i = 0; i + 1 < WP; ++i) acc[q][i] = acc[q][i + 1];
                }
                const float a1 = wx1[q], a0 = 1.f - a1;
                float* o = outb + ((size_t)(j - 1) * KW) * gg + (size_t)(tcy * t.TR + rg * P + q) * G + gx;
#pragma unroll
                for (int i = 0; i < KW; ++i) {
                    const float h = a0 * acc[q][i] + a1 * acc[q][i + 1];
                    if (j >= 1) st_stream(o + (size_t)i * gg, wy0[q] * hprev[q][i] + wy1[q] * h);
                    hprev[q][i] = h;
                }
#pragma unroll
                for (int i = 0; i < WP; ++i) acc[q][i] = 0.f;
            }
        }
    }
    // points that did not stream: window entirely outside the image -> zeros; segment outside the staged
    // box (wild flow) -> exact per-sample gathers
#pragma unroll
    for (int q = 0; q < P; ++q) {
        if (valid[q] && !fast[q]) {
            const int gy = tcy * t.TR + rg * P + q;
            float* o = outb + (size_t)gy * G + gx;
            for (int k = 0; k < KW * KW; ++k)
                st_stream(o + (size_t)k * gg, slow[q] ? lc_generic_point(p, b, k, gy, gx) : 0.f);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Quad kernel: the hot kernel of the 448/560 path (C = 16 * CS).
//
// The P = 1 kernel above reads every f1 value from shared memory once per lattice point that uses it:
// one LDS.128 feeds 4 FFMA, and the shared-memory pipe (128 B/clk/SM) saturates at a quarter of the FP32
// rate (ncu: 35-50 % LSU wavefronts, 25 % FMA pipe).  Here a lane owns a QUAD of 4 vertically adjacent
// lattice points: their windows cover (nearly) the same columns, so one segment load feeds 4 points, and
// the multiply-adds are packed fma.rn.f32x2 (FFMA2: two FMAs per issue slot).  The lane keeps the f0
// values of its 4 points x 16 channels in registers; for C = 32 / 64, CS = 2 / 4 neighbouring lanes split
// the channels of one quad, and at the end of every streamed row a reduce-scatter over those lanes hands
// each lane the complete row of D for the 4 / CS points it finishes (shift, lerp x, lerp y, store).
//   VEC  floats per shared-memory load (4: LDS.128, 16 B aligned segments; 2: LDS.64, 8 B aligned)
//   WS   segment width in floats = W + SHMAX: alignment slack (VEC - 1) + spread of the 4 points' window
//        starts (shear of the flow over 3 lattice rows + per-point jitter)
// d (two packed fp32) += a * b (two packed fp32); the accumulators and the shared-memory segments stay in
// 64-bit registers so that FFMA2 needs no packing moves
__device__ __forceinline__ void ffma2(unsigned long long& d, float a, unsigned long long b) {
    asm("{\n\t.reg .b64 aa;\n\tmov.b64 aa, {%1, %1};\n\tfma.rn.f32x2 %0, aa, %2, %0;\n\t}" : "+l"(d) : "f"(a), "l"(b));
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

template <int N>
__device__ __forceinline__ float sel4(const float (&v)[N], int i) {  // v[i] for a lane-dependent i < N <= 4
    float r = v[0];
#pragma unroll
    for (int k = 1; k < N; ++k) r = (i == k) ? v[k] : r;
    return r;
}
template <int N>
__device__ __forceinline__ int sel4(const int (&v)[N], int i) {
    int r = v[0];
#pragma unroll
    for (int k = 1; k < N; ++k) r = (i == k) ? v[k] : r;
    return r;
}

template <int R, int CS, int VEC, int WS, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
lc_quad_kernel(const LcParams p, const LcTile t, const __grid_constant__ CUtensorMap tmap_f1) {
    constexpr int P = 4;
    constexpr int W = 2 * R + 2;
    constexpr int KW = 2 * R + 1;
    constexpr int SHMAX = WS - W;
    constexpr int CPL = 16;            // channels per lane
    constexpr int NOWN = P / CS;       // points a lane finishes
    constexpr int C = CPL * CS;
    constexpr int H2 = WS / 2;
    static_assert(WS % VEC == 0 && SHMAX >= VEC - 1 && SHMAX <= 15 && (VEC == 2 || VEC == 4), "bad segment");
    static_assert(CS == 1 || CS == 2 || CS == 4, "bad channel split");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);
    const int stage_floats = t.RPS * C * t.BW;   // [RPS rows][C channels][BW]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)t.NST * stage_floats * sizeof(float));
    uint64_t* empty_bar = full_bar + t.NST;
    __shared__ int s_red[8];
    __shared__ int s_nslow;
    __shared__ unsigned s_slow[1024];   // (gy << 16 | gx) of the tile's points that take the gather path (<= 4 * 256 / CS)

    const long long t_start = clock64();
    const int G = p.G;
    const int tiles_x = (G + t.TC - 1) / t.TC, tiles_y = (G + t.TR - 1) / t.TR;
    int tile = blockIdx.x;
    const int tcx = tile % tiles_x; tile /= tiles_x;
    const int tcy = tile % tiles_y;
    const int b = tile / tiles_y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NBR = t.TR / P;
    const size_t gg = (size_t)G * G;

    if (threadIdx.x == 0) {
        s_red[0] = INT_MAX; s_red[1] = INT_MIN; s_red[2] = 0; s_red[3] = 0;
        s_red[4] = INT_MAX; s_red[5] = INT_MIN; s_red[6] = 0; s_red[7] = 0; s_nslow = 0;
        for (int s = 0; s < t.NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], t.NCW); }
        mbar_fence_init();
    }
    if (threadIdx.x == 0) tma_prefetch_desc(&tmap_f1);

    // ---- per-lane quad setup (the CS lanes of a quad compute the same thing) ------------------------
    const int tid = threadIdx.x;
    const int cs = tid % CS;
    const int col = (tid / CS) % t.TC;
    const int rg = tid / (CS * t.TC);
    const bool has_pts = rg < NBR;
    const int gx = tcx * t.TC + col;
    const int gy0 = tcy * t.TR + rg * P;
    // f0 of the quad (this lane's 16 channels: CS * c + cs): issued first, consumed after the set-up phases
    float f0r[P][CPL];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const float* f0p = p.f0 + ((size_t)b * C + cs) * gg + (size_t)min(gy0 + q, G - 1) * G + min(gx, G - 1);
#pragma unroll
        for (int c = 0; c < CPL; ++c) f0r[q][c] = __ldg(f0p + (size_t)(c * CS) * gg);
    }
    int yb[P], sh[P];
    float wx1[P], wy0[P], wy1[P];
    bool fast[P], slow[P], valid[P];
    int basex = INT_MAX;
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int gy = gy0 + q;
        valid[q] = has_pts && gy < G && gx < G;
        fast[q] = false; slow[q] = false; yb[q] = 0; sh[q] = 0; wx1[q] = 0.f; wy0[q] = 0.f; wy1[q] = 0.f;
        if (valid[q]) {
            const float* fl = p.flow + (size_t)b * 2 * gg + (size_t)gy * G + gx;
            float sx = unnormalize(__ldg(fl), p.Ws), sy = unnormalize(__ldg(fl + gg), p.Hs);
            if (fabsf(sx) < 1e6f && fabsf(sy) < 1e6f) {
                float x0f = floorf(sx), y0f = floorf(sy);
                int xb = (int)x0f - R;
                yb[q] = (int)y0f - R;
                float tx = sx - x0f, ty = sy - y0f;
                wx1[q] = tx;
                wy0[q] = (1.f - ty) * p.inv_sqrt_c;
                wy1[q] = ty * p.inv_sqrt_c;
                fast[q] = !((xb >= p.Ws) || (xb + W <= 0) || (yb[q] >= p.Hs) || (yb[q] + W <= 0));
                sh[q] = xb;
                if (fast[q]) basex = min(basex, xb);
            }
        }
    }
    bool any_live = basex != INT_MAX;
    if (any_live) {
        basex &= ~(VEC - 1);
#pragma unroll
        for (int q = 0; q < P; ++q) {
            if (fast[q]) {
                sh[q] -= basex;
                if (sh[q] > SHMAX) { fast[q] = false; slow[q] = true; }
            }
        }
    } else {
        basex = 0;
    }
    __syncthreads();
    {   // where to put the staged box in x
        int xmn = warp_min(any_live ? basex : INT_MAX), xmx = warp_max(any_live ? basex + WS : INT_MIN);
        int sm = any_live ? basex : 0, cn = any_live ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sm += __shfl_xor_sync(0xffffffffu, sm, o); cn += __shfl_xor_sync(0xffffffffu, cn, o); }
        if (lane == 0 && cn) {
            atomicMin(&s_red[0], xmn); atomicMax(&s_red[1], xmx);
            atomicAdd(&s_red[2], sm); atomicAdd(&s_red[3], cn);
        }
    }
    __syncthreads();
    int xbox0 = 0;
    bool centred = false;
    if (s_red[3] > 0) {
        xbox0 = s_red[0] & ~3;
        if (s_red[1] - xbox0 > t.BW) {
            const int mean = s_red[2] / s_red[3];
            xbox0 = (mean + WS / 2 - t.BW / 2) & ~3;
            centred = true;
        }
    }
    if (any_live && (basex < xbox0 || basex + WS > xbox0 + t.BW)) {
#pragma unroll
        for (int q = 0; q < P; ++q) if (fast[q]) { fast[q] = false; slow[q] = true; }
    }
    {   // rows to stream = union of the streamed points' windows
        int ymin = INT_MAX, ymax = INT_MIN, nf = 0, ns = 0;
#pragma unroll
        for (int q = 0; q < P; ++q) {
            if (fast[q]) { ymin = min(ymin, yb[q]); ymax = max(ymax, yb[q] + W); ++nf; }
            ns += slow[q];
        }
        if (cs != 0) { nf = 0; ns = 0; }   // count each point once
        ymin = warp_min(ymin); ymax = warp_max(ymax);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { nf += __shfl_xor_sync(0xffffffffu, nf, o); ns += __shfl_xor_sync(0xffffffffu, ns, o); }
        if (lane == 0 && (nf | ns)) {
            if (nf) { atomicMin(&s_red[4], ymin); atomicMax(&s_red[5], ymax); }
            atomicAdd(&s_red[6], nf); atomicAdd(&s_red[7], ns);
        }
    }
    __syncthreads();
    const bool stream_any = s_red[6] > 0;
    const int ylo = s_red[4], yhi = s_red[5];
    if (threadIdx.x == 0) {
        atomicAdd(&g_lc_stats[0], 1ull);
        if (!stream_any) atomicAdd(&g_lc_stats[1], 1ull);
        if (s_red[7]) atomicAdd(&g_lc_stats[2], (unsigned long long)s_red[7]);
        if (centred) atomicAdd(&g_lc_stats[3], 1ull);
    }
    float* outb = p.out + ((size_t)b * p.k_total + p.k_offset) * gg;

    // ---- TMA producer = lane 0 of warp 0 (no dedicated warp: registers are allocated per warp slot) ----------
    const bool producer = threadIdx.x == 0;
    // One TMA operation per stage = RPS consecutive rows (an SM sustains only ~3 M TMA operations/s per CTA, so the
    // operations must be >= ~32 KB to keep up with HBM: tools/micro/tma_stream.cu).
    const int y_first = max(ylo, 0), n_rows = stream_any ? max(min(yhi, p.Hs) - y_first, 0) : 0;
    const int n_groups = (n_rows + t.RPS - 1) / t.RPS;
    const uint32_t stage_bytes = (uint32_t)stage_floats * sizeof(float);
    int p_stage = 0;                         // producer: stage of the next row group to issue
    int e_stage = 0; uint32_t e_phase = 0;   // producer: stage / phase of the next empty barrier to wait on
    int n_issued = 0;
    auto issue_group = [&]() {               // next RPS rows -> next stage (the stage must be free); rows >= Hs are zero-filled
        mbar_expect_tx(&full_bar[p_stage], stage_bytes);
        tma_load_3d(ring + (size_t)p_stage * stage_floats, &tmap_f1, &full_bar[p_stage], xbox0, b * C, y_first + n_issued * t.RPS);
        ++n_issued;
        if (++p_stage == t.NST) p_stage = 0;
    };
    if (producer)
        for (int i = 0; i < min(n_groups, t.NST); ++i) issue_group();

    // ---- the points this lane finishes -----------------------------------------------------------------
    int o_q[NOWN], o_sh[NOWN], o_yb[NOWN];
    float o_wx1[NOWN], o_wy0[NOWN], o_wy1[NOWN];
    bool o_fast[NOWN], o_slow[NOWN], o_valid[NOWN];
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
        const int q = (CS == 1) ? o : (CS == 2 ? cs * 2 + o : cs);
        o_q[o] = q;
        o_sh[o] = sel4(sh, q); o_yb[o] = sel4(yb, q);
        o_wx1[o] = sel4(wx1, q); o_wy0[o] = sel4(wy0, q); o_wy1[o] = sel4(wy1, q);
        int fl[P], sl[P], va[P];
#pragma unroll
        for (int k = 0; k < P; ++k) { fl[k] = fast[k]; sl[k] = slow[k]; va[k] = valid[k]; }
        o_fast[o] = sel4(fl, q) != 0; o_slow[o] = sel4(sl, q) != 0; o_valid[o] = sel4(va, q) != 0;
    }

    const long long t_setup = clock64();
    long long t_first = t_setup, t_loop = t_setup;
    if (stream_any) {
        unsigned long long acc[P][H2];
        float hprev[NOWN][KW];
#pragma unroll
        for (int q = 0; q < P; ++q) {
#pragma unroll
            for (int m = 0; m < H2; ++m) acc[q][m] = 0ull;
        }
#pragma unroll
        for (int o = 0; o < NOWN; ++o)
#pragma unroll
            for (int i = 0; i < KW; ++i) hprev[o][i] = 0.f;

        // lanes without streamed points still execute the warp's loads: keep their segment inside the stage
        const int xoff = (fast[0] || fast[1] || fast[2] || fast[3]) ? basex - xbox0 : 0;
        int seq = 0, s = 0, rin = 0;   // row group, its stage, row inside the group
        uint32_t ph = 0;
        for (int y = ylo; y < yhi; ++y) {
            unsigned wm = 0;   // which of the 4 points of the quads in this warp are inside their window rows
#pragma unroll
            for (int q = 0; q < P; ++q) {
                const bool a = fast[q] && (unsigned)(y - yb[q]) < (unsigned)W;
                if (__any_sync(0xffffffffu, a)) wm |= 1u << q;
            }
            if ((unsigned)y < (unsigned)p.Hs) {
                if (rin == 0) {
                    if (producer) {
                        // refill every stage that all warps have released; block only for the group this warp needs now
                        while (n_issued < n_groups) {
                            if (n_issued >= t.NST) {
                                if (n_issued == seq) mbar_wait(&empty_bar[e_stage], e_phase);
                                else if (!mbar_try_wait(&empty_bar[e_stage], e_phase)) break;
                                if (++e_stage == t.NST) { e_stage = 0; e_phase ^= 1; }
                            }
                            issue_group();
                        }
                    }
                    __syncwarp();
                    mbar_wait(&full_bar[s], ph);
                }
                if (seq == 0) t_first = clock64();
                if (wm && !(p.debug & 1)) {
                    // channels are interleaved over the CS lanes of a quad (lane cs: CS * c + cs) and BW = 8 mod 32,
                    // so the lanes of a quad read different banks
                    const float* srow = ring + (size_t)s * stage_floats + (size_t)(rin * C + cs) * t.BW + xoff;
                    const int cstride = CS * t.BW;
#pragma unroll
                    for (int c = 0; c < CPL; c += 2) {
                        unsigned long long sa[H2], sb[H2];
                        if (VEC == 4) {
#pragma unroll
                            for (int v = 0; v < WS / 4; ++v) {
                                const ulonglong2 u = *reinterpret_cast<const ulonglong2*>(srow + c * cstride + 4 * v);
                                const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(srow + (c + 1) * cstride + 4 * v);
                                sa[2 * v] = u.x; sa[2 * v + 1] = u.y;
                                sb[2 * v] = w.x; sb[2 * v + 1] = w.y;
                            }
                        } else {
#pragma unroll
                            for (int v = 0; v < H2; ++v) {
                                sa[v] = *reinterpret_cast<const unsigned long long*>(srow + c * cstride + 2 * v);
                                sb[v] = *reinterpret_cast<const unsigned long long*>(srow + (c + 1) * cstride + 2 * v);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < P; ++q) {
                            if (wm & (1u << q)) {
#pragma unroll
                                for (int m = 0; m < H2; ++m) { ffma2(acc[q][m], f0r[q][c], sa[m]); ffma2(acc[q][m], f0r[q][c + 1], sb[m]); }
                            }
                        }
                    }
                }
                if (++rin == t.RPS || y + 1 >= min(yhi, p.Hs)) {   // last row of the group: release the stage
                    rin = 0;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[s]);
                    ++seq;
                    if (++s == t.NST) { s = 0; ph ^= 1; }
                }
            }
            if (!wm) continue;
            // ---- row y is complete: reduce over the CS lanes of the quad, each lane keeps its own points ----
            float red[CS == 1 ? 1 : NOWN][WS];
            if constexpr (CS == 1) {
                // each lane finishes its own 4 points straight from the accumulators (below)
            } else if constexpr (CS == 2) {
                const bool hi = cs & 1;
#pragma unroll
                for (int o = 0; o < 2; ++o)
#pragma unroll
                    for (int m = 0; m < H2; ++m) {
                        const float2 keep = unpack2(hi ? acc[2 + o][m] : acc[o][m]);
                        const float2 send = unpack2(hi ? acc[o][m] : acc[2 + o][m]);
                        red[o][2 * m] = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 1);
                        red[o][2 * m + 1] = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 1);
                    }
            } else {
                const bool hi2 = cs & 2, hi1 = cs & 1;
#pragma unroll
                for (int m = 0; m < H2; ++m) {
                    float2 k0 = unpack2(hi2 ? acc[2][m] : acc[0][m]), k1 = unpack2(hi2 ? acc[3][m] : acc[1][m]);
                    const float2 s0 = unpack2(hi2 ? acc[0][m] : acc[2][m]), s1 = unpack2(hi2 ? acc[1][m] : acc[3][m]);
                    k0.x += __shfl_xor_sync(0xffffffffu, s0.x, 2); k0.y += __shfl_xor_sync(0xffffffffu, s0.y, 2);
                    k1.x += __shfl_xor_sync(0xffffffffu, s1.x, 2); k1.y += __shfl_xor_sync(0xffffffffu, s1.y, 2);
                    const float2 keep = hi1 ? k1 : k0, send = hi1 ? k0 : k1;
                    red[0][2 * m] = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 1);
                    red[0][2 * m + 1] = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 1);
                }
            }
            if constexpr (CS > 1) {
#pragma unroll
                for (int q = 0; q < P; ++q)
                    if (wm & (1u << q)) {
#pragma unroll
                        for (int m = 0; m < H2; ++m) acc[q][m] = 0ull;
                    }
            }
            // ---- finish D row j = y - yb of the owned points: shift, lerp x, lerp y, store ---------------------
#pragma unroll
            for (int o = 0; o < NOWN; ++o) {
                const int j = y - o_yb[o];
                float rloc[CS == 1 ? WS : 1];
                if constexpr (CS == 1) {
#pragma unroll
                    for (int m = 0; m < H2; ++m) { const float2 v = unpack2(acc[o][m]); rloc[2 * m] = v.x; rloc[2 * m + 1] = v.y; acc[o][m] = 0ull; }
                }
                if (!(o_fast[o] && (unsigned)j < (unsigned)W)) continue;
                float* rr = CS == 1 ? rloc : red[CS == 1 ? 0 : o];
                if (SHMAX >= 8 && (o_sh[o] & 8)) {
#pragma unroll
                    for (int i = 0; i + 8 < WS; ++i) rr[i] = rr[i + 8];
                }
                if (SHMAX >= 4 && (o_sh[o] & 4)) {
#pragma unroll
                    for (int i = 0; i + 4 < WS; ++i) rr[i] = rr[i + 4];
                }
                if (SHMAX >= 2 && (o_sh[o] & 2)) {
#pragma unroll
                    for (int i = 0; i + 2 < WS; ++i) rr[i] = rr[i + 2];
                }
                if (o_sh[o] & 1) {
#pragma unroll
                    for (int i = 0; i + 1 < WS; ++i) rr[i] = rr[i + 1];
                }
                const float a1 = o_wx1[o], a0 = 1.f - a1;
                float* op = outb + ((size_t)(j - 1) * KW) * gg + (size_t)(gy0 + o_q[o]) * G + gx;
#pragma unroll
                for (int i = 0; i < KW; ++i) {
                    const float h = a0 * rr[i] + a1 * rr[i + 1];
                    if (j >= 1 && !(p.debug & 2)) st_stream(op + (size_t)i * gg, o_wy0[o] * hprev[o][i] + o_wy1[o] * h);
                    hprev[o][i] = h;
                }
            }
        }
    }
    t_loop = clock64();
    // points that did not stream: window outside the image -> zeros; segment outside the staged box or quad
    // spread wider than the segment -> exact per-sample gathers, shared out over the whole CTA
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
        if (o_valid[o] && !o_fast[o]) {
            const int gy = gy0 + o_q[o];
            if (o_slow[o]) {
                const int slot = atomicAdd(&s_nslow, 1);
                s_slow[slot] = (unsigned)((gy << 16) | gx);
            } else {
                float* op = outb + (size_t)gy * G + gx;
                for (int k = 0; k < KW * KW; ++k) st_stream(op + (size_t)k * gg, 0.f);
            }
        }
    }
    __syncthreads();
    const int nslow = s_nslow;
    for (int e = threadIdx.x; e < nslow * KW * KW; e += blockDim.x) {
        const int k = e / nslow, pt = e - k * nslow;
        const int gy = (int)(s_slow[pt] >> 16), gxx = (int)(s_slow[pt] & 0xffffu);
        st_stream(outb + (size_t)k * gg + (size_t)gy * G + gxx, lc_generic_point(p, b, k, gy, gxx));
    }
    if (threadIdx.x == 0) {
        atomicAdd(&g_lc_stats[4], (unsigned long long)(t_setup - t_start));
        atomicAdd(&g_lc_stats[5], (unsigned long long)(t_first - t_setup));
        atomicAdd(&g_lc_stats[6], (unsigned long long)(t_loop - t_first));
        atomicAdd(&g_lc_stats[7], (unsigned long long)(clock64() - t_loop));
    }
}

__global__ void avg_pool2_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W) {
    const int Ho = H / 2, Wo = W / 2;
    const size_t total = (size_t)N * Ho * Wo;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        int xo = (int)(t % Wo), yo = (int)((t / Wo) % Ho);
        size_t n = t / ((size_t)Wo * Ho);
        const float* q = x + n * H * W + (size_t)(2 * yo) * W + 2 * xo;
        y[t] = (q[0] + q[1] + q[W] + q[W + 1]) * 0.25f;
    }
}

__global__ void pad_rows_kernel(const float* __restrict__ x, float* __restrict__ y, size_t rows, int W, int pitch) {
    const size_t total = rows * pitch;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int xx = (int)(t % pitch);
        const size_t r = t / pitch;
        y[t] = xx < W ? x[r * W + xx] : 0.f;
    }
}

template <int CCH, int R, int P, int WP, bool F0REG>
static int launch_stream(const LcParams& p, cudaStream_t st) {
    LcTile t;
    const int G = p.G;
    const float s = (float)p.Ws / (float)G;
    const int bw_full = ((p.Ws + 2 * WP + 3) / 4) * 4;
    // tile columns: whole lattice rows up to 80 wide, else a divisor of G; shrink until the staged box
    // (tile span x 1.6 slack for local magnification + segment + alignment) fits a 256-wide TMA box
    const int cand[] = {80, 64, 48, 40, 32, 24, 16, 8};
    t.TC = 0; t.BW = 0;
    for (int ci = -1; ci < 8 && !t.TC; ++ci) {
        const int tc = ci < 0 ? (G <= 80 ? G : 0) : cand[ci];
        if (tc <= 0 || tc > G || (G % tc != 0 && !(ci == 7))) continue;
        int bw = (((int)ceilf((float)tc * s * 1.6f) + WP + 8 + 3) / 4) * 4;
        bw = min(bw, bw_full);
        if (bw <= 256) { t.TC = tc; t.BW = bw; }
    }
    if (!t.TC) return GFB_EUNSUPPORTED;
    int NBR = max(1, min(8 / P, (P >= 4 ? 192 : 320) / t.TC));
    t.TR = NBR * P;
    t.NCW = (NBR * t.TC + 31) / 32;
    const size_t stage_bytes = (size_t)CCH * t.BW * sizeof(float);
    t.NST = (int)min((size_t)6, (size_t)(200 * 1024) / stage_bytes);
    if (t.NST < 2) return GFB_EUNSUPPORTED;
    // keep two CTAs per SM resident when the ring allows it
    if ((size_t)t.NST * stage_bytes > 100 * 1024) t.NST = (int)max((size_t)3, (size_t)(100 * 1024) / stage_bytes);
    const size_t smem = (size_t)t.NST * stage_bytes + 2 * t.NST * sizeof(uint64_t);

    CUtensorMap tmap;
    uint64_t dims[3] = {(uint64_t)p.Ws, (uint64_t)p.Hs, (uint64_t)p.B * p.C};
    uint64_t strides[2] = {(uint64_t)p.pitch * 4, (uint64_t)p.Hs * p.pitch * 4};
    uint32_t box[3] = {(uint32_t)t.BW, 1u, (uint32_t)CCH};
    int rc = gfb_encode_tmap_f32(&tmap, p.f1, 3, dims, strides, box, 0);
    if (rc != GFB_OK) return rc;

    auto kern = lc_stream_kernel<CCH, R, P, WP, F0REG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int tiles = p.B * ((G + t.TR - 1) / t.TR) * ((G + t.TC - 1) / t.TC);
    kern<<<tiles, (t.NCW + 1) * 32, smem, st>>>(p, t, tmap);
    GFB_LAUNCH_RESULT();
}

template <int R, int CS, int VEC, int WS>
static int launch_quad(const LcParams& p, cudaStream_t st, int div_override, int nbr_override, int nst_override, int rps_override) {
    constexpr int P = 4;
    constexpr int C = 16 * CS;
    LcTile t;
    const int G = p.G;
    const float s = (float)p.Ws / (float)G;
    const int bw_full = ((p.Ws + 2 * WS + 3) / 4) * 4;
    const int quad_rows = (G + P - 1) / P;
    // Tile = NBR quad rows x TC lattice columns.  Measured on B200 (tools/bench_kernels.py sweep, profiles/): small
    // tiles win -- one quad row of <= 32 lattice columns per CTA (1 warp at C = 16, 2 at C = 32, 4 at C = 64), several
    // CTAs per SM: the warps of a CTA advance through the row ring in lockstep, and the rows a warp needs shift with x
    // when the flow rotates, so wide or tall tiles leave most warps waiting.  TC = the largest divisor of G that is
    // <= 32 and a multiple of 4 (else G itself); the staged box = tile span x 1.6 slack for local magnification
    // + segment + alignment must fit a 256-wide TMA box.
    t.TC = 0; t.BW = 0; t.TR = 0;
    for (int div = 1; div <= 16; ++div) {
        if (G % div || (div_override && div != div_override)) continue;
        const int tc = G / div;
        if (tc * CS > 256) continue;
        if (!div_override && (tc > 32 || tc % 4) && div < 16 && G > 32) continue;
        int bw = (((int)ceilf((float)tc * s * 1.6f) + WS + 8 + 3) / 4) * 4;
        bw = min(bw, bw_full);
        if (CS > 1) bw += (8 - bw % 32 + 32) % 32;     // row pitch = 8 banks mod 32: see the channel interleave
        if (bw > 256 || (size_t)C * bw * sizeof(float) * 2 > 200 * 1024) continue;
        int nbr = nbr_override ? nbr_override : 1;
        nbr = max(1, min(nbr, min(quad_rows, 256 / (tc * CS))));
        t.TC = tc; t.BW = bw; t.TR = nbr * P;
        break;
    }
    if (!t.TC) return GFB_EUNSUPPORTED;
    const int NBR = t.TR / P;
    t.NCW = (NBR * t.TC * CS + 31) / 32;
    const size_t row_bytes = (size_t)C * t.BW * sizeof(float);
    // registers (<= 255 per lane) allow 8 warps per SM: CTAs of NCW warps run 8 / NCW at a time and share the 227 KB
    const int ctas_per_sm = max(1, 8 / t.NCW);
    const size_t budget = (size_t)(220 * 1024) / ctas_per_sm - 1024;
    t.RPS = rps_override ? rps_override : (int)max((size_t)1, min((size_t)4, (size_t)(32 * 1024 + row_bytes / 2) / row_bytes));
    while (t.RPS > 1 && 3 * t.RPS * row_bytes > budget) --t.RPS;
    const size_t stage_bytes = (size_t)t.RPS * row_bytes;
    t.NST = (int)min((size_t)(nst_override ? nst_override : 4), budget / stage_bytes);
    if (t.NST < 2) return GFB_EUNSUPPORTED;
    const size_t smem = (size_t)t.NST * stage_bytes + 2 * t.NST * sizeof(uint64_t);

    // f1 as a 3-D tensor (x, plane, row): a box {BW, C, RPS} lands in shared memory as [row][channel][x]
    CUtensorMap tmap;
    uint64_t dims[3] = {(uint64_t)p.Ws, (uint64_t)p.B * p.C, (uint64_t)p.Hs};
    uint64_t strides[2] = {(uint64_t)p.Hs * p.pitch * 4, (uint64_t)p.pitch * 4};
    uint32_t box[3] = {(uint32_t)t.BW, (uint32_t)C, (uint32_t)t.RPS};
    int rc = gfb_encode_tmap_f32(&tmap, p.f1, 3, dims, strides, box, 0);
    if (rc != GFB_OK) return rc;

    const int tiles = p.B * ((G + t.TR - 1) / t.TR) * ((G + t.TC - 1) / t.TC);
    cudaError_t e;
    if (t.NCW <= 4) {
        auto kern = lc_quad_kernel<R, CS, VEC, WS, 128, 2>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<tiles, t.NCW * 32, smem, st>>>(p, t, tmap);
    } else {
        auto kern = lc_quad_kernel<R, CS, VEC, WS, 256, 1>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        kern<<<tiles, t.NCW * 32, smem, st>>>(p, t, tmap);
    }
    GFB_LAUNCH_RESULT();
}

// quad kernel dispatch: C = 16 * CS; segment = W + 6 (LDS.128) or W + 4 (LDS.64)
template <int VEC>
static int launch_quad_any(const LcParams& p, cudaStream_t st, int div_override, int nbr_override, int nst_override, int rps_override) {
    constexpr int SL = VEC == 4 ? 6 : 4;
#define GFB_QUAD_CASE(RR, CSV) \
    if (p.r == RR && p.C == 16 * CSV) \
        return launch_quad<RR, CSV, VEC, (2 * RR + 2 + SL + VEC - 1) / VEC * VEC>(p, st, div_override, nbr_override, nst_override, rps_override);
    GFB_QUAD_CASE(2, 1) GFB_QUAD_CASE(4, 2) GFB_QUAD_CASE(6, 4) GFB_QUAD_CASE(7, 4)
#undef GFB_QUAD_CASE
    return GFB_EUNSUPPORTED;
}

// one point per thread, minimal segment: the default (robust to per-point flow jitter)
template <int CCH, bool F0REG>
static int launch_stream_p1(const LcParams& p, cudaStream_t st) {
    switch (p.r) {
        case 1: return launch_stream<CCH, 1, 1, 8, F0REG>(p, st);
        case 2: return launch_stream<CCH, 2, 1, 12, F0REG>(p, st);
        case 3: return launch_stream<CCH, 3, 1, 12, F0REG>(p, st);
        case 4: return launch_stream<CCH, 4, 1, 16, F0REG>(p, st);
        case 5: return launch_stream<CCH, 5, 1, 16, F0REG>(p, st);
        case 6: return launch_stream<CCH, 6, 1, 20, F0REG>(p, st);
        case 7: return launch_stream<CCH, 7, 1, 20, F0REG>(p, st);
        case 8: return launch_stream<CCH, 8, 1, 24, F0REG>(p, st);
        default: return GFB_EUNSUPPORTED;
    }
}

}  // namespace gfb

using namespace gfb;

extern "C" int gfb_avg_pool2_f32(const float* x, float* y, int N, int H, int W, gfb_stream_t stream) {
    GFB_CHECK_ARG(x && y && N > 0 && H >= 2 && W >= 2);
    size_t total = (size_t)N * (H / 2) * (W / 2);
    int blocks = (int)min((size_t)148 * 16, (total + 255) / 256);
    avg_pool2_kernel<<<blocks, 256, 0, gfb_cu(stream)>>>(x, y, N, H, W);
    GFB_LAUNCH_RESULT();
}

extern "C" int gfb_pad_rows_f32(const float* x, float* y, long long rows, int W, int pitch, gfb_stream_t stream) {
    GFB_CHECK_ARG(x && y && rows > 0 && W > 0 && pitch >= W);
    size_t total = (size_t)rows * pitch;
    int blocks = (int)min((size_t)148 * 16, (total + 255) / 256);
    pad_rows_kernel<<<blocks, 256, 0, gfb_cu(stream)>>>(x, y, (size_t)rows, W, pitch);
    GFB_LAUNCH_RESULT();
}

extern "C" int gfb_debug_local_corr_counters(unsigned long long* host_out4, int reset) {
    cudaError_t e = cudaSuccess;
    if (host_out4) e = cudaMemcpyFromSymbol(host_out4, g_lc_stats, 8 * sizeof(unsigned long long));
    if (e == cudaSuccess && reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        e = cudaMemcpyToSymbol(g_lc_stats, z, sizeof(z));
    }
    return e == cudaSuccess ? GFB_OK : (int)e;
}

extern "C" int gfb_local_corr_f32(const float* f0, const float* f1, const float* flow, float* out,
                                  int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                                  int win_h, int win_w, int sample_mode, int padding_mode,
                                  int k_total, int k_offset, int algo, gfb_stream_t stream) {
    GFB_CHECK_ARG(f0 && f1 && flow && out);
    GFB_CHECK_ARG(B > 0 && C > 0 && Hs > 0 && Ws > 0 && G > 0 && r >= 0 && win_h > 0 && win_w > 0);
    GFB_CHECK_ARG(f1_pitch == 0 || f1_pitch >= Ws);
    GFB_CHECK_ARG(sample_mode == 0 || sample_mode == 1);
    GFB_CHECK_ARG(padding_mode == 0 || padding_mode == 1);
    const int kk = (2 * r + 1) * (2 * r + 1);
    GFB_CHECK_ARG(k_offset >= 0 && k_offset + kk <= k_total);
    GFB_CHECK_ARG(algo >= 0);
    LcParams p;
    p.f0 = f0; p.f1 = f1; p.flow = flow; p.out = out;
    p.B = B; p.C = C; p.Hs = Hs; p.Ws = Ws; p.G = G; p.r = r;
    p.pitch = f1_pitch ? f1_pitch : Ws;
    p.k_total = k_total; p.k_offset = k_offset;
    p.sample_mode = sample_mode; p.padding_mode = padding_mode;
    // python: torch.linspace(-2*r/n, 2*r/n, 2r+1): endpoints are doubles rounded to fp32
    p.ox0 = (float)(-2.0 * r / win_w); p.ox1 = (float)(2.0 * r / win_w);
    p.oy0 = (float)(-2.0 * r / win_h); p.oy1 = (float)(2.0 * r / win_h);
    p.inv_sqrt_c = (float)(1.0 / sqrt((double)C));
    p.debug = (algo >> 20) & 3;
    cudaStream_t st = gfb_cu(stream);

    const bool stream_ok = win_h == Hs && win_w == Ws && sample_mode == 0 && padding_mode == 0 &&
                           r >= 1 && r <= 8 && (p.pitch % 4 == 0) && gfb_aligned(f1, 16) &&
                           (C % 16 == 0) && (size_t)B * C < (1ull << 31);
    // algo = base | variant << 4 | tile-column divisor override << 8 | quad-row override << 12 | ring-stage override << 16 | debug << 20 | rows-per-stage override << 24
    //   base 0 auto, 1 generic gather kernel, 2 streaming kernels (error if the shape is not eligible)
    //   variant 0 auto, 1 P = 1 stream kernel, 2 / 4 legacy P-point variants, 8 quad kernel LDS.128, 9 quad kernel LDS.64
    const int base = algo & 15, variant = (algo >> 4) & 15;
    GFB_CHECK_ARG(base <= 2);
    if (base == 2 && !stream_ok) return GFB_EUNSUPPORTED;
    if (base != 1 && stream_ok) {
        int rc = GFB_EUNSUPPORTED;
        // auto: the quad kernel where it measured faster than the P = 1 kernel (C = 64, or lattice rows that split
        // into whole warps); see profiles/r1_kbench_sweep.json
        const bool quad_auto = variant == 0 && (C == 64 || G % 32 == 0);
        if (quad_auto || variant == 8 || variant == 9) {
            rc = variant == 8 ? launch_quad_any<4>(p, st, (algo >> 8) & 15, (algo >> 12) & 15, (algo >> 16) & 15, (algo >> 24) & 15)
                              : launch_quad_any<2>(p, st, (algo >> 8) & 15, (algo >> 12) & 15, (algo >> 16) & 15, (algo >> 24) & 15);
            if (rc != GFB_EUNSUPPORTED || variant != 0) return rc;
        }
        if (variant <= 1) {
            if (C == 16) rc = launch_stream_p1<16, true>(p, st);
            else if (C == 32) rc = launch_stream_p1<32, true>(p, st);
            else if (C % 64 == 0) rc = launch_stream_p1<64, false>(p, st);
        } else if (variant == 2) {
            if (C == 16 && r == 2) rc = launch_stream<16, 2, 2, 12, true>(p, st);
            else if (C == 32 && r == 4) rc = launch_stream<32, 4, 2, 16, true>(p, st);
            else if (C % 64 == 0 && r == 6) rc = launch_stream<64, 6, 2, 20, false>(p, st);
            else if (C % 64 == 0 && r == 7) rc = launch_stream<64, 7, 2, 20, false>(p, st);
        } else if (variant == 4) {
            if (C == 16 && r == 2) rc = launch_stream<16, 2, 4, 12, true>(p, st);
            else if (C == 32 && r == 4) rc = launch_stream<32, 4, 4, 16, false>(p, st);
            else if (C % 64 == 0 && r == 6) rc = launch_stream<64, 6, 4, 20, false>(p, st);
            else if (C % 64 == 0 && r == 7) rc = launch_stream<64, 7, 4, 24, false>(p, st);
        }
        if (rc != GFB_EUNSUPPORTED || base == 2 || variant != 0) return rc;
    }
    if (base == 2) return GFB_EUNSUPPORTED;
    const size_t total = (size_t)B * kk * G * G;
    int blocks = (int)min((size_t)148 * 32, (total + 255) / 256);
    lc_generic_kernel<<<blocks, 256, 0, st>>>(p);
    GFB_LAUNCH_RESULT();
}
