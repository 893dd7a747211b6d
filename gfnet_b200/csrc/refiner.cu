// Refiner input assembly (reference: ConvRefiner.forward, model/network.py:533-558).
//
//   d = cat(grid_feature, x_hat, emb_in_displacement, local_corr)           [B, 2C + dd + K, G, G]   (:555)
//   grid_feature = grid_sample(x, lattice)          bilinear, zeros, align_corners=False              (:547)
//   x_hat        = grid_sample(y, flow)                                                               (:537)
//   emb          = disp_emb(40/32 * scale_factor * (flow - lattice))       1x1 conv, 2 -> dd          (:548-549)
//   local_corr   = local_correlation(grid_feature, y, flow)                                           (:553-554)
//
// refiner_assemble_kernel writes the first 2C + dd channels straight into d; the local-correlation kernels then read
// their feature0 from d's leading C channels (LcParams::f0_ctot) and write the last K channels (k_total / k_offset): the
// concatenation never exists as a separate copy and feature0 never crosses the host boundary.
#include "common.cuh"
#include "lc_common.cuh"

namespace gfb {

// F.grid_sample(bilinear, zeros, align_corners=False) taps of one sample position
struct Taps {
    int o[4];
    float w[4];
};
__device__ __forceinline__ Taps make_taps(float gx, float gy, int Hs, int Ws, int pitch) {
    Taps t;
    const float ix = unnormalize(gx, Ws), iy = unnormalize(gy, Hs);
#pragma unroll
    for (int k = 0; k < 4; ++k) { t.o[k] = 0; t.w[k] = 0.f; }
    if (!(fabsf(ix) < 1e8f) || !(fabsf(iy) < 1e8f)) return t;          // far outside / non-finite: all taps out
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float tx = ix - x0f, ty = iy - y0f;
    const float wx[2] = {1.f - tx, tx}, wy[2] = {1.f - ty, ty};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int xx = x0 + (k & 1), yy = y0 + (k >> 1);
        if (xx >= 0 && xx < Ws && yy >= 0 && yy < Hs) { t.o[k] = yy * pitch + xx; t.w[k] = wx[k & 1] * wy[k >> 1]; }
    }
    return t;
}

// One lattice point per thread, 256 consecutive points (row-major) per block: every store instruction of a block writes
// 1 KB contiguous per channel plane (DRAM-friendly bursts); all channels in a loop unrolled four times: 32 independent
// taps in flight per thread.  HBM-bound: reads x and (through the
// flow) y once, writes (2C + dd) G^2 floats per batch element.
__global__ void __launch_bounds__(256) refiner_assemble_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                               const float* __restrict__ flow, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ d,
                                                               int C, int Hs, int Ws, int y_pitch, int G, int dd, int Dtot,
                                                               float emb_scale, int keep_grid, int cpb) {
    const int pt = blockIdx.x * 256 + threadIdx.x, b = blockIdx.y;
    const int c_lo = blockIdx.z * cpb, c_hi = min(C, c_lo + cpb);        // channel group of this block (small lattices: more blocks)
    const int o_lo = min(dd, c_lo), o_hi = blockIdx.z + 1 == gridDim.z ? dd : min(dd, c_hi);
    if (pt >= G * G) return;
    const int gy = pt / G, gx = pt - gy * G;
    const size_t gg = (size_t)G * G;
    const float fx = __ldg(flow + ((size_t)b * 2) * gg + (size_t)gy * G + gx);
    const float fy = __ldg(flow + ((size_t)b * 2 + 1) * gg + (size_t)gy * G + gx);
    const float lxn = linspace_at(-1.f + 1.f / (float)G, 1.f - 1.f / (float)G, G, gx);
    const float lyn = linspace_at(-1.f + 1.f / (float)G, 1.f - 1.f / (float)G, G, gy);
    const Taps ta = make_taps(lxn, lyn, Hs, Ws, Ws), tb = make_taps(fx, fy, Hs, Ws, y_pitch);
    float* dp = d + (size_t)b * Dtot * gg + (size_t)gy * G + gx;
    const unsigned xplane = (unsigned)(Hs * Ws), yplane = (unsigned)(Hs * y_pitch);
    const float* xb = x + (size_t)b * C * xplane;
    const float* yb = y + (size_t)b * C * yplane;
    // two separate loops: interleaving the lattice taps of x with the flow taps of y halves the achieved bandwidth
    if (!keep_grid) {           // (later refiner iterations of a scale: x has not changed, d[:, 0:C] still holds the grid features)
#pragma unroll 8
        for (int c = c_lo; c < c_hi; ++c) {
            const float* xp = xb + c * xplane;
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) a = fmaf(__ldg(xp + ta.o[k]), ta.w[k], a);
            dp[(size_t)c * gg] = a;          // read again by the correlation kernel: no streaming hint
        }
    }
#pragma unroll 8
    for (int c = c_lo; c < c_hi; ++c) {
        const float* yp = yb + c * yplane;
        float h = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) h = fmaf(__ldg(yp + tb.o[k]), tb.w[k], h);
        __stcs(dp + (size_t)(C + c) * gg, h);
    }
    const float vx = emb_scale * (fx - lxn), vy = emb_scale * (fy - lyn);
#pragma unroll 4
    for (int o = o_lo; o < o_hi; ++o)
        __stcs(dp + (size_t)(2 * C + o) * gg, fmaf(__ldg(w + 2 * o + 1), vy, fmaf(__ldg(w + 2 * o), vx, __ldg(bias + o))));
}

}  // namespace gfb

using namespace gfb;

// x [B,C,Hs,Ws] (image-A feature map), y [B,C,Hs,y_pitch >= Ws] (image-B feature map), flow [B,2,G,G], w [dd,2], bias [dd]
// -> d[:, 0 : 2C + dd] of d [B,Dtot,G,G].  emb_scale = 40/32 * scale_factor.  keep_grid != 0: d[:, 0:C] already holds the grid
// features of the same x (later refiner iterations of a scale): only x_hat and the embedding are rewritten.
extern "C" int gfb_refiner_assemble_f32(const float* x, const float* y, const float* flow, const float* w, const float* bias,
                                        float* d, int B, int C, int Hs, int Ws, int y_pitch, int G, int dd, int Dtot,
                                        float emb_scale, int keep_grid, gfb_stream_t stream) {
    GFB_CHECK_ARG(x && y && flow && d && (dd == 0 || (w && bias)));
    GFB_CHECK_ARG(B > 0 && C > 0 && Hs > 0 && Ws > 0 && G > 0 && dd >= 0 && Dtot >= 2 * C + dd);
    GFB_CHECK_ARG(y_pitch == 0 || y_pitch >= Ws);
    GFB_CHECK_ARG((size_t)C * Hs * (y_pitch ? y_pitch : Ws) < (1ull << 31) && B <= 65535 && G <= 23170);
    // small lattices: split the channels over blockIdx.z so that every SM gets several blocks
    int cpb = C;
    while (cpb > 8 && cpb % 2 == 0 && (long long)((G * G + 255) / 256) * B * (C / cpb) < 148 * 8) cpb /= 2;
    dim3 grid((G * G + 255) / 256, B, (C + cpb - 1) / cpb), block(256);
    refiner_assemble_kernel<<<grid, block, 0, gfb_cu(stream)>>>(x, y, flow, w, bias, d, C, Hs, Ws, y_pitch ? y_pitch : Ws, G, dd,
                                                                  Dtot, emb_scale, keep_grid, cpb);
    GFB_LAUNCH_RESULT();
}
