// Shared by the local-correlation kernels: parameters and the exact per-sample gather (reference arithmetic).
#pragma once
#include "common.cuh"
#include <math.h>

namespace gfb {

struct LcParams {
    const float* f0;
    const float* f1;
    const float* flow;
    float* out;
    int B, C, Hs, Ws, G, r;
    int Ctot, c0;              // channel slice: the C channels correlated are c0 .. c0 + C of tensors with Ctot channels
    int f0_ctot;               // channels per batch element of the tensor feature0 lives in (= Ctot unless feature0 is the
                               // leading channel block of the refiner-input buffer [B, 2C+d+K, G, G], model/network.py:555)
    int accumulate;            // out += instead of out = (second and later channel slices of one correlation)
    int pitch;                 // floats between rows of f1 (>= Ws)
    int k_total, k_offset;
    int sample_mode, padding_mode;
    float ox0, ox1, oy0, oy1;  // torch.linspace endpoints of the window offsets (fp32)
    float inv_sqrt_c;
    int debug;                 // profiling aids of the debug entry points only (0 in every production call)
};

// torch.linspace(start, end, steps)[i] in fp32 (ATen RangeFactories: symmetric evaluation).
__device__ __forceinline__ float linspace_at(float start, float end, int steps, int i) {
    if (steps <= 1) return start;
    float step = (end - start) / (float)(steps - 1);
    return (i < steps / 2) ? start + step * (float)i : end - step * (float)(steps - 1 - i);
}

__device__ __forceinline__ float unnormalize(float c, int size) {  // align_corners = False
    return ((c + 1.f) * (float)size - 1.f) / 2.f;
}

// flow of one lattice point -> window origin, bilinear weights, liveness (window intersects the image)
struct PointGeom {
    int xb, yb;
    float fx, fy;
    bool live;
};
__device__ __forceinline__ PointGeom point_geom(const LcParams& p, int b, int gy, int gx, bool valid, int R) {
    PointGeom g;
    g.xb = 0; g.yb = 0; g.fx = 0.f; g.fy = 0.f; g.live = false;
    if (!valid) return g;
    const size_t gg = (size_t)p.G * p.G;
    const float* fl = p.flow + (size_t)b * 2 * gg + (size_t)gy * p.G + gx;
    const float sx = unnormalize(__ldg(fl), p.Ws), sy = unnormalize(__ldg(fl + gg), p.Hs);
    if (fabsf(sx) < 1e6f && fabsf(sy) < 1e6f) {
        const float x0f = floorf(sx), y0f = floorf(sy);
        const int W = 2 * R + 2;
        g.xb = (int)x0f - R; g.yb = (int)y0f - R;
        g.fx = sx - x0f; g.fy = sy - y0f;
        g.live = !(g.xb >= p.Ws || g.xb + W <= 0 || g.yb >= p.Hs || g.yb + W <= 0);
    }
    return g;
}

// One output element with the reference's exact per-sample coordinate arithmetic.
static __device__ float lc_generic_point(const LcParams& p, int b, int k, int gy, int gx) {
    const int kw = 2 * p.r + 1;
    const int iy = k / kw, ix = k - iy * kw;
    const size_t gg = (size_t)p.G * p.G;
    const float* fl = p.flow + (size_t)b * 2 * gg + (size_t)gy * p.G + gx;
    float px = fl[0] + linspace_at(p.ox0, p.ox1, kw, ix);
    float py = fl[gg] + linspace_at(p.oy0, p.oy1, kw, iy);
    float sx = unnormalize(px, p.Ws), sy = unnormalize(py, p.Hs);
    if (p.padding_mode == 1) {
        sx = fminf((float)(p.Ws - 1), fmaxf(sx, 0.f));
        sy = fminf((float)(p.Hs - 1), fmaxf(sy, 0.f));
    }
    const float* f0 = p.f0 + ((size_t)b * p.f0_ctot + p.c0) * gg + (size_t)gy * p.G + gx;
    const size_t plane = (size_t)p.Hs * p.pitch;
    const float* f1 = p.f1 + ((size_t)b * p.Ctot + p.c0) * plane;
    float acc = 0.f;
    if (p.sample_mode == 1) {
        float rx = rintf(sx), ry = rintf(sy);
        if (rx >= 0.f && rx < (float)p.Ws && ry >= 0.f && ry < (float)p.Hs) {
            size_t o = (size_t)(int)ry * p.pitch + (int)rx;
            for (int c = 0; c < p.C; ++c) acc = fmaf(f0[c * gg], f1[c * plane + o], acc);
        }
        return acc * p.inv_sqrt_c;
    }
    if (!(fabsf(sx) < 1e8f) || !(fabsf(sy) < 1e8f)) return 0.f;  // far outside / non-finite: all taps out
    float x0f = floorf(sx), y0f = floorf(sy);
    int x0 = (int)x0f, y0 = (int)y0f;
    float tx = sx - x0f, ty = sy - y0f;
    float w00 = (1.f - tx) * (1.f - ty), w01 = tx * (1.f - ty), w10 = (1.f - tx) * ty, w11 = tx * ty;
    bool xa = x0 >= 0 && x0 < p.Ws, xb = x0 + 1 >= 0 && x0 + 1 < p.Ws;
    bool ya = y0 >= 0 && y0 < p.Hs, yb = y0 + 1 >= 0 && y0 + 1 < p.Hs;
    if (!xa) { w00 = 0.f; w10 = 0.f; }
    if (!xb) { w01 = 0.f; w11 = 0.f; }
    if (!ya) { w00 = 0.f; w01 = 0.f; }
    if (!yb) { w10 = 0.f; w11 = 0.f; }
    int xc0 = min(max(x0, 0), p.Ws - 1), xc1 = min(max(x0 + 1, 0), p.Ws - 1);
    int yc0 = min(max(y0, 0), p.Hs - 1), yc1 = min(max(y0 + 1, 0), p.Hs - 1);
    size_t o00 = (size_t)yc0 * p.pitch + xc0, o01 = (size_t)yc0 * p.pitch + xc1;
    size_t o10 = (size_t)yc1 * p.pitch + xc0, o11 = (size_t)yc1 * p.pitch + xc1;
    for (int c = 0; c < p.C; ++c) {
        const float* q = f1 + c * plane;
        float s = w00 * __ldg(q + o00) + w01 * __ldg(q + o01) + w10 * __ldg(q + o10) + w11 * __ldg(q + o11);
        acc = fmaf(__ldg(f0 + c * gg), s, acc);
    }
    return acc * p.inv_sqrt_c;
}

// local_corr_mma.cu: launch the mma.sync kernel on filled parameters (shape 0 = auto CTA shape); GFB_EUNSUPPORTED for (r, C)
// it is not instantiated for
int lc_mma_launch(const LcParams& p, int shape, cudaStream_t st);

}  // namespace gfb
