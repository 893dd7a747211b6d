// K1 -- local correlation, general entry (reference: utils/local_correlation.py:4-72, call site model/network.py:553).
//
// corr[b,k,gy,gx] = (1/sqrt C) sum_c f0[b,c,gy,gx] * sample(f1[b,c], flow[b,:,gy,gx] + off_k)
//
// lc_generic_kernel: one thread per output element, direct gathers.  Mirrors the reference's coordinate arithmetic
// exactly (per-k fp32 `flow + offset`, unnormalise, floor) and covers every mode the reference can be called with that we
// support (bilinear / nearest, zeros / border, grid-based windows, any C and r).  It is the path for configurations
// outside the hot kernels of local_corr_v2.cu (lc_rot_kernel, lc_pt_kernel, lc_tc2_kernel) and the cross-check the
// parity tests compare them with at full size.  The round-1 streaming kernels (lc_stream / lc_quad and the converter-warp
// tcgen05 kernel) were superseded by local_corr_v2.cu and removed.
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

// local_correlation.py:71: F.avg_pool2d(feature1, 2, 2) between pyramid levels
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

// rows of W floats -> rows of `pitch` floats (zero tail): the TMA descriptors need 16-byte global strides
__global__ void pad_rows_kernel(const float* __restrict__ x, float* __restrict__ y, size_t rows, int W, int pitch) {
    const size_t total = rows * pitch;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int xx = (int)(t % pitch);
        const size_t r = t / pitch;
        y[t] = xx < W ? x[r * W + xx] : 0.f;
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

extern "C" int gfb_local_corr_f32(const float* f0, const float* f1, const float* flow, float* out,
                                  int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                                  int win_h, int win_w, int sample_mode, int padding_mode,
                                  int k_total, int k_offset, gfb_stream_t stream) {
    GFB_CHECK_ARG(f0 && f1 && flow && out);
    GFB_CHECK_ARG(B > 0 && C > 0 && Hs > 0 && Ws > 0 && G > 0 && r >= 0 && win_h > 0 && win_w > 0);
    GFB_CHECK_ARG(f1_pitch == 0 || f1_pitch >= Ws);
    GFB_CHECK_ARG(sample_mode == 0 || sample_mode == 1);
    GFB_CHECK_ARG(padding_mode == 0 || padding_mode == 1);
    const int kk = (2 * r + 1) * (2 * r + 1);
    GFB_CHECK_ARG(k_offset >= 0 && k_offset + kk <= k_total);
    LcParams p;
    p.f0 = f0; p.f1 = f1; p.flow = flow; p.out = out;
    p.B = B; p.C = C; p.Hs = Hs; p.Ws = Ws; p.G = G; p.r = r;
    p.Ctot = C; p.c0 = 0; p.accumulate = 0; p.f0_ctot = C;
    p.pitch = f1_pitch ? f1_pitch : Ws;
    p.k_total = k_total; p.k_offset = k_offset;
    p.sample_mode = sample_mode; p.padding_mode = padding_mode;
    // python: torch.linspace(-2*r/n, 2*r/n, 2r+1): endpoints are doubles rounded to fp32
    p.ox0 = (float)(-2.0 * r / win_w); p.ox1 = (float)(2.0 * r / win_w);
    p.oy0 = (float)(-2.0 * r / win_h); p.oy1 = (float)(2.0 * r / win_h);
    p.inv_sqrt_c = (float)(1.0 / sqrt((double)C));
    p.debug = 0;
    const size_t total = (size_t)B * kk * G * G;
    int blocks = (int)min((size_t)148 * 32, (total + 255) / 256);
    lc_generic_kernel<<<blocks, 256, 0, gfb_cu(stream)>>>(p);
    GFB_LAUNCH_RESULT();
}
