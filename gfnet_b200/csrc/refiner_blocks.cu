// Refiner convolution blocks (reference: ConvRefiner.create_block / forward tail, model/network.py:505-531,
// 557-563) and the flow update + between-scale upsampling of the decoder loop (model/network.py:262-285).
//
//   block(h) = conv2_1x1( relu( bn( dwconv5x5(h) ) ) )            9 blocks per refiner (block1 + 8 hidden), in_dim = hidden_dim
//   delta    = out_conv_1x1( blocks(d).float() )                  hidden -> 3 (dx, dy, d_certainty), fp32
//
// The reference runs the blocks under fp16 autocast (amp=True at every call site, model/network.py:90).  Here activations are
// stored as fp16 NHWC ([B, G*G, Cp], Cp = C rounded up to 16, pad channels zero) and every sum is accumulated in fp32:
//   rb_pack_kernel   d [B,C,G,G] fp32 NCHW -> fp16 NHWC (the refiner input written by the assemble + correlation kernels)
//   rb_dw_kernel     depth-wise 5x5 + batch norm (eval, folded into the taps + a shift) + ReLU; lane = channel pair, a warp
//                    walks a 4-pixel-wide strip downwards with the 5 pending output rows in registers (packed f32x2 FMAs)
//   rb_pw_kernel     the 1x1 convolution as a GEMM on tcgen05: A = 128 pixels x K (TMA, 128B swizzle), B = W2 rows (TMA),
//                    fp32 accumulator in TMEM, epilogue adds the bias and writes fp16 NHWC; persistent over pixel tiles
//   rb_out_kernel    out_conv (hidden -> 3) in fp32, NCHW output
// gfb_refiner_blocks_f16 runs the chain over two ping-pong activation buffers (the whole batch at once; `chunk` bounds the
// buffers: L2-sized chunks were measured slower, see gfb_refiner_blocks_chunk).
#include "common.cuh"
#include <cuda_fp16.h>

namespace gfb {
namespace rb {

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
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long h2_to_f2(uint32_t h) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
    return pack2(f.x, f.y);
}
__device__ __forceinline__ uint32_t f2_to_h2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// ---------------------------------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC fp16, 64 channels x 32 pixels per block through a shared-memory transpose
__global__ void __launch_bounds__(256) rb_pack_kernel(const float* __restrict__ d, __half* __restrict__ out, int C, int Cp, int P) {
    __shared__ float t[64][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 64; i += 8) {
        const int c = c0 + i, p = p0 + tx;
        t[i][tx] = (c < C && p < P) ? __ldcs(d + ((size_t)b * C + c) * P + p) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int p = p0 + i, c = c0 + 2 * tx;
        if (p < P && c < Cp)
            *reinterpret_cast<uint32_t*>(out + ((size_t)b * P + p) * Cp + c) = f2_to_h2(t[2 * tx][i], t[2 * tx + 1][i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// depth-wise 5x5 (zero padding 2) + folded batch norm + ReLU on NHWC fp16.  wf [25][Cp] fp32 = tap * bn_scale, shift [Cp] =
// (conv_bias - running_mean) * bn_scale + bn_bias.  A warp = 64 channels (lane = channel pair) x a strip of 4 x th outputs;
// input row j of the strip feeds the five output rows j-4..j, whose accumulators live in a register ring (static slots: the
// row loop is unrolled by five).
constexpr int DW_TH_MIN = 16;  // output rows per strip: run-time `th` with (th + 4) % 5 == 0, chosen by launch_dw
constexpr int DW_WARPS = 4;    // neighbouring strips of one channel chunk: their x halos meet in L1

// lp = channel pairs per chunk (8, 16 or 32 lanes); the other 32 / lp lane groups of a warp take neighbouring strips, so that
// narrow refiners (Cp = 32, 80) keep every lane busy.
// CP = the channel pitch as a compile-time constant for the refiners GFNet builds (every tap / pixel offset an immediate: the
// run-time pitch spent 120 of 320 instructions per input row on 64-bit addresses), 0 = run-time pitch.
template <int CP>
__global__ void __launch_bounds__(DW_WARPS * 32, 3) rb_dw_kernel(const __half* __restrict__ in, const float* __restrict__ wf,
                                                                 const float* __restrict__ shift, __half* __restrict__ out,
                                                                 int G, int Cp_rt, int nchunk, int lp, int th) {
    const int Cp = CP ? CP : Cp_rt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = blockIdx.x % nchunk, xs = blockIdx.x / nchunk;
    const int spw = 32 / lp;                                          // strips per warp
    const int c = (chunk * lp + lane % lp) * 2;
    const int x0 = ((xs * DW_WARPS + warp) * spw + lane / lp) * 4, y0 = blockIdx.y * th, b = blockIdx.z;
    const bool cok = c < Cp && x0 < G;
    unsigned long long w[25];
#pragma unroll
    for (int t = 0; t < 25; ++t) w[t] = cok ? pack2(__ldg(wf + t * Cp + c), __ldg(wf + t * Cp + c + 1)) : 0ull;
    const float sh0 = cok ? __ldg(shift + c) : 0.f, sh1 = cok ? __ldg(shift + c + 1) : 0.f;
    unsigned long long acc[5][4];
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[s][p] = 0ull;
    const __half* inb = in + (size_t)b * G * G * Cp + c;
    __half* outb = out + (size_t)b * G * G * Cp + c;
    bool xok[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xok[i] = cok && x0 - 2 + i >= 0 && x0 - 2 + i < G;

    // the loads of input row j + 1 are issued before the arithmetic of row j (one row of look-ahead per warp)
    uint32_t raw[8];
    const __half* row = inb + ((long long)(y0 - 2) * G + (x0 - 2)) * Cp;      // walks down one image row per input row
    const long long row_pitch = (long long)G * Cp;
    auto load_row = [&](int j) {
        const int iy = y0 - 2 + j;
        const bool yok = iy >= 0 && iy < G && j < th + 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            raw[i] = 0u;
            if (yok && xok[i]) raw[i] = __ldg(reinterpret_cast<const uint32_t*>(row + i * Cp));
        }
        row += row_pitch;
    };
    load_row(0);
    __half* orow = outb + ((long long)(y0 - 4) * G + x0) * Cp;               // output row j - 4
#pragma unroll 1
    for (int base = 0; base < th + 4; base += 5) {
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int j = base + s;
            unsigned long long v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = h2_to_f2(raw[i]);
            load_row(j + 1);
#pragma unroll
            for (int ky = 0; ky < 5; ++ky) {
                const int slot = (s - ky + 5) % 5;       // output row j - ky
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int kx = 0; kx < 5; ++kx) acc[slot][p] = fma2(w[ky * 5 + kx], v[p + kx], acc[slot][p]);
            }
            // output row o = j - 4 is complete (its last tap row ky = 4 was this input row): slot (s + 1) % 5
            const int o = j - 4, oy = y0 + o, slot = (s + 1) % 5;
            if (o >= 0 && oy < G && cok) {
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float2 a = unpack2(acc[slot][p]);
                    if (x0 + p < G)
                        *reinterpret_cast<uint32_t*>(orow + p * Cp) = f2_to_h2(fmaxf(a.x + sh0, 0.f), fmaxf(a.y + sh1, 0.f));
                }
            }
            orow += row_pitch;
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[slot][p] = 0ull;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// 1x1 convolution = GEMM  out[p, n] = fp16( sum_k act[p, k] * w2[n, k] + bias[n] ),  p < P, n, k < Cp.
// any-shape CUDA-core version (cross-check of the tcgen05 kernel; algo = 1)
__global__ void __launch_bounds__(128) rb_pw_simt_kernel(const __half* __restrict__ act, const __half* __restrict__ w2,
                                                         const float* __restrict__ bias, __half* __restrict__ out, size_t P, int Cp) {
    const size_t p = blockIdx.x;
    for (int n = threadIdx.x; n < Cp; n += 128) {
        float a = 0.f;
        for (int k = 0; k < Cp; ++k) a = fmaf(__half2float(act[p * Cp + k]), __half2float(w2[(size_t)n * Cp + k]), a);
        out[p * Cp + n] = __float2half_rn(a + bias[n]);
    }
}

// tcgen05 version -------------------------------------------------------------------------------------------------------
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
// instruction descriptor: D = F32, A = B = F16, both K-major, M = 128, N runtime
__device__ __forceinline__ uint32_t idesc_f16(int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// mma.sync version for narrow refiners (Cp <= 96: scales 2 and 1) -----------------------------------------------------------
// With K = N <= 96 the 1x1 convolution is a streaming operation (arithmetic intensity <= 48 flop/B): the tcgen05 kernel spends
// its time on per-tile hand-offs (TMA -> MMA -> TMEM -> registers) of 8-20 KB tiles.  Here a warp owns 16 pixels at a time and
// nothing is staged: a lane's coalesced 16-byte loads of the NHWC rows ARE its A fragments, and its D fragments ARE 16-byte
// row segments, because the k and n orders of the MMA are permuted consistently in the (shared-memory) weight fragments:
//   lane (g = lane / 4, t = lane % 4), load l, half u:   k-slot 2t + e   <->  channel 32 l + 8 t + 4 u + e
//                                                        k-slot 2t + 8 + e <->  channel 32 l + 8 t + 4 u + 2 + e
//   n-tile (q, j), column n                              <->  output channel 32 q + 8 (n / 2) + 2 j + n % 2
// so that over j = 0..3 lane t holds output channels 32 q + 8 t .. + 7 of rows g and g + 8.
__device__ __forceinline__ void mma_m16n8k16_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                                 uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int PWM_THREADS = 128;

template <int KL>      // 32-channel groups: Cp <= 32 KL
__global__ void __launch_bounds__(PWM_THREADS, 4) rb_pw_mma_kernel(const __half* __restrict__ act, const __half* __restrict__ w2,
                                                                   const float* __restrict__ bias, __half* __restrict__ out,
                                                                   long long P, int Cp) {
    __shared__ uint2 wfrag[KL * 2 * KL * 4][32];            // [(l, u, q, j)][lane] = (b0, b1)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < KL * 2 * KL * 4 * 32; i += PWM_THREADS) {
        const int L = i & 31, f = i >> 5, j = f & 3, q = (f >> 2) % KL, u = (f / (4 * KL)) & 1, l = f / (8 * KL);
        const int co = 32 * q + 8 * ((L >> 2) >> 1) + 2 * j + ((L >> 2) & 1), ch = 32 * l + 8 * (L & 3) + 4 * u;
        uint2 v = make_uint2(0u, 0u);
        if (co < Cp && ch < Cp) v = __ldg(reinterpret_cast<const uint2*>(w2 + (size_t)co * Cp + ch));
        wfrag[f][L] = v;
    }
    float bs[KL][8];
#pragma unroll
    for (int q = 0; q < KL; ++q)
#pragma unroll
        for (int e = 0; e < 8; ++e) bs[q][e] = 32 * q + 8 * t + e < Cp ? __ldg(bias + 32 * q + 8 * t + e) : 0.f;
    __syncthreads();

    const long long nblk = (P + 15) / 16;
    for (long long blk = (long long)blockIdx.x * (PWM_THREADS / 32) + warp; blk < nblk; blk += (long long)gridDim.x * (PWM_THREADS / 32)) {
        const long long r0 = blk * 16 + g, r1 = r0 + 8;
        uint4 alo[KL], ahi[KL];
#pragma unroll
        for (int l = 0; l < KL; ++l) {
            alo[l] = make_uint4(0u, 0u, 0u, 0u);
            ahi[l] = make_uint4(0u, 0u, 0u, 0u);
            if (32 * l + 8 * t < Cp) {
                if (r0 < P) alo[l] = __ldcs(reinterpret_cast<const uint4*>(act + (size_t)r0 * Cp + 32 * l + 8 * t));
                if (r1 < P) ahi[l] = __ldcs(reinterpret_cast<const uint4*>(act + (size_t)r1 * Cp + 32 * l + 8 * t));
            }
        }
#pragma unroll
        for (int q = 0; q < KL; ++q) {
            float acc[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll
            for (int l = 0; l < KL; ++l) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint2 b_u0 = wfrag[((l * 2 + 0) * KL + q) * 4 + j][lane];
                    const uint2 b_u1 = wfrag[((l * 2 + 1) * KL + q) * 4 + j][lane];
                    mma_m16n8k16_f16(acc[j], alo[l].x, ahi[l].x, alo[l].y, ahi[l].y, b_u0.x, b_u0.y);
                    mma_m16n8k16_f16(acc[j], alo[l].z, ahi[l].z, alo[l].w, ahi[l].w, b_u1.x, b_u1.y);
                }
            }
            if (32 * q + 8 * t < Cp) {
                uint4 o0, o1;
                o0.x = f2_to_h2(acc[0][0] + bs[q][0], acc[0][1] + bs[q][1]); o1.x = f2_to_h2(acc[0][2] + bs[q][0], acc[0][3] + bs[q][1]);
                o0.y = f2_to_h2(acc[1][0] + bs[q][2], acc[1][1] + bs[q][3]); o1.y = f2_to_h2(acc[1][2] + bs[q][2], acc[1][3] + bs[q][3]);
                o0.z = f2_to_h2(acc[2][0] + bs[q][4], acc[2][1] + bs[q][5]); o1.z = f2_to_h2(acc[2][2] + bs[q][4], acc[2][3] + bs[q][5]);
                o0.w = f2_to_h2(acc[3][0] + bs[q][6], acc[3][1] + bs[q][7]); o1.w = f2_to_h2(acc[3][2] + bs[q][6], acc[3][3] + bs[q][7]);
                if (r0 < P) *reinterpret_cast<uint4*>(out + (size_t)r0 * Cp + 32 * q + 8 * t) = o0;
                if (r1 < P) *reinterpret_cast<uint4*>(out + (size_t)r1 * Cp + 32 * q + 8 * t) = o1;
            }
        }
    }
}

template <int KL>
static int launch_pw_mma(const __half* act, const __half* w2, const float* bias, __half* out, long long P, int Cp, cudaStream_t st) {
    const long long nblk = (P + 15) / 16;
    const long long want = (nblk + PWM_THREADS / 32 - 1) / (PWM_THREADS / 32);
    const int grid = (int)min(want, (long long)148 * 4);              // one resident wave (4 blocks per SM): the weight fragments
                                                                      // are built once per block, warps stride over the pixels
    rb_pw_mma_kernel<KL><<<grid, PWM_THREADS, 0, st>>>(act, w2, bias, out, P, Cp);
    return (int)cudaGetLastError();
}

struct PwParams {
    const float* bias;
    __half* out;
    long long P;       // pixels (rows of act / out)
    int Cp;            // padded channels: K extent, N extent, row pitch of act / w2 / out in halves
    int NT;            // N tile (multiple of 16, <= 256); nsplit tiles cover Cp
    int nsplit;
    int KA;            // K atoms of 64 halves (128 B)
    int klast;         // K steps of 16 in the last atom
    int nstage;
    int nwork;         // pixel tiles * nsplit
};
constexpr int PW_THREADS = 6 * 32;   // warp 0: TMA, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int PW_MAXSTAGE = 6;
constexpr int PW_TMEM = 256;

__global__ void __launch_bounds__(PW_THREADS, 2) rb_pw_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB, const PwParams g) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[PW_MAXSTAGE], empty[PW_MAXSTAGE], d_full[4], d_empty[4];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t a_bytes = 128 * 128, b_bytes = (uint32_t)g.NT * 128, stage_bytes = a_bytes + b_bytes;
    // accumulators in flight (256 TMEM columns): narrow tiles are latency-bound per tile, so keep up to four going
    const uint32_t nacc = g.NT <= 64 ? 4 : g.NT <= 128 ? 2 : 1, acc_cols = PW_TMEM / nacc;

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 4; ++a) { mbar_init(&d_full[a], 1); mbar_init(&d_empty[a], 4); }
        mbar_fence_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, PW_TMEM);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int wk = blockIdx.x; wk < g.nwork; wk += gridDim.x) {
                const int half = wk % g.nsplit, mt = wk / g.nsplit;
                for (int ka = 0; ka < g.KA; ++ka, ++it) {
                    const uint32_t s = it % g.nstage;
                    mbar_wait(&empty[s], ((it / g.nstage) & 1) ^ 1);
                    unsigned char* st = smem + (size_t)s * stage_bytes;
                    mbar_expect_tx(&full[s], stage_bytes);
                    tma_load_2d(st, &tmA, &full[s], ka * 32, mt * 128);
                    tma_load_2d(st + a_bytes, &tmB, &full[s], ka * 32, half * g.NT);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = idesc_f16(g.NT);
            uint32_t it = 0, t = 0;
            for (int wk = blockIdx.x; wk < g.nwork; wk += gridDim.x, ++t) {
                const uint32_t a = t % nacc;
                mbar_wait(&d_empty[a], ((t / nacc) & 1) ^ 1);
                fence_after_sync();
                const uint32_t dt = tmem_base + a * acc_cols;
                uint32_t accum = 0;
                for (int ka = 0; ka < g.KA; ++ka, ++it) {
                    const uint32_t s = it % g.nstage;
                    mbar_wait(&full[s], (it / g.nstage) & 1);
                    fence_after_sync();
                    const uint32_t a_addr = smem_u32(smem + (size_t)s * stage_bytes), b_addr = a_addr + a_bytes;
                    const int nk = ka + 1 == g.KA ? g.klast : 4;
                    for (int ks = 0; ks < nk; ++ks) {
                        mma_f16(dt, smem_desc_k128(a_addr + ks * 32), smem_desc_k128(b_addr + ks * 32), idesc, accum);
                        accum = 1;
                    }
                    mma_commit(&empty[s]);
                }
                mma_commit(&d_full[a]);
            }
        }
    } else {
        const int q = warp & 3;                        // TMEM lane quarter this warp may read
        uint32_t t = 0;
        int half_s = -1;
        for (int wk = blockIdx.x; wk < g.nwork; wk += gridDim.x, ++t) {
            const int half = wk % g.nsplit, mt = wk / g.nsplit;
            if (half != half_s) {                      // the bias of this N tile (a CTA keeps its tile: the grid is even)
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int e = threadIdx.x - 64;
                for (int n = e; n < g.NT; n += 128) bias_s[n] = half * g.NT + n < g.Cp ? __ldg(g.bias + half * g.NT + n) : 0.f;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                half_s = half;
            }
            const uint32_t a = t % nacc;
            mbar_wait(&d_full[a], (t / nacc) & 1);
            fence_after_sync();
            const long long p = (long long)mt * 128 + q * 32 + lane;
            __half* orow = g.out + (size_t)p * g.Cp;
            const int nchunks = (g.NT + 31) / 32;
            for (int ch = 0; ch < nchunks; ++ch) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + a * acc_cols + (uint32_t)ch * 32u, r);
                tmem_ld_wait();
                const int n0 = half * g.NT + ch * 32;
#pragma unroll
                for (int v8 = 0; v8 < 4; ++v8) {
                    const int n = n0 + v8 * 8;
                    if (ch * 32 + v8 * 8 < g.NT && n < g.Cp && p < g.P) {
                        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + ch * 32 + v8 * 8);
                        const float4 b1 = *reinterpret_cast<const float4*>(bias_s + ch * 32 + v8 * 8 + 4);
                        uint4 o;
                        o.x = f2_to_h2(__uint_as_float(r[v8 * 8 + 0]) + b0.x, __uint_as_float(r[v8 * 8 + 1]) + b0.y);
                        o.y = f2_to_h2(__uint_as_float(r[v8 * 8 + 2]) + b0.z, __uint_as_float(r[v8 * 8 + 3]) + b0.w);
                        o.z = f2_to_h2(__uint_as_float(r[v8 * 8 + 4]) + b1.x, __uint_as_float(r[v8 * 8 + 5]) + b1.y);
                        o.w = f2_to_h2(__uint_as_float(r[v8 * 8 + 6]) + b1.z, __uint_as_float(r[v8 * 8 + 7]) + b1.w);
                        *reinterpret_cast<uint4*>(orow + n) = o;
                    }
                }
            }
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[a]);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, PW_TMEM);
}

// ---------------------------------------------------------------------------------------------------------------------
// Depth-wise 5x5 on the tensor cores, channel-planar fp16 (EXPERIMENT): per plane and ky the convolution along
// x is a banded (Toeplitz) 16 x 8 matrix T_ky[k][n] = w[ky][k - n], so 16 output rows x 8 output pixels are five
// mma.sync.m16n8k16 with A_ky[m][k] = in[y0 + m + ky - 2][x0 - 2 + k] read straight from the plane (a lane's 4-byte loads are
// its A fragments; neighbouring 8-pixel tiles share half of them) and T_ky a per-channel constant in registers.
// A warp = 16 rows x 32 pixels of one plane: 50 loads, 20 MMAs, 8 stores for 512 outputs (the FFMA2 kernel issues ~500
// instructions for as many).  Taps are rounded to fp16 (as the reference's autocast does), sums are fp32.
constexpr int DWM_PITCH = 36;      // words per staged row (48 pixels + pad): pitch = 4 (mod 32) -> conflict-free fragment reads
constexpr int DWM_ROWS = 20;       // 16 output rows + 4 halo rows per step

// A warp owns a 32-pixel-wide column strip of one plane and walks it downwards in steps of 16 rows: the B fragments are built
// once, every step stages 20 rows x 48 pixels (16-byte chunks, prefetched into registers one step ahead) in a warp-private
// shared-memory tile, reads its 50 A-fragment words, issues 20 MMAs and stores 16 x 32 outputs: ~135 instructions per 512
// outputs.  No block-wide barrier.
__global__ void __launch_bounds__(128) rb_dwm_kernel(const __half* __restrict__ in, const float* __restrict__ wf,
                                                     const float* __restrict__ shift, __half* __restrict__ out,
                                                     int C, int Cp, int G, int seg) {
    __shared__ __align__(16) uint32_t tiles[4][DWM_ROWS * DWM_PITCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int plane = blockIdx.z, c = plane % C;
    const int x0 = (blockIdx.x * 4 + warp) * 32;
    if (x0 >= G) return;
    const int ys = blockIdx.y * seg, ye = min(G, ys + seg);
    const __half* ip = in + (size_t)plane * G * G;
    __half* op = out + (size_t)plane * G * G;
    uint32_t* tile = tiles[warp];
    uint32_t b0[5], b1[5];
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
        float e[4];
        const int d[4] = {2 * t - g, 2 * t + 1 - g, 2 * t + 8 - g, 2 * t + 9 - g};
#pragma unroll
        for (int i = 0; i < 4; ++i) e[i] = (d[i] >= 0 && d[i] <= 4) ? __ldg(wf + (ky * 5 + d[i]) * Cp + c) : 0.f;
        b0[ky] = f2_to_h2(e[0], e[1]);
        b1[ky] = f2_to_h2(e[2], e[3]);
    }
    const float sh = __ldg(shift + c);
    // the lane's four 16-byte chunks of a step: chunk index lane + 32 q -> (row, chunk of the row); 120 chunks per step
    int ry[4], cx[4];
    bool cok[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int idx = lane + 32 * q;
        ry[q] = idx / 6;
        const int ch = idx - ry[q] * 6;
        cx[q] = x0 - 8 + 8 * ch;
        cok[q] = idx < DWM_ROWS * 6 && cx[q] >= 0 && cx[q] < G;
    }
    // per-lane pointers, advanced by 16 rows per step (no 64-bit address arithmetic inside the loop)
    const uint4* lp[4];
    uint32_t* sp[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        lp[q] = reinterpret_cast<const uint4*>(ip + ((long long)(ys - 2 + ry[q]) * G + cx[q]));
        sp[q] = tile + ry[q] * DWM_PITCH + ((cx[q] - x0 + 8) >> 1);
    }
    const long long step16 = (long long)2 * G;                 // 16 rows in uint4 units (8 halves each)
    uint4 pre[4];
    auto fetch = [&](int y0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int y = y0 - 2 + ry[q];
            pre[q] = make_uint4(0u, 0u, 0u, 0u);
            if (cok[q] && y >= 0 && y < G) pre[q] = __ldg(lp[q]);
            lp[q] += step16;
        }
    };
    fetch(ys);
    const uint32_t* fr = tile + g * DWM_PITCH + 3 + t;         // the lane's fragment words: fr[row * PITCH + 4 j]
    for (int y0 = ys; y0 < ye; y0 += 16) {
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (lane + 32 * q < DWM_ROWS * 6) *reinterpret_cast<uint4*>(sp[q]) = pre[q];
        __syncwarp();
        if (y0 + 16 < ye) fetch(y0 + 16);
        // staged rows g + ky (i = ky) and g + 8 + ky (i = 5 + ky); word 3 + 4 j + t = pixels x0 - 2 + 8 j + 2 t, + 1
        uint32_t r[10][5];
#pragma unroll
        for (int i = 0; i < 10; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) r[i][j] = fr[(i < 5 ? i : i + 3) * DWM_PITCH + 4 * j];
        float d[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) d[n][e] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 5; ++ky)                          // four independent accumulator chains per ky
#pragma unroll
            for (int n = 0; n < 4; ++n) mma_m16n8k16_f16(d[n], r[ky][n], r[5 + ky][n], r[ky][n + 1], r[5 + ky][n + 1], b0[ky], b1[ky]);
        // the 16 x 32 outputs go back through the (now free) tile so that every lane stores two 16-byte row pieces: 16 store
        // wavefronts per step instead of 64 for 4-byte fragment stores (the LSU data pipe is what bounds this kernel)
        __syncwarp();
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            tile[g * DWM_PITCH + 4 * n + t] = f2_to_h2(fmaxf(d[n][0] + sh, 0.f), fmaxf(d[n][1] + sh, 0.f));
            tile[(g + 8) * DWM_PITCH + 4 * n + t] = f2_to_h2(fmaxf(d[n][2] + sh, 0.f), fmaxf(d[n][3] + sh, 0.f));
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int id = lane + 32 * q, orow = id >> 2, c4 = id & 3;
            const uint4 v = *reinterpret_cast<const uint4*>(tile + orow * DWM_PITCH + 4 * c4);
            if (y0 + orow < ye && x0 + 8 * c4 < G) *reinterpret_cast<uint4*>(op + ((long long)(y0 + orow) * G + x0 + 8 * c4)) = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Fused block for the narrowest refiner (Cp = 32, scale 1: the most pixels): depth-wise 5x5 + batch norm + ReLU as in
// rb_dw_kernel (16 lanes = the 16 channel pairs, two neighbouring strips per warp), then the 1x1 convolution of the same
// pixels on mma.sync without the activations leaving the SM: two finished output rows of a warp (2 x 8 pixels x 32 channels)
// go through a 1.25 KB warp-private shared-memory tile, from which every lane reads its A fragments as two 16-byte rows (the
// k / n permutation of rb_pw_mma_kernel), eight MMAs, bias, and each lane stores 16 contiguous bytes of two output rows.
constexpr int FU_PITCH = 80;      // bytes per pixel row of the tile: conflict-free 4-byte writes, 16-byte aligned reads

__global__ void __launch_bounds__(DW_WARPS * 32, 3) rb_dwpw32_kernel(const __half* __restrict__ in, const float* __restrict__ wf,
                                                                     const float* __restrict__ shift, const __half* __restrict__ w2,
                                                                     const float* __restrict__ b2, __half* __restrict__ out,
                                                                     int G, int th) {
    constexpr int Cp = 32;
    __shared__ uint2 wfrag[8][32];
    __shared__ float bias_s[32];
    __shared__ __align__(16) unsigned char tiles[DW_WARPS][16 * FU_PITCH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 32; i += DW_WARPS * 32) {
        const int L = i & 31, f = i >> 5, j = f & 3, u = f >> 2;
        const int co = 8 * ((L >> 2) >> 1) + 2 * j + ((L >> 2) & 1), ch = 8 * (L & 3) + 4 * u;
        wfrag[f][L] = __ldg(reinterpret_cast<const uint2*>(w2 + co * Cp + ch));
    }
    if (threadIdx.x < 32) bias_s[threadIdx.x] = __ldg(b2 + threadIdx.x);
    __syncthreads();

    const int sx = lane >> 4, c = (lane & 15) * 2;
    const int xw = (blockIdx.x * DW_WARPS + warp) * 8;                       // first pixel of the warp's two strips
    const int x0 = xw + sx * 4, y0 = blockIdx.y * th, b = blockIdx.z;
    const bool cok = x0 < G;
    unsigned long long w[25];
#pragma unroll
    for (int t = 0; t < 25; ++t) w[t] = pack2(__ldg(wf + t * Cp + c), __ldg(wf + t * Cp + c + 1));
    const float sh0 = __ldg(shift + c), sh1 = __ldg(shift + c + 1);
    unsigned long long acc[5][4];
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[s][p] = 0ull;
    const __half* inb = in + (size_t)b * G * G * Cp + c;
    __half* outb = out + (size_t)b * G * G * Cp;
    bool xok[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xok[i] = cok && x0 - 2 + i >= 0 && x0 - 2 + i < G;
    unsigned char* tile = tiles[warp];
    const int g = lane >> 2, t4 = lane & 3;

    // 1x1 stage for output rows oA (tile rows 0-7) and oA + 1 (tile rows 8-15) of this warp's 8 pixels
    auto pointwise = [&](int oA) {
        __syncwarp();
        const uint4 alo = *reinterpret_cast<const uint4*>(tile + g * FU_PITCH + t4 * 16);
        const uint4 ahi = *reinterpret_cast<const uint4*>(tile + (g + 8) * FU_PITCH + t4 * 16);
        float d[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) d[j][e] = 0.f;
            const uint2 bu0 = wfrag[j][lane], bu1 = wfrag[4 + j][lane];
            mma_m16n8k16_f16(d[j], alo.x, ahi.x, alo.y, ahi.y, bu0.x, bu0.y);
            mma_m16n8k16_f16(d[j], alo.z, ahi.z, alo.w, ahi.w, bu1.x, bu1.y);
        }
        const float4 bl = *reinterpret_cast<const float4*>(bias_s + 8 * t4), bh = *reinterpret_cast<const float4*>(bias_s + 8 * t4 + 4);
        uint4 o0, o1;
        o0.x = f2_to_h2(d[0][0] + bl.x, d[0][1] + bl.y); o1.x = f2_to_h2(d[0][2] + bl.x, d[0][3] + bl.y);
        o0.y = f2_to_h2(d[1][0] + bl.z, d[1][1] + bl.w); o1.y = f2_to_h2(d[1][2] + bl.z, d[1][3] + bl.w);
        o0.z = f2_to_h2(d[2][0] + bh.x, d[2][1] + bh.y); o1.z = f2_to_h2(d[2][2] + bh.x, d[2][3] + bh.y);
        o0.w = f2_to_h2(d[3][0] + bh.z, d[3][1] + bh.w); o1.w = f2_to_h2(d[3][2] + bh.z, d[3][3] + bh.w);
        const int x = xw + g, yA = y0 + oA, yB = yA + 1;
        if (x < G && yA < G) *reinterpret_cast<uint4*>(outb + ((size_t)yA * G + x) * Cp + 8 * t4) = o0;
        if (x < G && oA + 1 < th && yB < G) *reinterpret_cast<uint4*>(outb + ((size_t)yB * G + x) * Cp + 8 * t4) = o1;
        __syncwarp();
    };

    uint32_t raw[8];
    const __half* row = inb + ((long long)(y0 - 2) * G + (x0 - 2)) * Cp;
    const long long row_pitch = (long long)G * Cp;
    auto load_row = [&](int j) {
        const int iy = y0 - 2 + j;
        const bool yok = iy >= 0 && iy < G && j < th + 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            raw[i] = 0u;
            if (yok && xok[i]) raw[i] = __ldg(reinterpret_cast<const uint32_t*>(row + i * Cp));
        }
        row += row_pitch;
    };
    load_row(0);
#pragma unroll 1
    for (int base = 0; base < th + 4; base += 5) {
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int j = base + s;
            unsigned long long v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = h2_to_f2(raw[i]);
            load_row(j + 1);
#pragma unroll
            for (int ky = 0; ky < 5; ++ky) {
                const int slot = (s - ky + 5) % 5;
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int kx = 0; kx < 5; ++kx) acc[slot][p] = fma2(w[ky * 5 + kx], v[p + kx], acc[slot][p]);
            }
            const int o = j - 4, slot = (s + 1) % 5;
            if (o >= 0) {                                 // finished output row o -> tile rows (o & 1) * 8 + sx * 4 + p
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float2 a = unpack2(acc[slot][p]);
                    *reinterpret_cast<uint32_t*>(tile + ((o & 1) * 8 + sx * 4 + p) * FU_PITCH + c * 2) =
                        f2_to_h2(fmaxf(a.x + sh0, 0.f), fmaxf(a.y + sh1, 0.f));
                }
                if (o & 1) pointwise(o - 1);
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) acc[slot][p] = 0ull;
        }
    }
    if (th & 1) pointwise(th - 1);                         // odd strip height: the last row has no partner
}

static int launch_dwpw32(const __half* in, const float* wf, const float* shift, const __half* w2, const float* b2, __half* out,
                         int B, int G, cudaStream_t st) {
    const int xblocks = (G + 8 * DW_WARPS - 1) / (8 * DW_WARPS);
    int th = DW_TH_MIN;
    long long best = (long long)((G + th - 1) / th) * (th + 4);
    for (int t = DW_TH_MIN + 5; t <= G + 4; t += 5) {
        const int segs = (G + t - 1) / t;
        if ((long long)xblocks * segs * B < 2 * 3 * 148) break;
        if ((long long)segs * (t + 4) < best) { best = (long long)segs * (t + 4); th = t; }
    }
    dim3 grid(xblocks, (G + th - 1) / th, B);
    rb_dwpw32_kernel<<<grid, DW_WARPS * 32, 0, st>>>(in, wf, shift, w2, b2, out, G, th);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// out_conv: fp32 1x1 convolution hidden -> OC (<= 4) on the fp16 activations, NCHW fp32 output.  Four lanes per pixel, each
// reading 16-byte pieces of the NHWC row (a warp load covers 8 rows x 64 contiguous bytes), weights broadcast from shared memory.
__global__ void __launch_bounds__(256) rb_out_kernel(const __half* __restrict__ act, const float* __restrict__ w /*[OC][Cp]*/,
                                                     const float* __restrict__ bias, float* __restrict__ out /*[B,OC,P]*/,
                                                     int B, int P, int Cp, int OC) {
    extern __shared__ float ws[];                           // [4][Cp], rows >= OC zero
    for (int i = threadIdx.x; i < 4 * Cp; i += 256) ws[i] = i < OC * Cp ? __ldg(w + i) : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, t = lane & 3;
    const long long total = (long long)B * P;
    // warp-uniform trip count (the shuffles below need the whole warp): 8 pixels per warp and round
    for (long long w0 = ((long long)blockIdx.x * 256 + (threadIdx.x & ~31)) >> 2; w0 < total; w0 += ((long long)gridDim.x * 256) >> 2) {
        const long long pix = w0 + (lane >> 2);
        const bool ok = pix < total;
        const __half* row = act + (size_t)(ok ? pix : 0) * Cp;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 8 * t; c < Cp && ok; c += 32) {
            const uint4 v = __ldcs(reinterpret_cast<const uint4*>(row + c));
            const uint32_t hv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hv[e]));
#pragma unroll
                for (int o = 0; o < 4; ++o) acc[o] = fmaf(f.y, ws[o * Cp + c + 2 * e + 1], fmaf(f.x, ws[o * Cp + c + 2 * e], acc[o]));
            }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 1);
            acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], 2);
        }
        const long long bb = pix / P, p = pix - bb * P;
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (ok && o == t && o < OC) out[((size_t)bb * OC + o) * P + p] = acc[o] + __ldg(bias + o);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// decoder-loop glue (model/network.py:262-285)
// displacement = int(scale) * (delta / (4 W0), delta / (4 H0)) as torch evaluates it on CUDA (division by a Python scalar =
// multiplication by its fp32 reciprocal); eval-mode zeroing of |disp - pre| / |pre| < 1e-6; flow += disp; certainty += delta_c
__global__ void __launch_bounds__(256) flow_update_kernel(const float* __restrict__ delta /*[B,3,P]*/, float* __restrict__ flow,
                                                          float* __restrict__ cert, float* __restrict__ pre, int B, int P,
                                                          float scale, float inv4w, float inv4h, int zero_rule) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)B * P) return;
    const long long b = i / P, p = i - b * P;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const float dl = delta[((size_t)b * 3 + ch) * P + p];
        float disp = __fmul_rn(scale, __fmul_rn(dl, ch == 0 ? inv4w : inv4h));
        const size_t o = ((size_t)b * 2 + ch) * P + p;
        const float dp = pre[o];
        if (zero_rule && __fdiv_rn(fabsf(__fsub_rn(disp, dp)), fabsf(dp)) < 1e-6f) disp = 0.f;
        flow[o] = __fadd_rn(flow[o], disp);
        pre[o] = disp;
    }
    cert[(size_t)b * P + p] = __fadd_rn(cert[(size_t)b * P + p], delta[((size_t)b * 3 + 2) * P + p]);
}

// F.interpolate(mode="bilinear", align_corners=False, size=(Ho, Wo)) as ATen's upsample_bilinear2d evaluates it
__global__ void __launch_bounds__(256) upsample_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out, int planes,
                                                                int Hi, int Wi, int Ho, int Wo, float rh, float rw) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)planes * Ho * Wo) return;
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
    const long long pl = i / ((long long)Wo * Ho);
    const float sy = fmaxf(__fmaf_rn(rh, (float)oy + 0.5f, -0.5f), 0.f), sx = fmaxf(__fmaf_rn(rw, (float)ox + 0.5f, -0.5f), 0.f);
    const int y1 = (int)sy, x1 = (int)sx;
    const int yp = y1 < Hi - 1 ? 1 : 0, xp = x1 < Wi - 1 ? 1 : 0;
    const float ly = sy - (float)y1, lx = sx - (float)x1, hy = 1.f - ly, hx = 1.f - lx;
    const float* s = in + (size_t)pl * Hi * Wi;
    const float v00 = s[(size_t)y1 * Wi + x1], v01 = s[(size_t)y1 * Wi + x1 + xp];
    const float v10 = s[(size_t)(y1 + yp) * Wi + x1], v11 = s[(size_t)(y1 + yp) * Wi + x1 + xp];
    out[i] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
}

// ---------------------------------------------------------------------------------------------------------------------
static inline int pad16(int c) { return (c + 15) / 16 * 16; }

static int launch_pack(const float* d, __half* out, int B, int C, int Cp, int P, cudaStream_t st) {
    dim3 grid((P + 31) / 32, (Cp + 63) / 64, B);
    rb_pack_kernel<<<grid, 256, 0, st>>>(d, out, C, Cp, P);
    return (int)cudaGetLastError();
}
static int launch_dw(const __half* in, const float* wf, const float* shift, __half* out, int B, int G, int Cp, cudaStream_t st) {
    // lanes per chunk: the widest of 32 / 16 / 8 channel pairs that wastes at most ~5 % of the lanes on channel padding
    const int pairs = Cp / 2;
    int lp = 32;
    while (lp > 8 && ((pairs + lp - 1) / lp) * lp * 20 > pairs * 21) lp /= 2;
    const int nchunk = (pairs + lp - 1) / lp, spw = 32 / lp;
    const int xblocks = (G + 4 * DW_WARPS * spw - 1) / (4 * DW_WARPS * spw);
    // strip height: every strip re-reads 4 halo rows, so taller strips cost less -- as long as the grid still fills the SMs
    // (3 blocks per SM, at least two waves) and the last strip of a column is not mostly empty
    int th = DW_TH_MIN;
    long long best = (long long)((G + th - 1) / th) * (th + 4);
    for (int t = DW_TH_MIN + 5; t <= G + 4; t += 5) {
        const int segs = (G + t - 1) / t;
        if ((long long)xblocks * nchunk * segs * B < 2 * 3 * 148) break;
        if ((long long)segs * (t + 4) < best) { best = (long long)segs * (t + 4); th = t; }
    }
    dim3 grid(xblocks * nchunk, (G + th - 1) / th, B);
    // (four blocks per SM under a 128-register cap -- small spills -- measured 9 % slower than three at 168 registers)
#define DW_LAUNCH(CPV) rb_dw_kernel<CPV><<<grid, DW_WARPS * 32, 0, st>>>(in, wf, shift, out, G, Cp, nchunk, lp, th)
    switch (Cp) {
        case 32: DW_LAUNCH(32); break;
        case 80: DW_LAUNCH(80); break;
        case 192: DW_LAUNCH(192); break;
        case 368: DW_LAUNCH(368); break;
        case 432: DW_LAUNCH(432); break;
        default: DW_LAUNCH(0); break;
    }
    return (int)cudaGetLastError();
}
static int launch_pw(const __half* act, const __half* w2, const float* bias, __half* out, long long P, int Cp, int algo,
                     cudaStream_t st) {
    if (algo == 1) {
        rb_pw_simt_kernel<<<(unsigned)P, 128, 0, st>>>(act, w2, bias, out, (size_t)P, Cp);
        return (int)cudaGetLastError();
    }
    if (algo == 0 && Cp <= 96) {                 // narrow refiners: the streaming mma.sync kernel
        if (Cp <= 32) return launch_pw_mma<1>(act, w2, bias, out, P, Cp, st);
        if (Cp <= 64) return launch_pw_mma<2>(act, w2, bias, out, P, Cp, st);
        return launch_pw_mma<3>(act, w2, bias, out, P, Cp, st);
    }
    PwParams g;
    g.bias = bias; g.out = out; g.P = P; g.Cp = Cp;
    // N tiles of at most 128 columns keep two accumulators in flight (epilogue of one tile under the MMAs of the next); measured
    // 4 % faster at Cp = 192, but not worth a fourfold re-read of the activations at Cp = 368 / 432 (two halves there)
    const int nt_max = Cp <= 256 ? 128 : 256;
    g.nsplit = (Cp + nt_max - 1) / nt_max;
    g.NT = pad16((Cp + g.nsplit - 1) / g.nsplit);
    g.KA = (Cp + 63) / 64;
    g.klast = (Cp - 64 * (g.KA - 1)) / 16;
    const int stage_bytes = 128 * 128 + g.NT * 128;
    g.nstage = max(2, min(PW_MAXSTAGE, (100 * 1024) / stage_bytes));
    const long long mtiles = (P + 127) / 128;
    if (mtiles * g.nsplit > 0x7fffffffLL) return GFB_EUNSUPPORTED;
    g.nwork = (int)(mtiles * g.nsplit);
    CUtensorMap tmA, tmB;
    {   // fp16 pairs addressed as 32-bit words: inner extent Cp / 2, box 32 words = one 128-byte swizzle row
        uint64_t dims[2] = {(uint64_t)Cp / 2, (uint64_t)P};
        uint64_t strides[1] = {(uint64_t)Cp * 2};
        uint32_t box[2] = {32u, 128u};
        int rc = gfb_encode_tmap_f32(&tmA, act, 2, dims, strides, box, 3);
        if (rc != GFB_OK) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)Cp / 2, (uint64_t)Cp};
        uint64_t strides[1] = {(uint64_t)Cp * 2};
        uint32_t box[2] = {32u, (uint32_t)g.NT};
        int rc = gfb_encode_tmap_f32(&tmB, w2, 2, dims, strides, box, 3);
        if (rc != GFB_OK) return rc;
    }
    const int smem = g.nstage * stage_bytes + 1024;
    int dev = 0;
    cudaGetDevice(&dev);
    static int sms_of[64] = {0};           // immutable per-device cache (SM count; the attribute is set once per device)
    if (dev < 0 || dev >= 64) return GFB_EUNSUPPORTED;
    if (sms_of[dev] == 0) {
        cudaError_t e = cudaFuncSetAttribute(rb_pw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
        if (e != cudaSuccess) return (int)e;
        int sms = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess || sms <= 0) return e != cudaSuccess ? (int)e : GFB_ENODEVICE;
        sms_of[dev] = sms;
    }
    const int grid = (int)min((long long)g.nwork, (long long)2 * sms_of[dev]);
    rb_pw_kernel<<<grid, PW_THREADS, smem, st>>>(tmA, tmB, g);
    return (int)cudaGetLastError();
}
static int launch_out(const __half* act, const float* w, const float* bias, float* out, int B, int P, int Cp, int OC, cudaStream_t st) {
    const long long total = (long long)B * P;
    const long long want = (total * 4 + 255) / 256;
    const int grid = (int)min(want, (long long)148 * 8 * 4);          // grid-stride over the pixels: the weights are staged once per block
    rb_out_kernel<<<grid, 256, 4 * Cp * sizeof(float), st>>>(act, w, bias, out, B, P, Cp, OC);
    return (int)cudaGetLastError();
}

// packed weights of one refiner (host-side packer: gfnet_b200/refiner.py): per block
//   [wf 25*Cp f32][shift Cp f32][b2 Cp f32][w2 Cp*Cp f16]   then   [wout OC*Cp f32][bout 4 f32]
static inline size_t block_bytes(int Cp) { return (size_t)27 * Cp * 4 + (size_t)Cp * Cp * 2; }

}  // namespace rb
}  // namespace gfb

using namespace gfb;
using namespace gfb::rb;

#define RB_TRY(x) do { int rc__ = (x); if (rc__ != 0) return rc__; } while (0)

extern "C" int gfb_refiner_pack_f16(const float* d, void* out, int B, int C, int P, gfb_stream_t stream) {
    GFB_CHECK_ARG(d && out && B > 0 && B <= 65535 && C > 0 && P > 0);
    return launch_pack(d, (__half*)out, B, C, pad16(C), P, gfb_cu(stream));
}

extern "C" int gfb_refiner_dw5_f16(const void* in, const float* wf, const float* shift, void* out, int B, int G, int Cp,
                                   gfb_stream_t stream) {
    GFB_CHECK_ARG(in && wf && shift && out && in != out && B > 0 && B <= 65535 && G > 0 && Cp > 0 && Cp % 16 == 0);
    return launch_dw((const __half*)in, wf, shift, (__half*)out, B, G, Cp, gfb_cu(stream));
}

// EXPERIMENT: channel-planar depth-wise stage on the tensor cores; in / out [B*C][G][G] fp16, G % 8 == 0
extern "C" int gfb_debug_refiner_dw5_planar_f16(const void* in, const float* wf, const float* shift, void* out, int B, int C, int Cp,
                                                int G, gfb_stream_t stream) {
    GFB_CHECK_ARG(in && wf && shift && out && in != out && B > 0 && C > 0 && Cp >= C && G > 0 && G % 8 == 0);
    GFB_CHECK_ARG((long long)B * C <= 65535);
    const int seg = G >= 256 ? 128 : ((G + 15) / 16) * 16;          // rows per block: whole columns unless the plane is tall
    dim3 grid((G + 127) / 128, (G + seg - 1) / seg, B * C);
    rb_dwm_kernel<<<grid, 128, 0, gfb_cu(stream)>>>((const __half*)in, wf, shift, (__half*)out, C, Cp, G, seg);
    GFB_LAUNCH_RESULT();
}

extern "C" int gfb_refiner_pw_f16(const void* act, const void* w2, const float* bias, void* out, long long P, int Cp, int algo,
                                  gfb_stream_t stream) {
    GFB_CHECK_ARG(act && w2 && bias && out && act != out && P > 0 && P < (1ll << 31) - 128 && Cp > 0 && Cp % 16 == 0 && Cp <= 512);
    GFB_CHECK_ARG(algo >= 0 && algo <= 2);
    if (!gfb_aligned(act, 16) || !gfb_aligned(w2, 16) || !gfb_aligned(out, 16) || !gfb_aligned(bias, 16)) return GFB_EALIGN;
    return launch_pw((const __half*)act, (const __half*)w2, bias, (__half*)out, P, Cp, algo, gfb_cu(stream));
}

extern "C" int gfb_refiner_out_f32(const void* act, const float* w, const float* bias, float* out, int B, int P, int Cp, int OC,
                                   gfb_stream_t stream) {
    GFB_CHECK_ARG(act && w && bias && out && B > 0 && P > 0 && Cp > 0 && Cp % 16 == 0 && OC > 0 && OC <= 4);
    return launch_out((const __half*)act, w, bias, out, B, P, Cp, OC, gfb_cu(stream));
}

extern "C" size_t gfb_refiner_blocks_weight_bytes(int C, int nblocks, int out_dim) {
    const int Cp = pad16(C);
    return (size_t)nblocks * block_bytes(Cp) + (size_t)out_dim * Cp * 4 + 16;
}

extern "C" int gfb_refiner_blocks_chunk(int B, int C, int G) {
    // The whole batch at once unless an activation buffer would exceed 1 GiB.  Chunks sized for the L2 (two 24 MB buffers) were
    // measured 25-45 % slower at op batch 64: short grids lose to wave quantisation and the kernels are not HBM-bound.
    const size_t per = (size_t)G * G * pad16(C) * 2;
    size_t n = ((size_t)1 << 30) / per;
    if (n < 1) n = 1;
    if (n > (size_t)B) n = (size_t)B;
    const size_t nchunks = ((size_t)B + n - 1) / n;          // equal chunks
    return (int)(((size_t)B + nchunks - 1) / nchunks);
}

extern "C" size_t gfb_refiner_blocks_workspace_bytes(int B, int C, int G, int chunk) {
    if (chunk <= 0) chunk = gfb_refiner_blocks_chunk(B, C, G);
    return 2 * (size_t)chunk * G * G * pad16(C) * 2 + 512;
}

// d [B,C,G,G] fp32 (the refiner input of model/network.py:555) -> out [B,out_dim,G,G] fp32 = out_conv(hidden_blocks(block1(d)))
extern "C" int gfb_refiner_blocks_f16(const float* d, const void* weights, float* out, int B, int C, int G, int nblocks,
                                      int out_dim, void* workspace, size_t ws_bytes, int chunk, int algo, gfb_stream_t stream) {
    GFB_CHECK_ARG(d && weights && out && workspace && B > 0 && B <= 65535 && C > 0 && C <= 512 && G > 0 && G <= 4096);
    GFB_CHECK_ARG(nblocks > 0 && out_dim > 0 && out_dim <= 4 && algo >= 0 && algo <= 2);
    if (chunk <= 0) chunk = gfb_refiner_blocks_chunk(B, C, G);
    if (chunk > B) chunk = B;
    if (ws_bytes < gfb_refiner_blocks_workspace_bytes(B, C, G, chunk)) return GFB_EWORKSPACE;
    if (!gfb_aligned(weights, 16)) return GFB_EALIGN;
    const int Cp = pad16(C), P = G * G;
    cudaStream_t st = gfb_cu(stream);
    unsigned char* ws = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    const size_t buf_bytes = (size_t)chunk * P * Cp * 2;
    __half* h = reinterpret_cast<__half*>(ws);
    __half* a = reinterpret_cast<__half*>(ws + buf_bytes);
    const unsigned char* wb = reinterpret_cast<const unsigned char*>(weights);
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = min(chunk, B - b0);
        RB_TRY(launch_pack(d + (size_t)b0 * C * P, h, nb, C, Cp, P, st));
        for (int k = 0; k < nblocks; ++k) {
            const unsigned char* blk = wb + (size_t)k * block_bytes(Cp);
            const float* wf = reinterpret_cast<const float*>(blk);
            const float* shift = wf + 25 * Cp;
            const float* b2 = shift + Cp;
            const __half* w2 = reinterpret_cast<const __half*>(b2 + Cp);
            if (Cp == 32 && algo == 0) {          // scale 1: depth-wise and 1x1 stages in one kernel, buffers swap roles
                RB_TRY(launch_dwpw32(h, wf, shift, w2, b2, a, nb, G, st));
                __half* tmp = h; h = a; a = tmp;
                continue;
            }
            RB_TRY(launch_dw(h, wf, shift, a, nb, G, Cp, st));
            RB_TRY(launch_pw(a, w2, b2, h, (long long)nb * P, Cp, algo, st));
        }
        const float* wout = reinterpret_cast<const float*>(wb + (size_t)nblocks * block_bytes(Cp));
        RB_TRY(launch_out(h, wout, wout + (size_t)out_dim * Cp, out + (size_t)b0 * out_dim * P, nb, P, Cp, out_dim, st));
    }
    return GFB_OK;
}

extern "C" int gfb_flow_update_f32(const float* delta, float* flow, float* certainty, float* disp_pre, int B, int G,
                                   int scale, int H0, int W0, int zero_rule, gfb_stream_t stream) {
    GFB_CHECK_ARG(delta && flow && certainty && disp_pre && B > 0 && G > 0 && H0 > 0 && W0 > 0);
    const long long n = (long long)B * G * G;
    flow_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, gfb_cu(stream)>>>(delta, flow, certainty, disp_pre, B, G * G,
                                                                              (float)scale, 1.0f / (float)(4 * W0),
                                                                              1.0f / (float)(4 * H0), zero_rule);
    GFB_LAUNCH_RESULT();
}

extern "C" int gfb_upsample_bilinear_f32(const float* in, float* out, int planes, int Hi, int Wi, int Ho, int Wo,
                                         gfb_stream_t stream) {
    GFB_CHECK_ARG(in && out && in != out && planes > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0);
    const long long n = (long long)planes * Ho * Wo;
    upsample_bilinear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, gfb_cu(stream)>>>(in, out, planes, Hi, Wi, Ho, Wo,
                                                                                    (float)Hi / (float)Ho, (float)Wi / (float)Wo);
    GFB_LAUNCH_RESULT();
}
