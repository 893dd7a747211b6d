/*
 * gfnet_b200 -- C ABI of the B200-native GFNet hot path (dense matching + homography).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every pointer is a
 * DEVICE pointer unless said otherwise; tensors are contiguous in the stated layout; `stream` is
 * a cudaStream_t (0 = legacy default stream).  The library never allocates device memory, never
 * synchronises the device and never touches a stream other than the one passed.  All entry
 * points are re-entrant (no mutable global state except an immutable driver-entry-point cache).
 *
 * Return convention: 0 = OK, <0 = GFB_E* (bad/unsupported argument, nothing launched),
 * >0 = the cudaError_t of the failed launch.  `gfb_strerror` explains both.
 *
 * Each function cites the reference interface (KN-Zhang/GFNet @ 2281c4b) it replaces.
 */
#ifndef GFNET_B200_H
#define GFNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFB_ABI_VERSION 2

#define GFB_OK 0
#define GFB_EINVAL (-1)       /* bad shape / null pointer / bad enum          */
#define GFB_EUNSUPPORTED (-2) /* valid in the reference, not implemented here */
#define GFB_EALIGN (-3)       /* pointer not aligned as required              */
#define GFB_EWORKSPACE (-4)   /* workspace too small                          */
#define GFB_ENODEVICE (-5)    /* no sm_100 device / driver entry point missing */

typedef void* gfb_stream_t; /* cudaStream_t */

int gfb_abi_version(void);
const char* gfb_strerror(int code);
/* sm count and compute capability of the current device (host-side query). */
int gfb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- K1: local correlation --------------------------------------------------------------
 * replaces utils/local_correlation.py:4-72 `local_correlation(featuremap_size, feature0,
 * feature1, local_radius, num_grid, padding_mode, flow, im_A_coords, sample_mode,
 * grid_based_correlation, num_level)`; call site model/network.py:553-554.
 *   f0   [B,C,G,G]   grid features of image A        (fp32, NCHW)
 *   f1   [B,C,Hs,Ws] feature map of image B          (fp32, NCHW)
 *   flow [B,2,G,G]   normalised (x,y) targets in B   (fp32; ch0 = x, ch1 = y)
 *   out  [B,Ktot,G,G]; this call writes channels [k_offset, k_offset + (2r+1)^2)
 *        out[b, k_offset + iy*(2r+1)+ix, gy, gx] =
 *            (1/sqrt(C)) * sum_c f0[b,c,gy,gx] * sample(f1[b,c], flow[b,:,gy,gx] + off(ix,iy))
 *   off(ix,iy) = ((ix-r)*2/win_w, (iy-r)*2/win_h) built like torch.linspace(-2r/n, 2r/n, 2r+1);
 *        win_w/win_h = (Ws,Hs) of `featuremap_size` normally, or num_grid when
 *        grid_based_correlation=True (local_correlation.py:33-52).
 *   sample_mode 0 = bilinear, 1 = nearest; padding_mode 0 = zeros, 1 = border; align_corners=False.
 *   f1_pitch: floats between consecutive rows of f1 (0 = Ws).
 * This general entry is the per-sample gather kernel (any C, r, mode); the hot kernels for the configuration GFNet uses
 * are below.  num_level > 1 (local_correlation.py:61-71) = one call per level with `gfb_avg_pool2_f32` between. */
int gfb_local_corr_f32(const float* f0, const float* f1, const float* flow, float* out,
                       int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                       int win_h, int win_w, int sample_mode, int padding_mode,
                       int k_total, int k_offset, gfb_stream_t stream);
/* Hot kernels for the same operator (utils/local_correlation.py:4-72; bilinear, zero padding, window == (Ws,Hs) only --
 * the configuration GFNet uses, model/network.py:553-554).  Same tensors as gfb_local_corr_f32.
 *
 * gfb_local_corr_pt_f32: one lattice point per thread, the whole (2r+2)^2 integer patch of dot products in registers, f1
 *   staged per tile by TMA (zero fill outside the image).  (r, C) = (2,16): lc_rot_kernel (window cells visited in a
 *   rotated order that makes every shared-memory load bank-conflict free; producer / consumer warps); (4,32), (1,16),
 *   (1,8), (2,8): lc_pt_kernel.  f1_pitch % 4 == 0, f0 / f1 16-byte aligned (TMA strides) else GFB_EALIGN; G % 4 == 0 else
 *   GFB_EUNSUPPORTED.  tune 0 = default kernel and box sizes (other values select tilings, see local_corr_v2.cu).
 *
 * gfb_local_corr_tc2_f32: banded GEMM on tcgen05 fed by TMA.  A pre-pass rewrites f0 / f1 once as position-major
 *   rows [bf16 hi(C) | bf16 lo(C)] into `workspace` (processed in groups of `group` batch elements; 0 = auto), the main
 *   kernel accumulates hi*hi + hi*lo + lo*hi in fp32 (relative error ~1e-5, inside the 1e-4 fp32 bar).
 *   (r, C) in {(2,32), (3,32), (4,32), (2,64), (3,64), (4,64), (5,64), (6,64), (7,64), (8,64)}; any Ws / pitch.
 *   workspace >= gfb_local_corr_tc2_workspace_bytes(...) bytes, 128-byte aligned.
 * gfb_local_corr_tc2_slice_f32: the same kernel on the channel slice [c0, c0 + C) of tensors with Ctot channels, scaled by
 *   1/sqrt(Ctot), stored (accumulate = 0) or added to out (accumulate = 1): correlation is linear in the channels, so the
 *   C-channel kernels cover any Ctot that is a multiple of C (C = 128 .. 512 of the microbench sweep).
 * All return GFB_EUNSUPPORTED for other (r, C); points whose windows leave the staged box use the exact gather. */
int gfb_local_corr_pt_f32(const float* f0, const float* f1, const float* flow, float* out,
                          int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                          int k_total, int k_offset, int tune, gfb_stream_t stream);
size_t gfb_local_corr_tc2_workspace_bytes(int B, int C, int Hs, int Ws, int G, int r, int group);
int gfb_local_corr_tc2_f32(const float* f0, const float* f1, const float* flow, float* out,
                           int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                           int k_total, int k_offset, int group,
                           void* workspace, size_t workspace_bytes, gfb_stream_t stream);
int gfb_local_corr_tc2_slice_f32(const float* f0, const float* f1, const float* flow, float* out,
                                 int B, int C, int Ctot, int c0, int accumulate, int Hs, int Ws, int f1_pitch, int G, int r,
                                 int k_total, int k_offset, void* workspace, size_t workspace_bytes, gfb_stream_t stream);
/* Split form for callers that correlate the same feature0 / feature1 against several flows -- the iterations of one
 * refiner scale (model/network.py:230-281 calls local_correlation num_itr times per scale with unchanged features):
 * `prepare` runs the pre-pass into `workspace` (whole batch in one group, else GFB_EUNSUPPORTED), `run` plans and
 * correlates one flow against it; f0 / f1 are still passed to `run` for the exact-gather fallback. */
int gfb_local_corr_tc2_prepare_f32(const float* f0, const float* f1, int B, int C, int Hs, int Ws, int f1_pitch,
                                   int G, int r, void* workspace, size_t workspace_bytes, gfb_stream_t stream);
int gfb_local_corr_tc2_run_f32(const float* f0, const float* f1, const float* flow, float* out,
                               int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                               int k_total, int k_offset,
                               void* workspace, size_t workspace_bytes, gfb_stream_t stream);
/* gfb_local_corr_mma_f32: the same operator on the warp-level tensor-core path (mma.sync m16n8k16, bf16 hi/lo split of the
 *   fp32 operands done in-kernel: hi*hi + hi*lo + lo*hi, fp32 accumulate, relative error ~1e-5) -- reads f0 / f1 / flow as
 *   they are (fp32 NCHW), no pre-pass, no workspace.  A warp owns 8 x 4 lattice points, the CTA streams the image rows of its
 *   bounding box through shared memory (cp.async, zero fill), every lane bilerps and stores its own point.
 *   Channel-slice form like gfb_local_corr_tc2_slice_f32: channels [c0, c0 + C) of tensors with Ctot channels, scaled by
 *   1/sqrt(Ctot), stored (accumulate = 0) or added (accumulate = 1); Ctot = C (or 0), c0 = 0, accumulate = 0 is the plain
 *   operator.  (r, C) in {(1,16), (2,16), (2,32), (3,32), (4,32), (2..8, 64)}; f1_pitch % 4 == 0 and f1 16-byte aligned else
 *   GFB_EALIGN.  shape 0 = auto CTA shape, else warps_x | warps_y << 4 (4|2<<4, 4|1<<4, 2|1<<4).
 *   gfb_debug_local_corr_mma_counters (synchronises): host_out4[0] = points that took the exact gather. */
int gfb_local_corr_mma_f32(const float* f0, const float* f1, const float* flow, float* out,
                           int B, int C, int Ctot, int c0, int accumulate, int Hs, int Ws, int f1_pitch, int G, int r,
                           int k_total, int k_offset, int shape, gfb_stream_t stream);
int gfb_debug_local_corr_mma_counters(unsigned long long* host_out4, int reset);
/* ---- refiner input assembly -------------------------------------------------
 * replaces the body of ConvRefiner.forward up to the concatenation, model/network.py:537-555:
 *   d = cat(grid_sample(x, lattice), grid_sample(y, flow), disp_emb(40/32 * scale_factor * (flow - lattice)),
 *           local_correlation(grid_feature, y, flow))                       d [B, 2C + dd + (2r+1)^2, G, G]
 * gfb_refiner_assemble_f32 writes d[:, 0 : 2C + dd] (bilinear, zeros, align_corners=False like F.grid_sample :537, :547;
 *   the 1x1 conv `disp_emb` :549 as w [dd,2] + bias [dd]; emb_scale = 40/32 * scale_factor); x [B,C,Hs,Ws], y [B,C,Hs,y_pitch];
 *   keep_grid != 0 leaves d[:, 0:C] as it is (later refiner iterations of one scale: same x, same grid features).
 * gfb_local_corr_cat_f32 correlates feature0 = d[:, 0:C] with f1 along `flow` and writes d[:, k_offset : k_offset + (2r+1)^2]
 *   (the call of :553-554 with its output placed where torch.cat :555 would copy it).  Kernel choice: the point kernels for
 *   their (r, C), the mma.sync kernel for C = 32, the tcgen05 kernel for C = 64 when a workspace is given (phase 0 = one
 *   shot, 1 = pre-pass only, 2 = plan + main on the prepared workspace: the iterations of one scale) else mma.sync. */
int gfb_refiner_assemble_f32(const float* x, const float* y, const float* flow, const float* w, const float* bias,
                             float* d, int B, int C, int Hs, int Ws, int y_pitch, int G, int dd, int Dtot,
                             float emb_scale, int keep_grid, gfb_stream_t stream);
int gfb_local_corr_cat_f32(float* d, int Dtot, const float* f1, const float* flow,
                           int B, int C, int Hs, int Ws, int f1_pitch, int G, int r, int k_offset,
                           int phase, void* workspace, size_t workspace_bytes, gfb_stream_t stream);
/* ---- refiner convolution blocks + decoder-loop glue ---------------------------
 * replaces the tail of ConvRefiner.forward, model/network.py:557-563 (`d = block1(d); d = hidden_blocks(d)` under fp16
 * autocast, `d = out_conv(d.float())`), each block = create_block :505-531 with dw=True: depth-wise 5x5 conv -> BatchNorm2d
 * (eval) -> ReLU -> 1x1 conv.  Activations are fp16 NHWC [B, G*G, Cp] with Cp = C rounded up to 16 (pad channels zero), all
 * sums accumulate in fp32 (the reference rounds to fp16 after every operator; stated tolerance in tests/test_refiner_blocks.py).
 *   gfb_refiner_pack_f16   d [B,C,P] fp32 NCHW -> out [B,P,Cp] fp16
 *   gfb_refiner_dw5_f16    out = relu(dwconv5x5(in) * bn_scale + shift): wf [25][Cp] fp32 (tap-major, bn_scale folded),
 *                          shift [Cp] = (conv_bias - running_mean) * bn_scale + bn_bias; zero padding 2
 *   gfb_refiner_pw_f16     out[p,n] = fp16(sum_k act[p,k] w2[n,k] + bias[n]); w2 [Cp][Cp] fp16; algo 0 = auto (Cp <= 96: streaming
 *                          mma.sync kernel, above: tcgen05, TMA-fed, TMEM accumulator), 2 = tcgen05 at every width,
 *                          1 = CUDA-core cross-check; all pointers 16-byte aligned
 *   gfb_refiner_out_f32    out [B,OC,P] fp32 = w [OC][Cp] . act + bias   (out_conv on d.float(), OC <= 4)
 *   gfb_refiner_blocks_f16 the whole tail for d [B,C,G,G] -> out [B,out_dim,G,G] over two ping-pong activation buffers in the
 *                          workspace (`chunk` batch elements at a time, 0 = all unless a buffer exceeds 1 GiB); weights = the packed blob
 *                          per block [wf 25*Cp f32][shift Cp f32][b2 Cp f32][w2 Cp*Cp f16], then [wout out_dim*Cp f32][bout 4 f32]
 *                          (gfnet_b200/refiner.py packs it from a ConvRefiner); chunk 0 = gfb_refiner_blocks_chunk.
 * gfb_flow_update_f32: the loop body model/network.py:265-274 for delta [B,3,G,G] = (dx, dy, d_certainty):
 *   disp = int(scale) * (dx / (4 W0), dy / (4 H0)) evaluated as torch does on CUDA, zero_rule != 0: disp[|disp - pre| / |pre|
 *   < 1e-6] = 0 (:270-271, eval mode), flow += disp, certainty += d_certainty, disp_pre = disp.
 * gfb_upsample_bilinear_f32: F.interpolate(x, size=(Ho,Wo), mode="bilinear", align_corners=False) on [planes,Hi,Wi]
 *   (:238-249, 276-285). */
int gfb_refiner_pack_f16(const float* d, void* out, int B, int C, int P, gfb_stream_t stream);
int gfb_refiner_dw5_f16(const void* in, const float* wf, const float* shift, void* out, int B, int G, int Cp,
                        gfb_stream_t stream);
/* experiment: the depth-wise stage on channel-planar fp16 [B*C][G][G] as banded MMAs */
int gfb_debug_refiner_dw5_planar_f16(const void* in, const float* wf, const float* shift, void* out, int B, int C, int Cp,
                                     int G, gfb_stream_t stream);
int gfb_refiner_pw_f16(const void* act, const void* w2, const float* bias, void* out, long long P, int Cp, int algo,
                       gfb_stream_t stream);
int gfb_refiner_out_f32(const void* act, const float* w, const float* bias, float* out, int B, int P, int Cp, int OC,
                        gfb_stream_t stream);
size_t gfb_refiner_blocks_weight_bytes(int C, int nblocks, int out_dim);
int gfb_refiner_blocks_chunk(int B, int C, int G);
size_t gfb_refiner_blocks_workspace_bytes(int B, int C, int G, int chunk);
int gfb_refiner_blocks_f16(const float* d, const void* weights, float* out, int B, int C, int G, int nblocks,
                           int out_dim, void* workspace, size_t workspace_bytes, int chunk, int algo, gfb_stream_t stream);
int gfb_flow_update_f32(const float* delta, float* flow, float* certainty, float* disp_pre, int B, int G,
                        int scale, int H0, int W0, int zero_rule, gfb_stream_t stream);
int gfb_upsample_bilinear_f32(const float* in, float* out, int planes, int Hi, int Wi, int Ho, int Wo,
                              gfb_stream_t stream);
/* how many (pre-pass, main) launch pairs one gfb_local_corr_tc2_f32 call issues for these shapes */
int gfb_local_corr_tc2_groups(int B, int C, int Hs, int Ws, int G, int group);
/* F.avg_pool2d(x, 2, 2) on [N,H,W] planes -> [N,H/2,W/2] (local_correlation.py:71). */
int gfb_avg_pool2_f32(const float* x, float* y, int N, int H, int W, gfb_stream_t stream);
/* y[rows, pitch] = x[rows, W] with zero fill of the tail of each row. */
int gfb_pad_rows_f32(const float* x, float* y, long long rows, int W, int pitch, gfb_stream_t stream);
/* ---- profiling aids, NOT part of the drop-in surface (results are wrong with any switch set) ----
 * gfb_debug_local_corr_v2_counters (synchronises): host_out8 = {points on the global-memory path of the point kernels,
 *   lc_tc2 points on the gather path, lc_tc2 gather tiles, 0, and with debug bit 0 of the tc2 debug entry: SM clocks
 *   epilogue warp 0 of every CTA spent waiting / pulling TMEM / emitting rows, chunk count}; reset != 0 zeroes.
 * gfb_debug_local_corr_tc2_f32: gfb_local_corr_tc2_f32 with `debug` switches (bit 0 per-phase clocks, 1 no output stores,
 *   2 no staging stores, 3 no B loads, 4 no MMAs, 5-7 polling flavours / no row arithmetic).
 * gfb_debug_local_corr_pt_f32: gfb_local_corr_pt_f32 with `debug` switches for lc_rot_kernel (bit 0 no window reads /
 *   FMAs, bit 1 no TMA loads). */
int gfb_debug_local_corr_v2_counters(unsigned long long* host_out8, int reset);
int gfb_debug_local_corr_tc2_f32(const float* f0, const float* f1, const float* flow, float* out,
                                 int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                                 int k_total, int k_offset, int group, int debug,
                                 void* workspace, size_t workspace_bytes, gfb_stream_t stream);
int gfb_debug_local_corr_pt_f32(const float* f0, const float* f1, const float* flow, float* out,
                                int B, int C, int Hs, int Ws, int f1_pitch, int G, int r,
                                int k_total, int k_offset, int tune, int debug, gfb_stream_t stream);

/* ---- K2: coarse global match --------------------------------------------------------------
 * replaces GFNet.corr_volume + GFNet.pos_embed, model/network.py:415-440 (call site :251-252).
 *   f0 [B,C,H0,W0], f1 [B,C,H1,W1] fp32 NCHW
 *   flow_out [B,2,H0,W0]: softmax over the N1 = H1*W1 positions of image B of
 *        <f0[:,i], f1[:,j]>/sqrt(C), then expectation of grid[j] = (-1+(2x+1)/W1, -1+(2y+1)/H1)
 *   vol_out  [B,N1,N0] (= the reference's [B,H1,W1,H0,W0]) or NULL to skip materialising it
 *   precision 0 = 3xTF32 split (fp32-faithful), 1 = single TF32 pass; tcgen05 tensor cores.
 *   algo 0 = auto (tcgen05), 1 = SIMT fp32 kernel (any C), 2 = tcgen05 (C % 8 == 0, C <= 128). */
int gfb_global_match_f32(const float* f0, const float* f1, float* flow_out, float* vol_out,
                         int B, int C, int H0, int W0, int H1, int W1,
                         int precision, int algo, gfb_stream_t stream);

/* GFNet.pos_embed on a materialised volume vol [B,N1,N0] (model/network.py:430-440). */
int gfb_pos_embed_f32(const float* vol, float* flow_out, int B, int H0, int W0, int H1, int W1,
                      gfb_stream_t stream);

/* ---- K3: kernel density ---------------------------------------------------------------------
 * replaces utils/kde.py:4-13 `kde(x, std, half, down)`; call site model/network.py:406-408.
 *   x [B,M,D] fp32 (D <= 8; the path uses D = 4), density [B,M]
 *   density[b,m] = sum_{m' = 0, down, 2*down, ...} exp(-||x[b,m]-x[b,m']||^2 / (2 std^2))
 * fp32 throughout (the parity target is the reference's half=False path). */
int gfb_kde_f32(const float* x, float* density, int B, int M, int D, int down, float std,
                gfb_stream_t stream);
/* Same operator for the configuration the reference uses on CUDA (down = 1, i.e. y = x; D = 4; model/network.py:405-408):
 * the kernel is symmetric, so only tile pairs (I, J >= I) are evaluated and every value is added to its row and its
 * column -- half the exponentials.  Sums meet in a 64-bit fixed-point accumulator (deterministic).
 *   cut_sigmas > 0: the points are first sorted along a Hilbert curve (the library's own top-k kernel) and tile pairs
 *   whose bounding boxes are further apart than cut_sigmas * std are skipped; every skipped term is below
 *   exp(-cut_sigmas^2 / 2) (2.3e-11 at 7: < 5e-7 of a density >= 1 for M = 20000).  0 evaluates every pair.
 *   workspace >= gfb_kde_sym_workspace_bytes(B, M) bytes, 256-byte aligned; zeroed by the call. */
size_t gfb_kde_sym_workspace_bytes(int B, int M);
int gfb_kde_sym_f32(const float* x, float* density, int B, int M, float std, float cut_sigmas,
                    void* workspace, size_t workspace_bytes, gfb_stream_t stream);

/* ---- match post-process (tail of GFNet.match, model/network.py:358-384) -----------------------
 *   flow [b,2,G,G], cert_logits [b,1,G,G], attenuation [b,1,G,G] or NULL (the low_res_certainty
 *   term, :334-340, already upsampled).  symmetric != 0: b = 2*Bp, A->B maps first, output
 *   warp [Bp,G,2G,4], cert [Bp,G,2G]; else warp [b,G,G,4], cert [b,G,G]. */
int gfb_match_postprocess_f32(const float* flow, const float* cert_logits, const float* attenuation,
                              float* warp, float* cert, int b, int G, int symmetric,
                              gfb_stream_t stream);

/* ---- balanced sampling pieces (GFNet.sample, model/network.py:385-414) ------------------------
 * torch.multinomial(p, n, replacement=False) == topk(p / q, n), q ~ Exp(1) (ATen); the Exp(1)
 * draw stays with the caller's generator, these kernels do everything around it.
 *   gfb_sample_keys_f32:   key[i] = (cert[i] > thresh ? 1 : cert[i]) / noise[i]      (:391-394,:400)
 *   gfb_balance_keys_f32:  p = 1/(rho+1), p[rho < min_density] = 1e-7; key = p/noise  (:409-411)
 *   gfb_gather_matches_f32: out_m[b,i,:] = warp[b, idx[b,i], :], out_c[b,i] = thresholded cert
 *   gfb_topk_desc_f32: per row of keys [B,n], the k largest in descending order (ties: lower
 *        index first), int64 indices like torch.topk; workspace from gfb_topk_workspace_bytes. */
int gfb_sample_keys_f32(const float* cert, const float* noise, float* key, long long n, float thresh,
                        gfb_stream_t stream);
int gfb_balance_keys_f32(const float* density, const float* noise, float* key, long long n,
                         float min_density, gfb_stream_t stream);
int gfb_gather_matches_f32(const float* warp, const float* cert, const int64_t* idx,
                           float* out_m, float* out_c, int B, long long n_src, int n_sel,
                           float thresh, gfb_stream_t stream);
size_t gfb_topk_workspace_bytes(int B, long long n, int k);
int gfb_topk_desc_f32(const float* keys, int64_t* idx_out, int B, long long n, int k,
                      void* workspace, size_t workspace_bytes, gfb_stream_t stream);

/* ---- K5: homography estimation + metric --------------------------------------------------------
 * replaces estimation.py:60-92: convert_coordinates (:26-45), cv2.findHomography(RANSAC, thr 3,
 * conf 0.99999) (:66-72), the diag(0,0,1) fallback (:73-77) and the 4-corner error (:79-92).
 *   matches [B,N,4] normalised (xA,yA,xB,yB) in [-1,1]; weights [B,N] or NULL (= 1)
 *   pixel coords: pa = ((wq-1)(xA+1)/2, (hq-1)(yA+1)/2), pb likewise with (wsup,hsup);
 *   wq == 0 means `matches` already holds pixel coordinates (pos_a | pos_b), no conversion
 *   n_hyp > 0: RANSAC over n_hyp hash-drawn 4-point models (threshold `thresh` px), then the
 *              normalised DLT (OpenCV runKernel) on the best model's inliers and `gn_iters`
 *              Gauss-Newton steps on the reprojection error (OpenCV's LM refinement).
 *   n_hyp == 0: weighted DLT + refinement on all points (weights as given).
 *   H_out [B,9] float64 row-major, h33 = 1; status[B] 1 = solved, 0 = fallback diag(0,0,1);
 *   n_inliers[B]; mask_out [B,N] uint8 or NULL.  workspace >= gfb_homography_workspace_bytes. */
size_t gfb_homography_workspace_bytes(int B, int N, int n_hyp);
int gfb_homography_f32(const float* matches, const float* weights, int B, int N,
                       float wq, float hq, float wsup, float hsup,
                       int n_hyp, float thresh, int gn_iters, unsigned seed,
                       double* H_out, int* status, int* n_inliers, unsigned char* mask_out,
                       void* workspace, size_t workspace_bytes, gfb_stream_t stream);
/* cv2.findHomography(pos_a, pos_b, cv2.RANSAC, ransacReprojThreshold, maxIters, confidence) restated on the device, batched
 * over pairs -- replaces the host call at estimation.py:66-72 (OpenCV itself is an un-vendored, un-pinned dependency of the
 * reference, requirements.txt:2; the restatement follows modules/calib3d/src/ptsetreg.cpp + fundam.cpp of 4.x and is pinned
 * on cv2 4.13.0 fixtures): RNG((uint64)-1) subset draws, checkSubset (collinearity + orientation), 4-point models scored by
 * the float32 transfer error, adaptive iteration count (RANSACUpdateNumIters), refit on the winner's inliers (normalised
 * DLT) + refinement, and the returned mask = inliers of the REFINED model.  iters_out [B] (or NULL) = iterations OpenCV's
 * loop runs.  matches [B,N,4] normalised (wq == 0: already pixels).  workspace >= gfb_homography_workspace_bytes(B, N, 1). */
int gfb_homography_cv_f32(const float* matches, int B, int N, float wq, float hq, float wsup, float hsup,
                          float thresh, int max_iters, double confidence, int gn_iters,
                          double* H_out, int* status, int* n_inliers, unsigned char* mask_out, int* iters_out,
                          void* workspace, size_t workspace_bytes, gfb_stream_t stream);
/* err[b] = min(clip, mean_k || proj(H_gt[b] c_k) - proj(H_pred[b] c_k) ||), c_k the 4 corners. */
int gfb_corner_error_f64(const double* H_pred, const double* H_gt, float* err, int B,
                         float w, float h, float clip, gfb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GFNET_B200_H */
