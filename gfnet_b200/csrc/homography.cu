// K5 -- homography estimation and corner-error metric (reference: estimation.py:26-45, 60-92).
//
// The reference converts the sampled matches to pixels on the host and calls
// cv2.findHomography(RANSAC, thr 3 px, conf 0.99999).  Here the same three stages of that OpenCV call
// run on the device, batched over image pairs:
//   ransac_kernel : n_hyp hash-drawn 4-point minimal models per pair, each solved exactly in fp64 by
//                   one thread and scored (squared reprojection error <= thr^2) against all N points;
//                   the best (most inliers, lowest hypothesis index on ties) wins through a 64-bit
//                   atomicMax.
//   refit_kernel  : one CTA per pair: inlier mask of the winning model, the normalised DLT of OpenCV's
//                   runKernel (per-axis mean-absolute-deviation normalisation, 9x9 LtL, smallest
//                   eigenvector by cyclic Jacobi), then Gauss-Newton steps on the reprojection error
//                   (the HomographyRefineCallback residual/Jacobian), all accumulations in fp64.
// oracle/estimation.py restates the same algorithm in numpy (same hash, same stage order).
#include "common.cuh"
#include <math.h>

namespace gfb {

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

struct HgParams {
    const float* matches;   // [B,N,4] normalised
    const float* weights;   // [B,N] or null
    int B, N;
    float wq1, hq1, ws1, hs1;  // (w-1), (h-1) of image A and image B
    int pixel_in;
    int n_hyp;
    float thr2;
    int gn_iters;
    uint32_t seed;
    double* H_out;
    int* status;
    int* n_inl;
    unsigned char* mask;
    unsigned long long* best;  // [B] packed (count << 32 | ~hyp)
    double* hyp_H;             // [B, n_hyp, 9]
    int cv_mode;               // 1: OpenCV's own RANSAC loop (cv_ransac_kernel), float32 inlier test, mask from the final H
    int max_iters;             // cv_mode: cv2 maxIters (2000)
    double confidence;         // cv_mode: cv2 confidence
    int* iters_out;            // cv_mode: RANSAC iterations OpenCV's loop ran (or null)
};

// OpenCV HomographyEstimatorCallback::computeError, float32 exactly as compiled there (no contraction)
__device__ __forceinline__ float cv_err_f32(const float (&Hf)[8], float X, float Y, float x, float y) {
    const float den = __fadd_rn(__fadd_rn(__fmul_rn(Hf[6], X), __fmul_rn(Hf[7], Y)), 1.f);
    const float ww = __fdiv_rn(1.f, den);
    const float dx = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[0], X), __fmul_rn(Hf[1], Y)), Hf[2]), ww), x);
    const float dy = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[3], X), __fmul_rn(Hf[4], Y)), Hf[5]), ww), y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

// normalised -> pixel exactly as numpy float32 does it: (w-1) * (x + 1) / 2   (estimation.py:26-45)
__device__ __forceinline__ void to_pixels(const HgParams& p, float4 m, float& ax, float& ay, float& bx, float& by) {
    if (p.pixel_in) { ax = m.x; ay = m.y; bx = m.z; by = m.w; return; }
    ax = p.wq1 * (m.x + 1.f) / 2.f; ay = p.hq1 * (m.y + 1.f) / 2.f;
    bx = p.ws1 * (m.z + 1.f) / 2.f; by = p.hs1 * (m.w + 1.f) / 2.f;
}

// Solve the 8x8 system M h = r in place (partial pivoting).  Returns false when singular.
__device__ bool solve8(double (&M)[8][9]) {
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
        int piv = c;
        double best = fabs(M[c][c]);
        for (int r = c + 1; r < 8; ++r) { double v = fabs(M[r][c]); if (v > best) { best = v; piv = r; } }
        if (!(best > 1e-300)) return false;
        if (piv != c) for (int k = c; k < 9; ++k) { double t = M[c][k]; M[c][k] = M[piv][k]; M[piv][k] = t; }
        const double inv = 1.0 / M[c][c];
        for (int r = c + 1; r < 8; ++r) {
            const double f = M[r][c] * inv;
            for (int k = c; k < 9; ++k) M[r][k] -= f * M[c][k];
        }
    }
#pragma unroll 1
    for (int c = 7; c >= 0; --c) {
        double s = M[c][8];
        for (int k = c + 1; k < 8; ++k) s -= M[c][k] * M[k][8];
        M[c][8] = s / M[c][c];
    }
    return true;
}

constexpr int RS_THREADS = 128;
constexpr int RS_SPLIT = 8;        // threads per hypothesis: each scores every RS_SPLIT-th point (fp64 scoring is latency-bound)

__global__ void __launch_bounds__(RS_THREADS) ransac_kernel(HgParams p) {
    extern __shared__ float4 spts[];   // pixel coords (ax, ay, bx, by) of all N points of this pair
    const int b = blockIdx.y;
    const int hyp = (blockIdx.x * RS_THREADS + threadIdx.x) / RS_SPLIT, sub = threadIdx.x % RS_SPLIT;
    const float4* mb = reinterpret_cast<const float4*>(p.matches) + (size_t)b * p.N;
    for (int i = threadIdx.x; i < p.N; i += RS_THREADS) {
        float4 m = __ldg(mb + i), q;
        to_pixels(p, m, q.x, q.y, q.z, q.w);
        spts[i] = q;
    }
    __syncthreads();
    if (hyp >= p.n_hyp) return;        // n_hyp * RS_SPLIT is a multiple of 32 or the tail warp exits as a whole group
    // four distinct indices from a counter-based hash (oracle.estimation.minimal_sample)
    int idx[4];
    {
        const uint32_t base = mix32(p.seed ^ mix32((uint32_t)b * 0x9E3779B1u + (uint32_t)hyp));
        int got = 0;
        uint32_t ctr = 0;
        while (got < 4) {
            const int v = (int)(mix32(base + ctr * 0x85EBCA6Bu) % (uint32_t)p.N);
            ++ctr;
            bool dup = false;
            for (int e = 0; e < got; ++e) dup |= idx[e] == v;
            if (!dup) idx[got++] = v;
        }
    }
    double M[8][9];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float4 q = spts[idx[e]];
        const double X = q.x, Y = q.y, x = q.z, y = q.w;
        M[2 * e][0] = X; M[2 * e][1] = Y; M[2 * e][2] = 1; M[2 * e][3] = 0; M[2 * e][4] = 0; M[2 * e][5] = 0;
        M[2 * e][6] = -x * X; M[2 * e][7] = -x * Y; M[2 * e][8] = x;
        M[2 * e + 1][0] = 0; M[2 * e + 1][1] = 0; M[2 * e + 1][2] = 0; M[2 * e + 1][3] = X; M[2 * e + 1][4] = Y; M[2 * e + 1][5] = 1;
        M[2 * e + 1][6] = -y * X; M[2 * e + 1][7] = -y * Y; M[2 * e + 1][8] = y;
    }
    bool ok = solve8(M);
    double h[9];
    for (int e = 0; e < 8; ++e) { h[e] = M[e][8]; ok = ok && isfinite(h[e]); }
    h[8] = 1.0;
    double* hs = p.hyp_H + ((size_t)b * p.n_hyp + hyp) * 9;
    if (sub == 0)
        for (int e = 0; e < 9; ++e) hs[e] = ok ? h[e] : 0.0;
    if (!ok) return;                   // the same decision in all RS_SPLIT threads of the hypothesis
    int cnt = 0;
    for (int i = sub; i < p.N; i += RS_SPLIT) {
        // |proj(X) - x|^2 <= thr^2 without the division: multiply through by den^2 (fp64 divisions dominate otherwise)
        const float4 q = spts[i];
        const double X = q.x, Y = q.y;
        const double den = h[6] * X + h[7] * Y + 1.0;
        const double ex = (h[0] * X + h[1] * Y + h[2]) - (double)q.z * den;
        const double ey = (h[3] * X + h[4] * Y + h[5]) - (double)q.w * den;
        cnt += (ex * ex + ey * ey <= (double)p.thr2 * den * den) && den != 0.0;
    }
    // the RS_SPLIT threads of a hypothesis are consecutive lanes of one warp
    const unsigned gmask = ((1u << RS_SPLIT) - 1u) << ((threadIdx.x & 31) / RS_SPLIT * RS_SPLIT);
#pragma unroll
    for (int o = RS_SPLIT / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(gmask, cnt, o);
    if (sub != 0) return;
    const unsigned long long packed = ((unsigned long long)(uint32_t)cnt << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)hyp);
    atomicMax(p.best + b, packed);
}


// =====================================================================================================================
// cv_ransac_kernel: OpenCV's RANSACPointSetRegistrator::run for the homography callback, restated step by step
// (modules/calib3d/src/ptsetreg.cpp, fundam.cpp, 4.x; the reference calls it through cv2.findHomography, estimation.py:66-72)
// =====================================================================================================================
//   RNG rng((uint64)-1): state = (uint32)state * 4164903690 + (state >> 32); uniform(0, n) = next() % n
//   getSubset: 4 distinct indices, redrawn while a duplicate; the attempt is rejected (and 4 new indices drawn) when
//              checkSubset fails: last point collinear with a pair of the others (in either image), or the
//              orientation of the 4 point triples not preserved
//   runKernel on the 4 points (exact 8x8 solve here: same model up to rounding), computeError in float32, inliers err <= thr^2
//   best = most inliers so far (strictly more than before and > 3); niters = RANSACUpdateNumIters(confidence, outlier ratio)
// The loop is sequential in OpenCV; here a CTA per pair runs it in rounds of CV_ATT attempts: thread 0 advances the RNG
// (the draws do not depend on the data), the subset checks, minimal solves and inlier counts of a round run in parallel,
// then thread 0 replays the round in order with OpenCV's update rule and stops where OpenCV stops.
constexpr int CV_ATT = 64, CV_THREADS = 256;

__device__ __forceinline__ bool cv_collinear_last(const float4 (&q)[4], bool second) {
    // haveCollinearPoints(m, 4): is point 3 on a line through two of points 0..2 (or too close)
    const float xi = second ? q[3].z : q[3].x, yi = second ? q[3].w : q[3].y;
    for (int j = 0; j < 3; ++j) {
        const double dx1 = (double)(second ? q[j].z : q[j].x) - xi, dy1 = (double)(second ? q[j].w : q[j].y) - yi;
        for (int k = 0; k < j; ++k) {
            const double dx2 = (double)(second ? q[k].z : q[k].x) - xi, dy2 = (double)(second ? q[k].w : q[k].y) - yi;
            if (fabs(dx2 * dy1 - dy2 * dx1) <= 1.1920928955078125e-07 * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
        }
    }
    return false;
}
__device__ __forceinline__ double cv_det3(double a0, double a1, double b0, double b1, double c0, double c1) {
    // determinant of [[a0,a1,1],[b0,b1,1],[c0,c1,1]] (cv::determinant of a Matx33d)
    return a0 * (b1 - c1) - a1 * (b0 - c0) + (b0 * c1 - b1 * c0);
}
__device__ bool cv_check_subset(const float4 (&q)[4]) {
    if (cv_collinear_last(q, false) || cv_collinear_last(q, true)) return false;
    const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
    int negative = 0;
    for (int i = 0; i < 4; ++i) {
        const int* t = tt[i];
        const double dA = cv_det3(q[t[0]].x, q[t[0]].y, q[t[1]].x, q[t[1]].y, q[t[2]].x, q[t[2]].y);
        const double dB = cv_det3(q[t[0]].z, q[t[0]].w, q[t[1]].z, q[t[1]].w, q[t[2]].z, q[t[2]].w);
        negative += dA * dB < 0;
    }
    return negative == 0 || negative == 4;
}
__device__ int cv_update_num_iters(double p, double ep, int max_iters) {   // RANSACUpdateNumIters, modelPoints = 4
    p = fmin(fmax(p, 0.), 1.); ep = fmin(fmax(ep, 0.), 1.);
    double num = fmax(1. - p, 2.2250738585072014e-308);
    double denom = 1. - pow(1. - ep, 4.0);
    if (denom < 2.2250738585072014e-308) return 0;
    num = log(num); denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

__global__ void __launch_bounds__(CV_THREADS) cv_ransac_kernel(HgParams p) {
    extern __shared__ float4 spts[];                 // pixel coords (ax, ay, bx, by) of all N points of this pair
    __shared__ int s_idx[CV_ATT][4];
    __shared__ int s_valid[CV_ATT], s_good[CV_ATT];
    __shared__ float s_Hf[CV_ATT][8];
    __shared__ double s_H[CV_ATT][9];
    __shared__ unsigned long long s_rng;
    __shared__ int s_iter, s_niters, s_maxgood, s_done, s_best_set;
    __shared__ double s_bestH[9];
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float4* mb = reinterpret_cast<const float4*>(p.matches) + (size_t)b * p.N;
    for (int i = tid; i < p.N; i += CV_THREADS) {
        float4 m = __ldg(mb + i), q;
        to_pixels(p, m, q.x, q.y, q.z, q.w);
        spts[i] = q;
    }
    if (tid == 0) { s_rng = 0xFFFFFFFFFFFFFFFFull; s_iter = 0; s_niters = p.max_iters; s_maxgood = 0; s_done = 0; s_best_set = 0; }
    __syncthreads();
    const float thr = p.thr2;
    for (int round = 0; round < 100000; ++round) {
        if (tid == 0) {                                // the RNG stream: 4 distinct indices per attempt
            unsigned long long st = s_rng;
            const unsigned n = (unsigned)p.N;
            for (int a = 0; a < CV_ATT; ++a) {
                int id[4];
                for (int i = 0; i < 4; ++i) {
                    bool dup;
                    int v;
                    do {
                        st = (unsigned long long)(unsigned)st * 4164903690ull + (st >> 32);
                        v = (int)((unsigned)st % n);
                        dup = false;
                        for (int e = 0; e < i; ++e) dup |= id[e] == v;
                    } while (dup);
                    id[i] = v;
                }
                for (int i = 0; i < 4; ++i) s_idx[a][i] = id[i];
            }
            s_rng = st;
        }
        __syncthreads();
        if (tid < CV_ATT) {                            // subset check + minimal model, one attempt per thread
            float4 q[4];
            for (int e = 0; e < 4; ++e) q[e] = spts[s_idx[tid][e]];
            int valid = cv_check_subset(q) ? 1 : 0;    // 0: rejected inside getSubset (not an iteration)
            if (valid) {
                double M[8][9];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double X = q[e].x, Y = q[e].y, x = q[e].z, y = q[e].w;
                    M[2 * e][0] = X; M[2 * e][1] = Y; M[2 * e][2] = 1; M[2 * e][3] = 0; M[2 * e][4] = 0; M[2 * e][5] = 0;
                    M[2 * e][6] = -x * X; M[2 * e][7] = -x * Y; M[2 * e][8] = x;
                    M[2 * e + 1][0] = 0; M[2 * e + 1][1] = 0; M[2 * e + 1][2] = 0; M[2 * e + 1][3] = X; M[2 * e + 1][4] = Y; M[2 * e + 1][5] = 1;
                    M[2 * e + 1][6] = -y * X; M[2 * e + 1][7] = -y * Y; M[2 * e + 1][8] = y;
                }
                bool ok = solve8(M);
                for (int e = 0; e < 8; ++e) ok = ok && isfinite(M[e][8]);
                if (ok) {
                    for (int e = 0; e < 8; ++e) { s_H[tid][e] = M[e][8]; s_Hf[tid][e] = (float)M[e][8]; }
                    s_H[tid][8] = 1.0;
                } else {
                    valid = 2;                         // an iteration without a model (runKernel returned 0 models)
                }
            }
            s_valid[tid] = valid;
            s_good[tid] = 0;
        }
        __syncthreads();
        for (int a = warp; a < CV_ATT; a += CV_THREADS / 32) {      // inlier counts, one attempt per warp at a time
            if (s_valid[a] != 1) continue;
            float Hf[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) Hf[e] = s_Hf[a][e];
            int cnt = 0;
            for (int i = lane; i < p.N; i += 32) {
                const float4 q = spts[i];
                cnt += cv_err_f32(Hf, q.x, q.y, q.z, q.w) <= thr;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (lane == 0) s_good[a] = cnt;
        }
        __syncthreads();
        if (tid == 0) {                                // OpenCV's loop over the round, in order
            int iter = s_iter, niters = s_niters, maxgood = s_maxgood;
            for (int a = 0; a < CV_ATT && iter < niters; ++a) {
                if (s_valid[a] == 0) continue;
                if (s_valid[a] == 1 && s_good[a] > max(maxgood, 3)) {
                    maxgood = s_good[a];
                    for (int e = 0; e < 9; ++e) s_bestH[e] = s_H[a][e];
                    s_best_set = 1;
                    niters = cv_update_num_iters(p.confidence, (double)(p.N - maxgood) / p.N, niters);
                }
                ++iter;
            }
            s_iter = iter; s_niters = niters; s_maxgood = maxgood;
            s_done = iter >= niters;
        }
        __syncthreads();
        if (s_done) break;
    }
    if (tid < 9) p.hyp_H[(size_t)b * 9 + tid] = s_best_set ? s_bestH[tid] : 0.0;
    if (tid == 0) {
        p.best[b] = ((unsigned long long)(unsigned)(s_best_set ? s_maxgood : 0) << 32) | 0xFFFFFFFFull;     // hypothesis 0
        if (p.iters_out) p.iters_out[b] = s_iter;
    }
}

// ---- block reduction of NACC doubles: result in sm_out[0..NACC) ---------------------------------
constexpr int RF_THREADS = 512;
template <int NACC>
__device__ void block_reduce(double (&acc)[NACC], double* sm_part /*[RF_THREADS/32][NACC]*/, double* sm_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < NACC; ++e) {
        double v = warp_sum(acc[e]);
        if (lane == 0) sm_part[warp * NACC + e] = v;
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        double s = 0;
        for (int w = 0; w < RF_THREADS / 32; ++w) s += sm_part[w * NACC + threadIdx.x];
        sm_out[threadIdx.x] = s;
    }
    __syncthreads();
}

// cyclic Jacobi on a symmetric 9x9 (shared memory, one thread); returns eigenvector of the smallest eigenvalue
__device__ void jacobi9_smallest(double* A /*81*/, double* V /*81*/, double* out /*9*/) {
    for (int i = 0; i < 81; ++i) V[i] = 0.0;
    for (int i = 0; i < 9; ++i) V[i * 9 + i] = 1.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0, diag = 0;
        for (int i = 0; i < 9; ++i) { diag += A[i * 9 + i] * A[i * 9 + i]; for (int j = i + 1; j < 9; ++j) off += A[i * 9 + j] * A[i * 9 + j]; }
        if (off <= 1e-30 * diag) break;
        for (int pi = 0; pi < 8; ++pi) {
            for (int q = pi + 1; q < 9; ++q) {
                const double apq = A[pi * 9 + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * 9 + q] - A[pi * 9 + pi]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 9; ++k) {
                    const double akp = A[k * 9 + pi], akq = A[k * 9 + q];
                    A[k * 9 + pi] = c * akp - s * akq;
                    A[k * 9 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 9; ++k) {
                    const double apk = A[pi * 9 + k], aqk = A[q * 9 + k];
                    A[pi * 9 + k] = c * apk - s * aqk;
                    A[q * 9 + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 9; ++k) {
                    const double vkp = V[k * 9 + pi], vkq = V[k * 9 + q];
                    V[k * 9 + pi] = c * vkp - s * vkq;
                    V[k * 9 + q] = s * vkp + c * vkq;
                }
            }
        }
    }
    int best = 0;
    for (int i = 1; i < 9; ++i) if (A[i * 9 + i] < A[best * 9 + best]) best = i;
    for (int k = 0; k < 9; ++k) out[k] = V[k * 9 + best];
}

// Smallest eigenvector of a symmetric positive semi-definite 9x9 by shifted inverse iteration (Cholesky of A + mu I, six
// solves): LtL of a homography fit has one eigenvalue far below the rest, so the iteration converges in 2-3 steps; a
// single thread needs ~1.5k flops instead of the ~60k of a cyclic Jacobi sweep series.  false -> caller falls back to Jacobi.
__device__ bool inv_iter9_smallest(const double* A /*81*/, double* L /*81*/, double* out /*9*/) {
    double tr = 0;
    for (int i = 0; i < 9; ++i) tr += A[i * 9 + i];
    if (!(tr > 0.0) || !isfinite(tr)) return false;
    const double mu = tr * 1e-15;
    for (int j = 0; j < 9; ++j) {
        double d = A[j * 9 + j] + mu;
        for (int k = 0; k < j; ++k) d -= L[j * 9 + k] * L[j * 9 + k];
        if (!(d > tr * 1e-18)) return false;
        d = sqrt(d);
        L[j * 9 + j] = d;
        for (int i = j + 1; i < 9; ++i) {
            double v = A[i * 9 + j];
            for (int k = 0; k < j; ++k) v -= L[i * 9 + k] * L[j * 9 + k];
            L[i * 9 + j] = v / d;
        }
    }
    double x[9], y[9];
    for (int i = 0; i < 9; ++i) x[i] = 1.0 / 3.0 + 0.01 * i;
    for (int it = 0; it < 6; ++it) {
        for (int i = 0; i < 9; ++i) { double v = x[i]; for (int k = 0; k < i; ++k) v -= L[i * 9 + k] * y[k]; y[i] = v / L[i * 9 + i]; }
        for (int i = 8; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < 9; ++k) v -= L[k * 9 + i] * x[k]; x[i] = v / L[i * 9 + i]; }
        double nrm = 0;
        for (int i = 0; i < 9; ++i) nrm += x[i] * x[i];
        nrm = sqrt(nrm);
        if (!(nrm > 0.0) || !isfinite(nrm)) return false;
        for (int i = 0; i < 9; ++i) x[i] /= nrm;
    }
    for (int i = 0; i < 9; ++i) out[i] = x[i];
    return true;
}

__global__ void __launch_bounds__(RF_THREADS) refit_kernel(HgParams p) {
    extern __shared__ float s_w[];           // [N] weight of every point (0 = not an inlier), filled by the first pass
    __shared__ double sm_part[(RF_THREADS / 32) * 30];
    __shared__ double sm_red[32];
    __shared__ double sA[81], sV[81], sh[9], sH[9];
    __shared__ int s_ok, s_cnt;
    const int b = blockIdx.x;
    const float4* mb = reinterpret_cast<const float4*>(p.matches) + (size_t)b * p.N;
    const float* wb = p.weights ? p.weights + (size_t)b * p.N : nullptr;
    unsigned char* mk = p.mask ? p.mask + (size_t)b * p.N : nullptr;

    auto fail = [&]() {   // estimation.py:73-77: H = diag(0,0,1)
        if (threadIdx.x < 9) p.H_out[(size_t)b * 9 + threadIdx.x] = threadIdx.x == 8 ? 1.0 : 0.0;
        if (threadIdx.x == 0) { p.status[b] = 0; p.n_inl[b] = 0; }
        if (mk) for (int i = threadIdx.x; i < p.N; i += RF_THREADS) mk[i] = 0;
    };

    // ---- stage 1: which points count (RANSAC winner's inliers, or every point with weight > 0)
    double hb[9];
    bool use_model = p.n_hyp > 0;
    if (use_model) {
        const unsigned long long packed = p.best[b];
        const int cnt = (int)(packed >> 32);
        const int hyp = (int)(0xFFFFFFFFu - (uint32_t)(packed & 0xFFFFFFFFull));
        if (cnt < 4) { fail(); return; }
        const double* hs = p.hyp_H + ((size_t)b * p.n_hyp + hyp) * 9;
        for (int e = 0; e < 9; ++e) hb[e] = hs[e];
    }
    float hbf[8];
    for (int e = 0; e < 8; ++e) hbf[e] = use_model ? (float)hb[e] : 0.f;
    auto weight_of = [&](int i, const float4& q) -> double {
        double w = wb ? (double)wb[i] : 1.0;
        if (use_model && p.cv_mode) {
            if (!(cv_err_f32(hbf, q.x, q.y, q.z, q.w) <= p.thr2)) w = 0.0;
        } else if (use_model) {
            const double X = q.x, Y = q.y;
            const double ww = 1.0 / (hb[6] * X + hb[7] * Y + 1.0);
            const double dx = (hb[0] * X + hb[1] * Y + hb[2]) * ww - (double)q.z;
            const double dy = (hb[3] * X + hb[4] * Y + hb[5]) * ww - (double)q.w;
            if (!(dx * dx + dy * dy <= (double)p.thr2)) w = 0.0;
        }
        return w > 0.0 ? w : 0.0;
    };

    // ---- stage 2: weighted centroids, then mean absolute deviations (OpenCV runKernel normalisation)
    double a6[6] = {0, 0, 0, 0, 0, 0};   // sw, cnt, sum ax, ay, bx, by
    for (int i = threadIdx.x; i < p.N; i += RF_THREADS) {
        float4 q; to_pixels(p, __ldg(mb + i), q.x, q.y, q.z, q.w);
        const double w = weight_of(i, q);
        s_w[i] = (float)w;                   // weights are floats (or 1): the round trip is exact
        if (mk) mk[i] = w > 0.0;
        a6[0] += w; a6[1] += (w > 0.0); a6[2] += w * q.x; a6[3] += w * q.y; a6[4] += w * q.z; a6[5] += w * q.w;
    }
    block_reduce<6>(a6, sm_part, sm_red);
    const double sw = sm_red[0];
    const int cnt = (int)(sm_red[1] + 0.5);
    const double cAx = sm_red[2] / sw, cAy = sm_red[3] / sw, cBx = sm_red[4] / sw, cBy = sm_red[5] / sw;
    __syncthreads();
    if (cnt < 4 || !(sw > 0.0)) { fail(); return; }
    double d4[4] = {0, 0, 0, 0};
    for (int i = threadIdx.x; i < p.N; i += RF_THREADS) {
        float4 q; to_pixels(p, __ldg(mb + i), q.x, q.y, q.z, q.w);
        const double w = (double)s_w[i];
        d4[0] += w * fabs(q.x - cAx); d4[1] += w * fabs(q.y - cAy); d4[2] += w * fabs(q.z - cBx); d4[3] += w * fabs(q.w - cBy);
    }
    block_reduce<4>(d4, sm_part, sm_red);
    const double sAx = sw / sm_red[0], sAy = sw / sm_red[1], sBx = sw / sm_red[2], sBy = sw / sm_red[3];
    __syncthreads();
    if (!(isfinite(sAx) && isfinite(sAy) && isfinite(sBx) && isfinite(sBy))) { fail(); return; }

    // ---- stage 3: LtL = sum w (Lx Lx^T + Ly Ly^T) in its block form: S0, Sx, Sy, Sq (6 unique each)
    double acc[24];
#pragma unroll
    for (int e = 0; e < 24; ++e) acc[e] = 0;
    for (int i = threadIdx.x; i < p.N; i += RF_THREADS) {
        float4 q; to_pixels(p, __ldg(mb + i), q.x, q.y, q.z, q.w);
        const double w = (double)s_w[i];
        if (w == 0.0) continue;
        const double X = (q.x - cAx) * sAx, Y = (q.y - cAy) * sAy, x = (q.z - cBx) * sBx, y = (q.w - cBy) * sBy;
        const double u[6] = {X * X, X * Y, X, Y * Y, Y, 1.0};   // upper triangle of u u^T, u = (X, Y, 1)
        const double wx = w * x, wy = w * y, wq = w * (x * x + y * y);
#pragma unroll
        for (int e = 0; e < 6; ++e) { acc[e] += w * u[e]; acc[6 + e] += wx * u[e]; acc[12 + e] += wy * u[e]; acc[18 + e] += wq * u[e]; }
    }
    block_reduce<24>(acc, sm_part, sm_part + (RF_THREADS / 32) * 24);
    if (threadIdx.x == 0) {
        const double* r = sm_part + (RF_THREADS / 32) * 24;
        const int tri[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const double s0 = r[tri[i][j]], sx = r[6 + tri[i][j]], sy = r[12 + tri[i][j]], sq = r[18 + tri[i][j]];
                sA[i * 9 + j] = s0;            sA[i * 9 + 3 + j] = 0.0;           sA[i * 9 + 6 + j] = -sx;
                sA[(3 + i) * 9 + j] = 0.0;     sA[(3 + i) * 9 + 3 + j] = s0;      sA[(3 + i) * 9 + 6 + j] = -sy;
                sA[(6 + i) * 9 + j] = -sx;     sA[(6 + i) * 9 + 3 + j] = -sy;     sA[(6 + i) * 9 + 6 + j] = sq;
            }
        if (!inv_iter9_smallest(sA, sV, sh)) jacobi9_smallest(sA, sV, sh);
        // H = invNormB * H0 * NormA, then / h33
        const double iB[9] = {1.0 / sBx, 0, cBx, 0, 1.0 / sBy, cBy, 0, 0, 1};
        const double nA[9] = {sAx, 0, -cAx * sAx, 0, sAy, -cAy * sAy, 0, 0, 1};
        double t[9], H[9];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += iB[i * 3 + k] * sh[k * 3 + j]; t[i * 3 + j] = s; }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += t[i * 3 + k] * nA[k * 3 + j]; H[i * 3 + j] = s; }
        bool ok = fabs(H[8]) > 1e-300;
        for (int e = 0; e < 9; ++e) { sH[e] = H[e] / H[8]; ok = ok && isfinite(sH[e]); }
        s_ok = ok; s_cnt = cnt;
    }
    __syncthreads();
    if (!s_ok) { fail(); return; }

    // ---- stage 4: Gauss-Newton on the reprojection error (OpenCV HomographyRefineCallback)
    double h[8];
    for (int e = 0; e < 8; ++e) h[e] = sH[e];
    double cost = -1.0;
    double gsys[8][9];   // thread 0 only
    for (int it = 0; it <= p.gn_iters; ++it) {
        double g[30];
#pragma unroll
        for (int e = 0; e < 30; ++e) g[e] = 0;
        for (int i = threadIdx.x; i < p.N; i += RF_THREADS) {
            float4 q; to_pixels(p, __ldg(mb + i), q.x, q.y, q.z, q.w);
            const double w = (double)s_w[i];
            if (w == 0.0) continue;
            const double X = q.x, Y = q.y;
            const double ww = 1.0 / (h[6] * X + h[7] * Y + 1.0);
            const double xi = (h[0] * X + h[1] * Y + h[2]) * ww, yi = (h[3] * X + h[4] * Y + h[5]) * ww;
            const double rx = xi - (double)q.z, ry = yi - (double)q.w;
            const double v0 = X * ww, v1 = Y * ww, v2 = ww;
            // V0 = sum w v v^T (00 01 02 11 12 22)
            g[0] += w * v0 * v0; g[1] += w * v0 * v1; g[2] += w * v0 * v2; g[3] += w * v1 * v1; g[4] += w * v1 * v2; g[5] += w * v2 * v2;
            const double wxi = w * xi, wyi = w * yi;
            // Vx = sum w xi v (v0,v1)^T  (3x2), Vy likewise
            g[6] += wxi * v0 * v0; g[7] += wxi * v0 * v1; g[8] += wxi * v1 * v0; g[9] += wxi * v1 * v1; g[10] += wxi * v2 * v0; g[11] += wxi * v2 * v1;
            g[12] += wyi * v0 * v0; g[13] += wyi * v0 * v1; g[14] += wyi * v1 * v0; g[15] += wyi * v1 * v1; g[16] += wyi * v2 * v0; g[17] += wyi * v2 * v1;
            const double wq = w * (xi * xi + yi * yi);
            g[18] += wq * v0 * v0; g[19] += wq * v0 * v1; g[20] += wq * v1 * v1;
            g[21] += w * rx * v0; g[22] += w * rx * v1; g[23] += w * rx * v2;
            g[24] += w * ry * v0; g[25] += w * ry * v1; g[26] += w * ry * v2;
            const double rr = w * (rx * xi + ry * yi);
            g[27] -= rr * v0; g[28] -= rr * v1;
            g[29] += w * (rx * rx + ry * ry);
        }
        block_reduce<30>(g, sm_part, sm_red);
        __shared__ int s_state;     // 0 = continue with new step in sh[0..8), 1 = stop
        if (threadIdx.x == 0) {
            const double* r = sm_red;
            const double new_cost = r[29];
            int state = 1;
            if (it > 0 && !(new_cost < cost)) {
                // candidate rejected: keep the previous estimate (already in sH)
                state = 1;
            } else {
                for (int e = 0; e < 8; ++e) sH[e] = h[e];   // accept current point
                if (it < p.gn_iters) {
                    const int t3[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
                    for (int i = 0; i < 8; ++i) for (int j = 0; j < 9; ++j) gsys[i][j] = 0;
                    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { gsys[i][j] = r[t3[i][j]]; gsys[3 + i][3 + j] = r[t3[i][j]]; }
                    for (int i = 0; i < 3; ++i) for (int j = 0; j < 2; ++j) {
                        gsys[i][6 + j] = -r[6 + i * 2 + j]; gsys[6 + j][i] = -r[6 + i * 2 + j];
                        gsys[3 + i][6 + j] = -r[12 + i * 2 + j]; gsys[6 + j][3 + i] = -r[12 + i * 2 + j];
                    }
                    gsys[6][6] = r[18]; gsys[6][7] = r[19]; gsys[7][6] = r[19]; gsys[7][7] = r[20];
                    for (int i = 0; i < 8; ++i) gsys[i][8] = r[21 + i];
                    if (solve8(gsys)) {
                        double mx = 0;
                        for (int e = 0; e < 8; ++e) { sh[e] = h[e] - gsys[e][8]; mx = fmax(mx, fabs(gsys[e][8])); }
                        bool fin = true;
                        for (int e = 0; e < 8; ++e) fin = fin && isfinite(sh[e]);
                        state = (fin && mx >= 1e-12) ? 0 : 1;
                    }
                }
            }
            s_state = state;
        }
        __syncthreads();
        if (s_state) break;
        if (it == 0 || true) cost = sm_red[29];
        for (int e = 0; e < 8; ++e) h[e] = sh[e];
        __syncthreads();
    }
    if (threadIdx.x < 9) p.H_out[(size_t)b * 9 + threadIdx.x] = threadIdx.x == 8 ? 1.0 : sH[threadIdx.x];
    if (p.cv_mode) {
        // cv2.findHomography returns the mask of the REFINED model (float32 computeError <= thr^2), checked against 4.13
        float hf[8];
        for (int e = 0; e < 8; ++e) hf[e] = (float)sH[e];
        double c1[1] = {0};
        for (int i = threadIdx.x; i < p.N; i += RF_THREADS) {
            float4 q; to_pixels(p, __ldg(mb + i), q.x, q.y, q.z, q.w);
            const bool in = cv_err_f32(hf, q.x, q.y, q.z, q.w) <= p.thr2;
            if (mk) mk[i] = in;
            c1[0] += in;
        }
        __syncthreads();
        block_reduce<1>(c1, sm_part, sm_red);
        if (threadIdx.x == 0) { p.status[b] = 1; p.n_inl[b] = (int)(sm_red[0] + 0.5); }
        return;
    }
    if (threadIdx.x == 0) { p.status[b] = 1; p.n_inl[b] = s_cnt; }
}

__global__ void init_best_kernel(unsigned long long* best, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) best[i] = 0ull;
}

// reference: estimation.py:79-92
__global__ void corner_error_kernel(const double* __restrict__ Hp, const double* __restrict__ Hg, float* __restrict__ err,
                                    int B, double w1, double h1, double clip) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double cx[4] = {0, 0, w1, w1}, cy[4] = {0, h1, 0, h1};
    const double* P = Hp + (size_t)b * 9;
    const double* Gt = Hg + (size_t)b * 9;
    double s = 0;
    for (int k = 0; k < 4; ++k) {
        const double pw = P[6] * cx[k] + P[7] * cy[k] + P[8], gw = Gt[6] * cx[k] + Gt[7] * cy[k] + Gt[8];
        const double px = (P[0] * cx[k] + P[1] * cy[k] + P[2]) / pw, py = (P[3] * cx[k] + P[4] * cy[k] + P[5]) / pw;
        const double gx = (Gt[0] * cx[k] + Gt[1] * cy[k] + Gt[2]) / gw, gy = (Gt[3] * cx[k] + Gt[4] * cy[k] + Gt[5]) / gw;
        s += sqrt((px - gx) * (px - gx) + (py - gy) * (py - gy));
    }
    s *= 0.25;
    if (s > clip) s = clip;
    err[b] = (float)s;
}

}  // namespace gfb

using namespace gfb;

extern "C" size_t gfb_homography_workspace_bytes(int B, int N, int n_hyp) {
    (void)N;
    if (B <= 0) return 0;
    return (size_t)B * sizeof(unsigned long long) + (size_t)B * (size_t)(n_hyp > 0 ? n_hyp : 0) * 9 * sizeof(double) + 16;
}

extern "C" int gfb_homography_f32(const float* matches, const float* weights, int B, int N,
                                  float wq, float hq, float wsup, float hsup,
                                  int n_hyp, float thresh, int gn_iters, unsigned seed,
                                  double* H_out, int* status, int* n_inliers, unsigned char* mask_out,
                                  void* workspace, size_t workspace_bytes, gfb_stream_t stream) {
    GFB_CHECK_ARG(matches && H_out && status && n_inliers && B > 0 && N >= 4 && n_hyp >= 0 && gn_iters >= 0 && gn_iters <= 100);
    GFB_CHECK_ARG(B <= 65535 && thresh > 0.f);
    if (!gfb_aligned(matches, 16)) return GFB_EALIGN;
    if (!workspace || workspace_bytes < gfb_homography_workspace_bytes(B, N, n_hyp) || !gfb_aligned(workspace, 8)) return GFB_EWORKSPACE;
    const size_t smem = (size_t)N * sizeof(float4);
    if (n_hyp > 0 && smem > 200 * 1024) return GFB_EUNSUPPORTED;
    HgParams p;
    p.matches = matches; p.weights = weights; p.B = B; p.N = N;
    p.pixel_in = (wq == 0.f);
    p.wq1 = wq - 1.f; p.hq1 = hq - 1.f; p.ws1 = wsup - 1.f; p.hs1 = hsup - 1.f;
    p.n_hyp = n_hyp; p.thr2 = thresh * thresh; p.gn_iters = gn_iters; p.seed = seed;
    p.cv_mode = 0; p.max_iters = 0; p.confidence = 0.0; p.iters_out = nullptr;
    p.H_out = H_out; p.status = status; p.n_inl = n_inliers; p.mask = mask_out;
    p.best = reinterpret_cast<unsigned long long*>(workspace);
    p.hyp_H = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + (((size_t)B * 8 + 15) / 16) * 16);
    cudaStream_t st = gfb_cu(stream);
    if (n_hyp > 0) {
        init_best_kernel<<<(B + 127) / 128, 128, 0, st>>>(p.best, B);
        cudaError_t e = cudaFuncSetAttribute(ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        dim3 grid((n_hyp * RS_SPLIT + RS_THREADS - 1) / RS_THREADS, B);
        ransac_kernel<<<grid, RS_THREADS, smem, st>>>(p);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    const size_t rsmem = (size_t)N * sizeof(float);
    if (rsmem > 200 * 1024) return GFB_EUNSUPPORTED;
    cudaError_t e2 = cudaFuncSetAttribute(refit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
    if (e2 != cudaSuccess) return (int)e2;
    refit_kernel<<<B, RF_THREADS, rsmem, st>>>(p);
    GFB_LAUNCH_RESULT();
}

// cv2.findHomography(pos_a, pos_b, cv2.RANSAC, thresh, maxIters, confidence) restated on the device (estimation.py:66-72):
// OpenCV's RNG, subset checks, adaptive iteration count, refit on the inliers, refinement, mask of the refined model.
extern "C" int gfb_homography_cv_f32(const float* matches, int B, int N, float wq, float hq, float wsup, float hsup,
                                     float thresh, int max_iters, double confidence, int gn_iters,
                                     double* H_out, int* status, int* n_inliers, unsigned char* mask_out, int* iters_out,
                                     void* workspace, size_t workspace_bytes, gfb_stream_t stream) {
    GFB_CHECK_ARG(matches && H_out && status && n_inliers && B > 0 && N >= 5 && max_iters > 0 && gn_iters >= 0 && gn_iters <= 100);
    GFB_CHECK_ARG(B <= 65535 && thresh > 0.f && confidence > 0.0 && confidence < 1.0);
    if (!gfb_aligned(matches, 16)) return GFB_EALIGN;
    if (!workspace || workspace_bytes < gfb_homography_workspace_bytes(B, N, 1) || !gfb_aligned(workspace, 8)) return GFB_EWORKSPACE;
    const size_t smem = (size_t)N * sizeof(float4);
    if (smem > 180 * 1024) return GFB_EUNSUPPORTED;
    HgParams p;
    p.matches = matches; p.weights = nullptr; p.B = B; p.N = N;
    p.pixel_in = (wq == 0.f);
    p.wq1 = wq - 1.f; p.hq1 = hq - 1.f; p.ws1 = wsup - 1.f; p.hs1 = hsup - 1.f;
    p.n_hyp = 1; p.thr2 = thresh * thresh; p.gn_iters = gn_iters; p.seed = 0;
    p.H_out = H_out; p.status = status; p.n_inl = n_inliers; p.mask = mask_out;
    p.best = reinterpret_cast<unsigned long long*>(workspace);
    p.hyp_H = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + (((size_t)B * 8 + 15) / 16) * 16);
    p.cv_mode = 1; p.max_iters = max_iters; p.confidence = confidence; p.iters_out = iters_out;
    cudaStream_t st = gfb_cu(stream);
    cudaError_t e = cudaFuncSetAttribute(cv_ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cv_ransac_kernel<<<B, CV_THREADS, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const size_t rsmem = (size_t)N * sizeof(float);
    e = cudaFuncSetAttribute(refit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem);
    if (e != cudaSuccess) return (int)e;
    refit_kernel<<<B, RF_THREADS, rsmem, st>>>(p);
    GFB_LAUNCH_RESULT();
}

extern "C" int gfb_corner_error_f64(const double* H_pred, const double* H_gt, float* err, int B,
                                    float w, float h, float clip, gfb_stream_t stream) {
    GFB_CHECK_ARG(H_pred && H_gt && err && B > 0);
    corner_error_kernel<<<(B + 127) / 128, 128, 0, gfb_cu(stream)>>>(H_pred, H_gt, err, B, (double)w - 1.0, (double)h - 1.0, (double)clip);
    GFB_LAUNCH_RESULT();
}
