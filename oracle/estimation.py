"""Oracle for the homography estimation tail (reference: estimation.py:12-45, 60-92).

TEST INFRASTRUCTURE ONLY.

The reference delegates the solve to ``cv2.findHomography(..., cv2.RANSAC, confidence=0.99999,
ransacReprojThreshold=3)`` (estimation.py:66-72).  OpenCV is a binary dependency that is neither
vendored under /root/reference nor version-pinned (requirements.txt:2: ``opencv-python``), and the
reference holds no test or golden vector for it: PARITY WITH THE REFERENCE'S RANSAC DRAW IS
UNPINNED.  What is restated here is OpenCV's published algorithm for that call (modules/calib3d/
src/fundam.cpp, 4.x):

* ``HomographyEstimatorCallback::runKernel`` -- per-axis mean-absolute-deviation normalisation,
  9x9 ``LtL`` accumulation, smallest eigenvector, de-normalise, divide by h33   -> ``weighted_dlt``
* ``HomographyRefineCallback`` + ``LMSolver(maxIters=10)`` -- minimise the reprojection error in
  the second image over the 8 free parameters                                   -> ``refine_homography_lm``
* RANSAC: 4-point minimal models scored by squared reprojection error <= thr^2, the best model's
  inlier set is re-fitted with runKernel and refined (findHomography, "if result && npoints > 4").

``find_homography_cv2`` calls the installed cv2 (4.13.0 in this image) with the reference's exact
arguments and is the anchor the restatement is checked against (tests/test_oracle_estimation.py).
"""
import numpy as np

FALLBACK_H = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.0], [0.0, 0.0, 1.0]])


def auc(errors, thresholds):
    """Area under the recall-vs-error curve, normalised by each threshold.

    reference: estimation.py:12-24 (sort, prepend 0, searchsorted, trapezoid / t).
    """
    e = np.sort(np.asarray(errors, dtype=np.float64))
    n = len(e)
    e = np.concatenate(([0.0], e))
    rec = np.concatenate(([0.0], (np.arange(n) + 1) / n))
    out = []
    for t in thresholds:
        last = int(np.searchsorted(e, t))
        x = np.concatenate((e[:last], [t]))
        y = np.concatenate((rec[:last], [rec[last - 1]]))
        out.append(float(np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) * 0.5) / t))
    return out


def convert_coordinates(im_A_coords, im_A_to_im_B, wq, hq, wsup, hsup):
    """Normalised [-1,1] -> pixel: ``px = (w-1)(x+1)/2``; reference: estimation.py:26-45."""
    a = np.stack(((wq - 1) * (im_A_coords[..., 0] + 1) / 2, (hq - 1) * (im_A_coords[..., 1] + 1) / 2), axis=-1)
    b = np.stack(((wsup - 1) * (im_A_to_im_B[..., 0] + 1) / 2, (hsup - 1) * (im_A_to_im_B[..., 1] + 1) / 2), axis=-1)
    return a, b


def fallback_homography():
    """``H = diag(0,0,1)`` when the solver fails; reference: estimation.py:73-77."""
    return FALLBACK_H.copy()


def corner_error(H_pred, H_gt, w, h, clip=70.0):
    """Mean L2 distance of the 4 warped image corners, clipped at 70 px.

    reference: estimation.py:79-92 (corners (0,0),(0,h-1),(w-1,0),(w-1,h-1); divide by the third
    homogeneous coordinate; ``mean_dist > 70 -> 70``).
    """
    c = np.array([[0, 0, 1], [0, h - 1, 1], [w - 1, 0, 1], [w - 1, h - 1, 1]], dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = c @ np.asarray(H_gt, dtype=np.float64).T
        a = a[:, :2] / a[:, 2:]
        p = c @ np.asarray(H_pred, dtype=np.float64).T
        p = p[:, :2] / p[:, 2:]
        d = float(np.mean(np.linalg.norm(a - p, axis=1)))
    if d > clip:
        d = clip
    return d


def find_homography_cv2(pos_a, pos_b, thresh=3.0, confidence=0.99999, method="ransac"):
    """The reference's call, verbatim arguments; reference: estimation.py:66-77."""
    import cv2
    try:
        H, mask = cv2.findHomography(np.asarray(pos_a, dtype=np.float32), np.asarray(pos_b, dtype=np.float32),
                                     method=cv2.RANSAC if method == "ransac" else 0,
                                     confidence=confidence, ransacReprojThreshold=thresh)
    except cv2.error:
        H, mask = None, None
    if H is None:
        return fallback_homography(), mask, False
    return H, mask, True


def _normalisation(p, w):
    """Weighted centroid and per-axis inverse mean-absolute-deviation (OpenCV runKernel)."""
    sw = w.sum()
    c = (p * w[:, None]).sum(0) / sw
    s = (np.abs(p - c) * w[:, None]).sum(0)
    return c, sw / s


def weighted_dlt(pa, pb, w=None):
    """Normalised DLT: smallest eigenvector of ``sum_i w_i (Lx_i Lx_i^T + Ly_i Ly_i^T)``.

    Restates OpenCV ``HomographyEstimatorCallback::runKernel`` (the solve behind estimation.py:66)
    with per-point weights (``w_i = 1`` reproduces it).  Returns (H 3x3 float64 with h33 = 1, ok).
    """
    pa = np.asarray(pa, dtype=np.float64)
    pb = np.asarray(pb, dtype=np.float64)
    w = np.ones(len(pa)) if w is None else np.asarray(w, dtype=np.float64)
    if (w > 0).sum() < 4:
        return fallback_homography(), False
    cA, sA = _normalisation(pa, w)
    cB, sB = _normalisation(pb, w)
    if not (np.all(np.isfinite(sA)) and np.all(np.isfinite(sB))):
        return fallback_homography(), False
    X, Y = ((pa - cA) * sA).T
    x, y = ((pb - cB) * sB).T
    one, zero = np.ones_like(X), np.zeros_like(X)
    Lx = np.stack((X, Y, one, zero, zero, zero, -x * X, -x * Y, -x), axis=1)
    Ly = np.stack((zero, zero, zero, X, Y, one, -y * X, -y * Y, -y), axis=1)
    LtL = (Lx * w[:, None]).T @ Lx + (Ly * w[:, None]).T @ Ly
    evals, evecs = np.linalg.eigh(LtL)
    h0 = evecs[:, 0].reshape(3, 3)
    inv_norm_b = np.array([[1 / sB[0], 0, cB[0]], [0, 1 / sB[1], cB[1]], [0, 0, 1]])
    norm_a = np.array([[sA[0], 0, -cA[0] * sA[0]], [0, sA[1], -cA[1] * sA[1]], [0, 0, 1]])
    H = inv_norm_b @ h0 @ norm_a
    if abs(H[2, 2]) < 1e-300:
        return fallback_homography(), False
    return H / H[2, 2], True


def reprojection_error_sq(H, pa, pb):
    """Squared transfer error in image B (OpenCV ``computeError``)."""
    pa = np.asarray(pa, dtype=np.float64)
    pb = np.asarray(pb, dtype=np.float64)
    ww = 1.0 / (H[2, 0] * pa[:, 0] + H[2, 1] * pa[:, 1] + H[2, 2])
    dx = (H[0, 0] * pa[:, 0] + H[0, 1] * pa[:, 1] + H[0, 2]) * ww - pb[:, 0]
    dy = (H[1, 0] * pa[:, 0] + H[1, 1] * pa[:, 1] + H[1, 2]) * ww - pb[:, 1]
    return dx * dx + dy * dy


def refine_homography_lm(H, pa, pb, w=None, iters=10):
    """Weighted Gauss-Newton on the reprojection error over h11..h32 (h33 = 1).

    Restates ``HomographyRefineCallback`` residuals/Jacobian (OpenCV fundam.cpp); OpenCV wraps them
    in a damped LM loop of at most 10 iterations whose damping falls to 0 after two accepted steps,
    so from a DLT start both converge to the same minimiser.  A step is kept only if it lowers the
    weighted cost.
    """
    pa = np.asarray(pa, dtype=np.float64)
    pb = np.asarray(pb, dtype=np.float64)
    w = np.ones(len(pa)) if w is None else np.asarray(w, dtype=np.float64)
    h = (np.asarray(H, dtype=np.float64) / H[2, 2]).reshape(9)[:8].copy()

    def residual(hv):
        ww = 1.0 / (hv[6] * pa[:, 0] + hv[7] * pa[:, 1] + 1.0)
        xi = (hv[0] * pa[:, 0] + hv[1] * pa[:, 1] + hv[2]) * ww
        yi = (hv[3] * pa[:, 0] + hv[4] * pa[:, 1] + hv[5]) * ww
        return ww, xi, yi, xi - pb[:, 0], yi - pb[:, 1]

    ww, xi, yi, rx, ry = residual(h)
    cost = float((w * (rx * rx + ry * ry)).sum())
    for _ in range(iters):
        X, Y = pa[:, 0] * ww, pa[:, 1] * ww
        z = np.zeros_like(X)
        Jx = np.stack((X, Y, ww, z, z, z, -X * xi, -Y * xi), axis=1)
        Jy = np.stack((z, z, z, X, Y, ww, -X * yi, -Y * yi), axis=1)
        A = (Jx * w[:, None]).T @ Jx + (Jy * w[:, None]).T @ Jy
        g = (Jx * w[:, None]).T @ rx + (Jy * w[:, None]).T @ ry
        try:
            d = np.linalg.solve(A, g)
        except np.linalg.LinAlgError:
            break
        hn = h - d
        ww2, xi2, yi2, rx2, ry2 = residual(hn)
        cost2 = float((w * (rx2 * rx2 + ry2 * ry2)).sum())
        if not cost2 < cost:
            break
        h, cost, ww, xi, yi, rx, ry = hn, cost2, ww2, xi2, yi2, rx2, ry2
        if np.max(np.abs(d)) < 1e-12:
            break
    return np.concatenate((h, [1.0])).reshape(3, 3)


# --- counter-based sampling shared bit-for-bit with the CUDA solver (gfnet_b200/csrc/homography.cu)
def _mix32(x):
    x = np.uint32(x)
    with np.errstate(over="ignore"):
        x ^= x >> np.uint32(16)
        x = np.uint32(x * np.uint32(0x7FEB352D))
        x ^= x >> np.uint32(15)
        x = np.uint32(x * np.uint32(0x846CA68B))
        x ^= x >> np.uint32(16)
    return x


def minimal_sample(seed, pair, hyp, n):
    """Four distinct indices in [0, n) for hypothesis ``hyp`` of pair ``pair`` (hash, no state)."""
    idx = []
    ctr = 0
    with np.errstate(over="ignore"):
        base = _mix32(np.uint32(seed) ^ _mix32(np.uint32(pair) * np.uint32(0x9E3779B1) + np.uint32(hyp)))
        while len(idx) < 4:
            v = int(_mix32(base + np.uint32(ctr) * np.uint32(0x85EBCA6B))) % n
            ctr += 1
            if v not in idx:
                idx.append(v)
    return idx


def four_point_homography(pa4, pb4):
    """Exact minimal solve, h33 = 1 (8x8 linear system); non-finite on degenerate samples."""
    A = np.zeros((8, 8))
    b = np.zeros(8)
    for i in range(4):
        X, Y = pa4[i]
        x, y = pb4[i]
        A[2 * i] = [X, Y, 1, 0, 0, 0, -x * X, -x * Y]
        A[2 * i + 1] = [0, 0, 0, X, Y, 1, -y * X, -y * Y]
        b[2 * i], b[2 * i + 1] = x, y
    try:
        h = np.linalg.solve(A, b)
    except np.linalg.LinAlgError:
        return None
    return np.concatenate((h, [1.0])).reshape(3, 3)


def homography_ransac_def(pa, pb, seed=0, pair=0, nhyp=512, thresh=3.0, gn_iters=10, w=None):
    """RANSAC -> re-fit on the best model's inliers -> refine, the structure of cv2.findHomography.

    Same stages as OpenCV (see module docstring) but with ``nhyp`` hash-drawn minimal samples
    evaluated exhaustively instead of OpenCV's adaptive sequential loop; the best hypothesis is the
    one with most inliers (lowest hypothesis index on ties).  Returns (H, mask uint8 [N], ok).
    """
    pa = np.asarray(pa, dtype=np.float64)
    pb = np.asarray(pb, dtype=np.float64)
    n = len(pa)
    if n < 4:
        return fallback_homography(), np.zeros(n, np.uint8), False
    t2 = thresh * thresh
    best_cnt, best_mask = -1, None
    for hyp in range(nhyp):
        idx = minimal_sample(seed, pair, hyp, n)
        H = four_point_homography(pa[idx], pb[idx])
        if H is None or not np.all(np.isfinite(H)):
            continue
        with np.errstate(all="ignore"):
            e = reprojection_error_sq(H, pa, pb)
        mask = e <= t2
        cnt = int(mask.sum())
        if cnt > best_cnt:
            best_cnt, best_mask = cnt, mask
    if best_cnt < 4:
        return fallback_homography(), np.zeros(n, np.uint8), False
    wi = best_mask.astype(np.float64) * (1.0 if w is None else np.asarray(w, dtype=np.float64))
    H, ok = weighted_dlt(pa, pb, wi)
    if not ok:
        return fallback_homography(), best_mask.astype(np.uint8), False
    H = refine_homography_lm(H, pa, pb, wi, iters=gn_iters)
    return H, best_mask.astype(np.uint8), True


# --- OpenCV's RANSAC for cv2.findHomography, restated step by step -------------------------------------------------
# Source restated: OpenCV 4.x modules/calib3d/src/ptsetreg.cpp (RANSACPointSetRegistrator::run / getSubset,
# RANSACUpdateNumIters) and fundam.cpp (HomographyEstimatorCallback::checkSubset / runKernel / computeError,
# findHomography's refit + LM refinement + mask).  Pinned on cv2 4.13.0: tests/golden/homography_cv2_grid.npz
# (generated by tests/golden/make_homography_grid.py with the reference's exact call, estimation.py:66-72).
CV_RNG_COEFF = 4164903690
_FLT_EPS = float(np.finfo(np.float32).eps)
_DBL_MIN = float(np.finfo(np.float64).tiny)


class CvRNG:
    """cv::RNG (multiply-with-carry); findHomography seeds it with (uint64)-1."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * CV_RNG_COEFF + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def _cv_collinear_last(pts):
    i = len(pts) - 1
    for j in range(i):
        dx1, dy1 = float(pts[j][0]) - float(pts[i][0]), float(pts[j][1]) - float(pts[i][1])
        for k in range(j):
            dx2, dy2 = float(pts[k][0]) - float(pts[i][0]), float(pts[k][1]) - float(pts[i][1])
            if abs(dx2 * dy1 - dy2 * dx1) <= _FLT_EPS * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def cv_check_subset(ms1, ms2):
    """HomographyEstimatorCallback::checkSubset for 4 correspondences."""
    if _cv_collinear_last(ms1) or _cv_collinear_last(ms2):
        return False
    negative = 0
    for t in ((0, 1, 2), (1, 2, 3), (0, 2, 3), (0, 1, 3)):
        A = np.array([[ms1[i][0], ms1[i][1], 1.0] for i in t], dtype=np.float64)
        B = np.array([[ms2[i][0], ms2[i][1], 1.0] for i in t], dtype=np.float64)
        negative += bool(np.linalg.det(A) * np.linalg.det(B) < 0)
    return negative in (0, 4)


def cv_update_num_iters(p, ep, model_points, max_iters):
    """RANSACUpdateNumIters."""
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, _DBL_MIN)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < _DBL_MIN:
        return 0
    num, denom = np.log(num), np.log(denom)
    return max_iters if (denom >= 0 or -num >= max_iters * (-denom)) else int(np.rint(num / denom))


def cv_compute_error_f32(H, m1, m2):
    """HomographyEstimatorCallback::computeError: float32 arithmetic on the float-cast model."""
    Hf = np.asarray(H, dtype=np.float64).reshape(9).astype(np.float32)
    M, m = np.asarray(m1, dtype=np.float32), np.asarray(m2, dtype=np.float32)
    one = np.float32(1)
    ww = one / (Hf[6] * M[:, 0] + Hf[7] * M[:, 1] + one)
    dx = (Hf[0] * M[:, 0] + Hf[1] * M[:, 1] + Hf[2]) * ww - m[:, 0]
    dy = (Hf[3] * M[:, 0] + Hf[4] * M[:, 1] + Hf[5]) * ww - m[:, 1]
    return dx * dx + dy * dy


def find_homography_cv_restated(pos_a, pos_b, thresh=3.0, confidence=0.99999, max_iters=2000, refine_iters=10):
    """cv2.findHomography(pos_a, pos_b, cv2.RANSAC, thresh, maxIters=2000, confidence) without cv2.

    Returns (H 3x3 float64, mask uint8 [N] of the refined model, ok, ransac iterations).
    """
    pa, pb = np.asarray(pos_a, dtype=np.float32), np.asarray(pos_b, dtype=np.float32)
    count = len(pa)
    if count < 5:
        raise ValueError("restated for count > 4 (the reference samples 5 000 matches)")
    rng = CvRNG()
    niters, it, max_good = max_iters, 0, 0
    t = np.float32(thresh * thresh)
    best_mask = None
    while it < niters:
        idx, attempts = None, 0
        while attempts < 10000:                      # getSubset
            cand = []
            for _ in range(4):
                v = rng.uniform(0, count)
                while v in cand:
                    v = rng.uniform(0, count)
                cand.append(v)
            if cv_check_subset(pa[cand], pb[cand]):
                idx = cand
                break
            attempts += 1
        if idx is None:
            if it == 0:
                return fallback_homography(), np.zeros(count, np.uint8), False, it
            break
        H4, ok = weighted_dlt(pa[idx].astype(np.float64), pb[idx].astype(np.float64))       # runKernel on the sample
        if ok:
            mask = cv_compute_error_f32(H4, pa, pb) <= t
            good = int(mask.sum())
            if good > max(max_good, 3):
                best_mask, max_good = mask, good
                niters = cv_update_num_iters(confidence, (count - good) / count, 4, niters)
        it += 1
    if best_mask is None:
        return fallback_homography(), np.zeros(count, np.uint8), False, it
    w = best_mask.astype(np.float64)
    H, ok = weighted_dlt(pa.astype(np.float64), pb.astype(np.float64), w)                   # runKernel on the inliers
    if not ok:
        return fallback_homography(), np.zeros(count, np.uint8), False, it
    H = refine_homography_lm(H, pa.astype(np.float64), pb.astype(np.float64), w, iters=refine_iters)   # LMSolver(10)
    final_mask = (cv_compute_error_f32(H, pa, pb) <= t).astype(np.uint8)                    # mask of the refined model
    return H, final_mask, True, it


def homography_from_matches(matches, wq, hq, wsup, hsup, H_gt=None, solver="cv2", **kw):
    """matches[N,4] (normalised) -> pixel coords -> H -> corner error; reference: estimation.py:60-92."""
    m = np.asarray(matches, dtype=np.float32)
    pa, pb = convert_coordinates(m[:, :2], m[:, 2:], wq, hq, wsup, hsup)
    if solver == "cv2":
        H, mask, ok = find_homography_cv2(pa, pb)
    else:
        H, mask, ok = homography_ransac_def(pa, pb, **kw)
    err = corner_error(H, H_gt, wq, hq) if H_gt is not None else None
    return H, mask, ok, err
