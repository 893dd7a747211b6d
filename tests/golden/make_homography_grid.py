"""cv2.findHomography fixtures over the SURVEY 8(d2) grid, called exactly as the reference does (estimation.py:66-72).

    python tests/golden/make_homography_grid.py        # needs cv2 (4.13.0 here); writes homography_cv2_grid.npz

Only cv2's OUTPUTS are stored (H, packed inlier mask); the inputs are regenerated from the seeds by ``grid_case`` below,
which tests import.  Cases: sigma in {0, 0.25, 0.5, 1} px x outliers in {0, 10, 20} % at N = 5 000 on 448^2, plus two
low-inlier-ratio cases (60 % / 75 % outliers) that need hundreds / the full 2 000 RANSAC iterations.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [(s, o) for s in (0.0, 0.25, 0.5, 1.0) for o in (0.0, 0.1, 0.2)] + [(0.5, 0.6), (0.5, 0.75)]
N, RES = 5000, 448


def _random_h(rs, jitter=0.15):
    """4-corner jitter recipe (datasets/generate_random_H_large_size.py:6-36) in pixel coordinates."""
    src = np.array([[0, 0], [RES - 1, 0], [RES - 1, RES - 1], [0, RES - 1]], dtype=np.float64)
    dst = src + rs.uniform(-jitter, jitter, (4, 2)) * RES
    A, b = np.zeros((8, 8)), np.zeros(8)
    for i in range(4):
        X, Y = src[i]; x, y = dst[i]
        A[2 * i] = [X, Y, 1, 0, 0, 0, -x * X, -x * Y]; b[2 * i] = x
        A[2 * i + 1] = [0, 0, 0, X, Y, 1, -y * X, -y * Y]; b[2 * i + 1] = y
    return np.concatenate((np.linalg.solve(A, b), [1.0])).reshape(3, 3)


def grid_case(i):
    """(pos_a [N,2] f32, pos_b [N,2] f32, H_true) of case i -- numpy RandomState only, so it reproduces anywhere."""
    sigma, outl = CASES[i]
    rs = np.random.RandomState(1000 + i)
    H = _random_h(rs)
    pa = rs.uniform(0, RES - 1, (N, 2))
    q = np.c_[pa, np.ones(N)] @ H.T
    pb = q[:, :2] / q[:, 2:]
    if sigma > 0:
        pb = pb + rs.normal(0, sigma, (N, 2))
    nout = int(round(outl * N))
    if nout:
        pb[:nout] = rs.uniform(0, RES - 1, (nout, 2))
    return pa.astype(np.float32), pb.astype(np.float32), H


def main():
    import cv2
    out = {"ncases": len(CASES), "cv2_version": cv2.__version__}
    for i in range(len(CASES)):
        pa, pb, Ht = grid_case(i)
        H, mask = cv2.findHomography(pa, pb, method=cv2.RANSAC, confidence=0.99999, ransacReprojThreshold=3)
        out[f"c{i}_H"] = H
        out[f"c{i}_mask"] = np.packbits(mask.ravel().astype(np.uint8))
        out[f"c{i}_ninl"] = int(mask.sum())
        print(i, CASES[i], "inliers", int(mask.sum()))
    np.savez_compressed(os.path.join(HERE, "homography_cv2_grid.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
