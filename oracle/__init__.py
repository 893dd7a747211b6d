"""CPU oracle for the GFNet dense-matching + homography hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gfnet_b200/`` may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and there only as the checker / the CPU baseline.

Each function restates, in numpy / torch-CPU, what the reference (KN-Zhang/GFNet @ 2281c4b)
computes on the path and cites the reference file:line it follows.  Two flavours exist:

* ``*_port``  -- follows the reference's own operator sequence (``F.grid_sample`` per batch
  element, ``torch.cdist`` ...) so that timing it on host cores is a fair stand-in for the
  reference's CPU path (``cpu_baseline.kind == "port"``).
* ``*_def``   -- the definition written out in float64 numpy (explicit bilinear taps, explicit
  pairwise distances).  Slow; used at small sizes to check both the port and the CUDA path.

Pinning: the reference ships no tests or golden vectors.  The oracle is
pinned against the reference's own functions imported in place from /root/reference by
``tests/golden/make_golden.py`` (run in the build container; fixtures committed under
``tests/golden/``), and -- for the homography solve, whose arithmetic lives in un-vendored,
un-pinned OpenCV (requirements.txt:2) -- against ``cv2.findHomography`` 4.13.0 called with the
arguments of estimation.py:66-72.
"""
from .local_correlation import local_correlation_port, local_correlation_def
from .kde import kde_port, kde_def
from .coarse_match import corr_volume_port, pos_embed_port, coarse_match_def
from .sampling import (match_postprocess_port, multinomial_from_noise, sample_port,
                       balanced_probability)
from .estimation import (convert_coordinates, corner_error, auc, fallback_homography,
                         find_homography_cv2, weighted_dlt, refine_homography_lm,
                         homography_from_matches, find_homography_cv_restated, cv_compute_error_f32)

__all__ = [n for n in dir() if not n.startswith("_")]
