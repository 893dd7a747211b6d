"""gfnet_b200 -- B200-native (sm_100a) kernels for GFNet's dense-matching + homography hot path.

Python surface = the reference's own (KN-Zhang/GFNet): ``local_correlation``, ``kde``,
``corr_volume`` / ``pos_embed``, ``match_postprocess`` / ``sample``, ``find_homography`` /
``corner_error``.  All of it calls hand-written CUDA through the C ABI in include/gfnet_b200.h.
Importing this package without the built extension raises ImportError (no fallback).
"""
from . import _lib
from .ops import (local_correlation, kde, coarse_match, corr_volume, pos_embed, LazyCorrVolume,
                  local_correlation_bytes, global_match_flops, local_correlation_prepare, refiner_input)
from .matcher import match_postprocess, sample, sample_batched, topk_desc, multinomial_from_noise
from .estimation import (convert_coordinates, estimate_homography, find_homography, corner_error, auc)

__version__ = "0.1.0"
