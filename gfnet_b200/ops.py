"""Drop-in operators with the reference's Python signatures, backed by the sm_100a kernels.

* ``local_correlation`` -- utils/local_correlation.py:4-16 (same positional/keyword arguments)
* ``kde``               -- utils/kde.py:4
* ``corr_volume`` / ``pos_embed`` / ``coarse_match`` -- GFNet.corr_volume / GFNet.pos_embed,
  model/network.py:415-440

CUDA fp32 tensors only.  Anything unsupported raises (ValueError / NotImplementedError): there is
no CPU, PyTorch or Triton fallback.
"""
import math

import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda_f32, stream_ptr

_SAMPLE_MODES = {"bilinear": 0, "nearest": 1}
_PADDING_MODES = {"zeros": 0, "border": 1}

ALGO_AUTO, ALGO_GENERIC = 0, 1
ALGO_PT = 4                         # one point per thread, whole patch in registers (TMA box per tile); algo = 4 | tune << 4
ALGO_TC2 = 5                        # tcgen05 banded GEMM fed by TMA from a bf16 hi/lo workspace; algo = 5 | group << 4
ALGO_MMA = 6                        # warp-level tensor cores (mma.sync bf16 hi/lo), operands straight from fp32 NCHW; algo = 6 | shape << 4

_TC2_SHAPES = {(2, 32), (3, 32), (4, 32), (2, 64), (3, 64), (4, 64), (5, 64), (6, 64), (7, 64), (8, 64)}   # csrc/local_corr_v2.cu
_TC2_RADII = {32: {2, 3, 4}, 64: {2, 3, 4, 5, 6, 7, 8}}
_PT_SHAPES = {(2, 16), (4, 32), (1, 16), (1, 8), (2, 8)}
_PT_AUTO = {(2, 16), (1, 16), (1, 8), (2, 8)}
_MMA_SHAPES = {(1, 16), (2, 16), (2, 32), (3, 32), (4, 32)} | {(r, 64) for r in range(2, 9)}            # csrc/local_corr_mma.cu
# (r, C) the default mode sends to the mma.sync kernel when the caller did not hoist the tcgen05 pre-pass with
# local_correlation_prepare: one launch instead of pre-pass + plan + main (measured, op batch 64: 0.133 vs 0.167 ms at
# (32,112,64,4), 0.193 vs 0.435 ms at (32,140,80,4); the 64-channel shapes stay on the tcgen05 kernel)
_MMA_AUTO = {(2, 32), (3, 32), (4, 32)}


def _tc2_slice_channels(r, c):
    """Channels per slice when C > 64 is served by the 64-channel tcgen05 kernel slice by slice (0 = not served)."""
    if c > 64 and c % 64 == 0 and r in _TC2_RADII[64]:
        return 64
    return 0


def _mma_slice_channels(r, c):
    return 64 if (c > 64 and c % 64 == 0 and (r, 64) in _MMA_SHAPES) else 0


class PreparedFeatures:
    """feature0 / feature1 of one refiner scale rewritten once for the tcgen05 kernel (``local_correlation_prepare``);
    pass it as ``prepared=`` to every ``local_correlation`` call of that scale (same features, different flow)."""

    def __init__(self, size, f0, f1, r, G, wsbuf, nws):
        self.size, self.f0, self.f1, self.r, self.G, self.wsbuf, self.nws = size, f0, f1, r, G, wsbuf, nws

    def matches(self, size, f0, f1, r, G):
        """Same tensors as at prepare time?  The handle keeps f0 / f1 alive, so equal (pointer, shape, stride) means the
        same storage; the version counter catches in-place writes where autograd tracks one (inference tensors --
        ``GFNet.match`` runs under ``torch.inference_mode``, model/network.py:285 -- have none)."""
        def same(a, b):
            if a.data_ptr() != b.data_ptr() or a.shape != b.shape or a.stride() != b.stride():
                return False
            if a.is_inference() or b.is_inference():
                return True
            return a._version == b._version
        return self.size == tuple(size) and self.r == r and self.G == G and same(self.f0, f0) and same(self.f1, f1)


def local_correlation_prepare(featuremap_size, feature0, feature1, local_radius, num_grid):
    """Hoist the feature pre-pass of the tcgen05 local-correlation kernel out of the iterations of a refiner scale
    (the reference calls ``local_correlation`` ``num_itr`` times per scale with the same ``feature0`` / ``feature1``,
    model/network.py:230-281).  Returns None when the shape is served by a kernel without a pre-pass."""
    B, c, h, w = (int(v) for v in featuremap_size)
    G, r = int(num_grid), int(local_radius)
    if (r, c) not in _TC2_SHAPES:
        return None
    f0 = require_cuda_f32("feature0", feature0)
    f1 = require_cuda_f32("feature1", feature1)
    if f0.shape != (B, c, G, G) or f1.shape != (B, c, h, w):
        raise ValueError("feature0 / feature1 do not match featuremap_size / num_grid")
    if int(lib.gfb_local_corr_tc2_groups(B, c, h, w, G, 0)) != 1:
        return None
    nws = int(lib.gfb_local_corr_tc2_workspace_bytes(B, c, h, w, G, r, 0))
    wsbuf = torch.empty(nws, device=f0.device, dtype=torch.uint8)
    with torch.cuda.device(f0.device):
        check(lib.gfb_local_corr_tc2_prepare_f32(ptr(f0), ptr(f1), B, c, h, w, 0, G, r, ptr(wsbuf), nws, stream_ptr(f0.device)),
              "local_correlation_prepare")
    return PreparedFeatures((B, c, h, w), f0, f1, r, G, wsbuf, nws)


def local_correlation(featuremap_size, feature0, feature1, local_radius, num_grid,
                      padding_mode="zeros", flow=None, im_A_coords=None, sample_mode="bilinear",
                      grid_based_correlation=False, num_level=1, *, algo=ALGO_AUTO, out=None, prepared=None):
    """Flow-displaced (2r+1)^2 window correlation; reference: utils/local_correlation.py:4-72.

    ``feature0 [B,c,G,G]``, ``feature1 [B,c,h,w]``, ``flow [B,2,G,G]`` (or None: identity lattice,
    :21-30) -> ``corr [B, (2r+1)^2 * num_level, G, G]``, ``k = iy*(2r+1)+ix``.  ``im_A_coords`` is
    accepted and ignored exactly as in the reference.
    """
    if sample_mode not in _SAMPLE_MODES:
        raise NotImplementedError(f"sample_mode={sample_mode!r} (supported: bilinear, nearest)")
    if padding_mode not in _PADDING_MODES:
        raise NotImplementedError(f"padding_mode={padding_mode!r} (supported: zeros, border)")
    B, c, h, w = (int(v) for v in featuremap_size)
    f0 = require_cuda_f32("feature0", feature0)
    f1 = require_cuda_f32("feature1", feature1)
    G, r = int(num_grid), int(local_radius)
    if f0.shape != (B, c, G, G):
        raise ValueError(f"feature0 must be [B,c,num_grid,num_grid] = {(B, c, G, G)}, got {tuple(f0.shape)}")
    if f1.dim() != 4 or f1.shape[0] != B or f1.shape[1] != c:
        raise ValueError(f"feature1 must be [B,c,hs,ws] with B={B}, c={c}, got {tuple(f1.shape)}")
    if flow is None:
        if h * w != G * G:
            raise ValueError("flow=None needs h*w == num_grid**2 (the reference reshapes the h x w lattice)")
        ys = torch.linspace(-1 + 1 / h, 1 - 1 / h, h, device=f0.device)
        xs = torch.linspace(-1 + 1 / w, 1 - 1 / w, w, device=f0.device)
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        flow = torch.stack((gx, gy), dim=0).reshape(1, 2, G, G).expand(B, 2, G, G)
    fl = require_cuda_f32("flow", flow)
    if fl.shape != (B, 2, G, G):
        raise ValueError(f"flow must be [B,2,num_grid,num_grid], got {tuple(fl.shape)}")
    kk = (2 * r + 1) ** 2
    num_level = int(num_level)
    if out is None:
        out = torch.empty((B, kk * num_level, G, G), device=f0.device, dtype=f0.dtype)
    elif out.shape != (B, kk * num_level, G, G) or not out.is_contiguous() or out.dtype != torch.float32:
        raise ValueError("out has the wrong shape/layout")
    win_h, win_w = (G, G) if grid_based_correlation else (h, w)
    st = stream_ptr(f0.device)
    base = int(algo) & 15
    if base not in (ALGO_AUTO, ALGO_GENERIC, ALGO_PT, ALGO_TC2, ALGO_MMA):
        raise NotImplementedError(f"local_correlation: unknown algo {algo}")

    def next_level(f1, hs, ws):
        nxt = torch.empty((B, c, hs // 2, ws // 2), device=f1.device, dtype=f1.dtype)    # local_correlation.py:71: avg_pool2d(2, 2)
        check(lib.gfb_avg_pool2_f32(ptr(f1), ptr(nxt), B * c, hs, ws, st), "avg_pool2")
        return nxt

    with torch.cuda.device(f0.device):
        for level in range(num_level):
            hs, ws = int(f1.shape[2]), int(f1.shape[3])
            plain = win_h == hs and win_w == ws and sample_mode == "bilinear" and padding_mode == "zeros"
            small = B * kk * num_level * G * G < 2 ** 31          # the tcgen05 kernel indexes its output with 32 bits
            kt, ko = kk * num_level, kk * level
            if (prepared is not None and level == 0 and num_level == 1 and plain and base == ALGO_AUTO and small
                    and prepared.matches((B, c, h, w), f0, f1, r, G)):
                rc = lib.gfb_local_corr_tc2_run_f32(ptr(f0), ptr(f1), ptr(fl), ptr(out), B, c, hs, ws, 0, G, r, kk, 0,
                                                    ptr(prepared.wsbuf), prepared.nws, st)
                check(rc, "local_correlation (tcgen05, prepared features)")
            elif base == ALGO_MMA or (base == ALGO_AUTO and plain and small and ((r, c) in _MMA_AUTO or (_mma_slice_channels(r, c) and (r, 64) in _MMA_AUTO))):
                cs = _mma_slice_channels(r, c)
                if not (plain and small and ((r, c) in _MMA_SHAPES or cs)):
                    raise NotImplementedError("local_correlation: the mma.sync kernel covers bilinear/zeros with "
                                              f"(r, C) in {sorted(_MMA_SHAPES)} or C a multiple of 64, got r={r}, C={c}")
                src, pitch = f1, 0
                if ws % 4:      # 16-byte cp.async sources: pad each row once (e.g. ws = 70 -> pitch 72)
                    pitch = (ws + 3) // 4 * 4
                    src = torch.empty((B, c, hs, pitch), device=f1.device, dtype=f1.dtype)
                    check(lib.gfb_pad_rows_f32(ptr(f1), ptr(src), B * c * hs, ws, pitch, st), "pad_rows")
                shape = (int(algo) >> 4) & 255 if base == ALGO_MMA else 0
                for c0 in range(0, c, cs or c):     # C = 128, 256, 512 ...: 64-channel slices, the first stores, the others accumulate
                    rc = lib.gfb_local_corr_mma_f32(ptr(f0), ptr(src), ptr(fl), ptr(out), B, cs or c, c, c0, int(c0 > 0), hs, ws, pitch,
                                                    G, r, kt, ko, shape, st)
                    check(rc, "local_correlation (mma.sync)")
            elif base == ALGO_TC2 or (base == ALGO_AUTO and plain and small and ((r, c) in _TC2_SHAPES or _tc2_slice_channels(r, c))):
                cs = _tc2_slice_channels(r, c)
                if not (plain and small and ((r, c) in _TC2_SHAPES or cs)):
                    raise NotImplementedError("local_correlation: the TMA-fed tcgen05 kernel covers bilinear/zeros with "
                                              f"(r, C) in {sorted(_TC2_SHAPES)} or C a multiple of 64, got r={r}, C={c}")
                group = (int(algo) >> 4) & 255 if base == ALGO_TC2 else 0
                if cs:      # C = 128, 256, 512 ...: 64-channel slices, the first stores, the others accumulate
                    nws = int(lib.gfb_local_corr_tc2_workspace_bytes(B, cs, hs, ws, G, r, 0))
                    wsbuf = torch.empty(nws, device=f0.device, dtype=torch.uint8)
                    for c0 in range(0, c, cs):
                        rc = lib.gfb_local_corr_tc2_slice_f32(ptr(f0), ptr(f1), ptr(fl), ptr(out), B, cs, c, c0, int(c0 > 0), hs, ws, 0,
                                                              G, r, kt, ko, ptr(wsbuf), nws, st)
                        check(rc, "local_correlation (tcgen05, channel slice)")
                else:
                    nws = int(lib.gfb_local_corr_tc2_workspace_bytes(B, c, hs, ws, G, r, group))
                    wsbuf = torch.empty(nws, device=f0.device, dtype=torch.uint8)
                    rc = lib.gfb_local_corr_tc2_f32(ptr(f0), ptr(f1), ptr(fl), ptr(out), B, c, hs, ws, 0, G, r, kt, ko, group,
                                                    ptr(wsbuf), nws, st)
                    check(rc, "local_correlation (tcgen05, TMA-fed)")
            elif base == ALGO_PT or (base == ALGO_AUTO and plain and (r, c) in _PT_AUTO and G % 4 == 0):
                if not (plain and (r, c) in _PT_SHAPES):
                    raise NotImplementedError("local_correlation: the point-per-thread kernels cover bilinear/zeros with "
                                              f"(r, C) in {sorted(_PT_SHAPES)}, got r={r}, C={c}")
                src, pitch = f1, 0
                if ws % 4:      # TMA needs 16-byte global strides: pad each row once (e.g. ws = 70 -> pitch 72)
                    pitch = (ws + 3) // 4 * 4
                    src = torch.empty((B, c, hs, pitch), device=f1.device, dtype=f1.dtype)
                    check(lib.gfb_pad_rows_f32(ptr(f1), ptr(src), B * c * hs, ws, pitch, st), "pad_rows")
                rc = lib.gfb_local_corr_pt_f32(ptr(f0), ptr(src), ptr(fl), ptr(out), B, c, hs, ws, pitch, G, r, kt, ko,
                                               (int(algo) >> 4) & 255 if base == ALGO_PT else 0, st)
                check(rc, "local_correlation (point-per-thread)")
            else:
                rc = lib.gfb_local_corr_f32(ptr(f0), ptr(f1), ptr(fl), ptr(out), B, c, hs, ws, 0, G, r, win_h, win_w,
                                            _SAMPLE_MODES[sample_mode], _PADDING_MODES[padding_mode], kt, ko, st)
                check(rc, "local_correlation")
            if level + 1 < num_level:
                f1 = next_level(f1, hs, ws)
    return out


def refiner_input(num_grid, x, y, flow, disp_weight, disp_bias, local_radius, scale_factor=1.0, *, out=None, prepared=None,
                  want_prepared=False, parts=3):
    """The refiner's input tensor ``d = cat(grid_feature, x_hat, emb_in_displacement, local_corr)`` written in place;
    reference: ConvRefiner.forward, model/network.py:537-555 (bilinear ``sample_mode``, ``has_displacement_emb`` and
    ``corr_in_other`` -- the configuration every GFNet scale with a radius uses, :76-135).

    ``x``/``y`` ``[B,c,hs,ws]`` feature maps of image A / B, ``flow [B,2,G,G]``, ``disp_weight [dd,2,1,1]`` / ``disp_bias [dd]``
    = ``ConvRefiner.disp_emb``.  Returns ``d [B, 2c+dd+(2r+1)^2, G, G]``; ``d[:, :c]`` is ``grid_feature``, ``d[:, -(2r+1)^2:]``
    the local correlation.  ``prepared`` / ``want_prepared``: the 64-channel scales run the tcgen05 kernel, whose feature
    pre-pass depends on ``x`` / ``y`` only; pass ``want_prepared=True`` on the first refiner iteration of a scale to get
    ``(d, handle)`` and hand ``prepared=handle`` (with ``out=d``) to the following ones (model/network.py:257-268).
    ``local_radius=None``: no correlation channels (``corr_in_other=False``, :557-558).
    ``parts``: bit 0 = assemble, bit 1 = correlate (bench.py times the two launches separately), bit 2 = ``out[:, :c]`` already
    holds the grid features of this ``x`` (later iterations of a scale: they are not recomputed)."""
    xx = require_cuda_f32("x", x)
    yy = require_cuda_f32("y", y)
    fl = require_cuda_f32("flow", flow)
    B, c, hs, ws = (int(v) for v in xx.shape)
    no_corr = local_radius is None           # corr_in_other=False (the scale-1 refiner, :139-153): d = cat(grid_feature, x_hat, emb)
    G, r = int(num_grid), 0 if no_corr else int(local_radius)
    if no_corr:
        parts &= ~2
    if yy.shape != xx.shape or fl.shape != (B, 2, G, G):
        raise ValueError("x / y must be [B,c,hs,ws] and flow [B,2,num_grid,num_grid]")
    w = require_cuda_f32("disp_weight", disp_weight.reshape(disp_weight.shape[0], -1).float())
    bi = require_cuda_f32("disp_bias", disp_bias.float())
    dd = int(w.shape[0])
    if w.shape != (dd, 2) or bi.shape != (dd,):
        raise ValueError("disp_emb must be a 1x1 convolution 2 -> dd")
    kk = 0 if no_corr else (2 * r + 1) ** 2
    dtot = 2 * c + dd + kk
    if out is None:
        out = torch.empty((B, dtot, G, G), device=xx.device, dtype=torch.float32)
    elif out.shape != (B, dtot, G, G) or not out.is_contiguous() or out.dtype != torch.float32:
        raise ValueError("out has the wrong shape/layout")
    st = stream_ptr(xx.device)
    handle = prepared
    with torch.cuda.device(xx.device):
        use_tc2 = c == 64 and (r, c) in _TC2_SHAPES and int(lib.gfb_local_corr_tc2_groups(B, c, hs, ws, G, 0)) == 1
        if parts & 1:
            check(lib.gfb_refiner_assemble_f32(ptr(xx), ptr(yy), ptr(fl), ptr(w), ptr(bi), ptr(out), B, c, hs, ws, 0, G, dd, dtot,
                                               1.25 * float(scale_factor), int(bool(parts & 4)), st), "refiner_assemble")
        if parts & 2:
            src, pitch = yy, 0
            if ws % 4 and not use_tc2:  # 16-byte row strides for the TMA / cp.async fed kernels: pad each row once
                if handle is not None:
                    src, pitch = handle["padded"], handle["pitch"]
                else:
                    pitch = (ws + 3) // 4 * 4
                    src = torch.empty((B, c, hs, pitch), device=yy.device, dtype=yy.dtype)
                    check(lib.gfb_pad_rows_f32(ptr(yy), ptr(src), B * c * hs, ws, pitch, st), "pad_rows")
            wsbuf, nws, phase = None, 0, 0
            if use_tc2:
                if handle is not None:
                    wsbuf, nws, phase = handle["ws"], handle["nws"], 2
                else:
                    nws = int(lib.gfb_local_corr_tc2_workspace_bytes(B, c, hs, ws, G, r, 0))
                    wsbuf = torch.empty(nws, device=xx.device, dtype=torch.uint8)
            rc = lib.gfb_local_corr_cat_f32(ptr(out), dtot, ptr(src), ptr(fl), B, c, hs, ws, pitch, G, r, 2 * c + dd, phase,
                                            ptr(wsbuf), nws, st)
            if rc == _lib.GFB_EUNSUPPORTED:     # any other (r, C): the generic gather kernel, feature0 copied out of d
                rc = lib.gfb_local_corr_f32(ptr(out[:, :c].contiguous()), ptr(yy), ptr(fl), ptr(out), B, c, hs, ws, 0, G, r, hs, ws,
                                            0, 0, dtot, 2 * c + dd, st)
            check(rc, "refiner_input (local correlation)")
            if handle is None:
                handle = {"ws": wsbuf, "nws": nws, "padded": src, "pitch": pitch, "x": xx, "y": yy}
    return (out, handle) if want_prepared else out


def refiner_input_launches(B, c, hs, ws, G, r, calls=1):
    """Kernels `calls` refiner_input calls on the same features launch (assemble + local correlation, pre-pass hoisted)."""
    n = calls                                                                  # assemble
    if c == 64 and (r, c) in _TC2_SHAPES and int(lib.gfb_local_corr_tc2_groups(B, c, hs, ws, G, 0)) == 1:
        return n + 2 + 2 * (calls - 1)                                         # (pre-pass + plan, main), then (plan, main)
    return n + calls + (1 if ws % 4 else 0)                                    # one kernel per call (+ pad_rows once)


def local_correlation_mma_counters(reset=True):
    """(points of the mma.sync kernel that took the exact gather, 0, 0, 0); synchronises."""
    import ctypes
    buf = (ctypes.c_ulonglong * 4)()
    check(lib.gfb_debug_local_corr_mma_counters(buf, int(reset)), "counters")
    return tuple(int(v) for v in buf)


def local_correlation_v2_counters(reset=True):
    """(lc_pt points on the global-memory path, lc_tc2 points on the gather path, lc_tc2 gather tiles, 0); synchronises."""
    import ctypes
    buf = (ctypes.c_ulonglong * 8)()
    check(lib.gfb_debug_local_corr_v2_counters(buf, int(reset)), "counters")
    return tuple(int(v) for v in buf)


KDE_AUTO, KDE_FULL, KDE_SYMMETRIC = 0, 1, 2
KDE_CUT_SIGMAS = 7.0     # symmetric kernel: tile pairs further apart than 7 std are skipped (terms < 2.3e-11); 0 = none


def kde(x, std=0.1, half=True, down=None, *, algo=KDE_AUTO, cut_sigmas=None):
    """Gaussian kernel density of the rows of ``x [M,D]`` (or batched ``[B,M,D]``).

    reference: utils/kde.py:4-13.  Arithmetic is fp32 (the parity target is the reference's
    ``half=False`` path).  ``half=True`` keeps the reference's dtype contract: the input is
    rounded to fp16 first and an fp16 density is returned, but the sum itself is fp32.
    """
    if not isinstance(x, torch.Tensor) or x.device.type != "cuda":
        raise RuntimeError("kde: x must be a CUDA tensor (no CPU path)")
    xin = x.half().float() if half else x
    xin = require_cuda_f32("x", xin if xin.dtype == torch.float32 else xin.float())
    batched = xin.dim() == 3
    if xin.dim() not in (2, 3):
        raise ValueError("x must be [M,D] or [B,M,D]")
    xb = xin if batched else xin[None]
    B, M, D = (int(v) for v in xb.shape)
    dens = torch.empty((B, M), device=xb.device, dtype=torch.float32)
    dn = 1 if down is None else int(down)
    sym_ok = D == 4 and dn == 1
    if algo == KDE_SYMMETRIC and not sym_ok:
        raise NotImplementedError("kde: the symmetric kernel needs D = 4 and down = 1 (y = x)")
    with torch.cuda.device(xb.device):
        if algo == KDE_SYMMETRIC or (algo == KDE_AUTO and sym_ok and M >= 2048):
            nws = int(lib.gfb_kde_sym_workspace_bytes(B, M))
            wsbuf = torch.empty(nws, device=xb.device, dtype=torch.uint8)
            cs = KDE_CUT_SIGMAS if cut_sigmas is None else float(cut_sigmas)
            rc = lib.gfb_kde_sym_f32(ptr(xb), ptr(dens), B, M, float(std), cs, ptr(wsbuf), nws, stream_ptr(xb.device))
        else:
            rc = lib.gfb_kde_f32(ptr(xb), ptr(dens), B, M, D, dn, float(std), stream_ptr(xb.device))
    check(rc, "kde")
    dens = dens if batched else dens[0]
    return dens.half() if half else dens


class LazyCorrVolume:
    """What ``corr_volume`` returns: the two feature maps, so that ``pos_embed`` can run the fused
    tensor-core kernel without the [B,H1,W1,H0,W0] volume ever touching HBM.  ``materialize()`` (or
    any tensor-like use through ``.tensor``) produces the reference's volume."""

    def __init__(self, feat0, feat1, precision):
        self.feat0, self.feat1, self.precision = feat0, feat1, precision
        B, C, H0, W0 = feat0.shape
        _, _, H1, W1 = feat1.shape
        self.shape = torch.Size((B, H1, W1, H0, W0))
        self._tensor = None

    def materialize(self):
        if self._tensor is None:
            _, self._tensor = _global_match(self.feat0, self.feat1, self.precision, want_volume=True)
        return self._tensor

    tensor = property(materialize)


def _global_match(feat0, feat1, precision=0, want_volume=False, algo=0):
    f0 = require_cuda_f32("feat0", feat0)
    f1 = require_cuda_f32("feat1", feat1)
    if f0.dim() != 4 or f1.dim() != 4 or f0.shape[:2] != f1.shape[:2]:
        raise ValueError("feat0/feat1 must be [B,C,H,W] with equal B and C")
    B, C, H0, W0 = (int(v) for v in f0.shape)
    H1, W1 = int(f1.shape[2]), int(f1.shape[3])
    flow = torch.empty((B, 2, H0, W0), device=f0.device, dtype=torch.float32)
    vol = torch.empty((B, H1, W1, H0, W0), device=f0.device, dtype=torch.float32) if want_volume else None
    with torch.cuda.device(f0.device):
        rc = lib.gfb_global_match_f32(ptr(f0), ptr(f1), ptr(flow), ptr(vol), B, C, H0, W0, H1, W1,
                                      int(precision), int(algo), stream_ptr(f0.device))
    check(rc, "global_match")
    return flow, vol


def coarse_match(feat0, feat1, precision=0, algo=0):
    """``pos_embed(corr_volume(feat0, feat1))`` fused; reference: model/network.py:251-252, 415-440.

    ``precision`` 0 = 3xTF32 split (matches the reference's fp32 einsum to ~1e-6), 1 = one TF32 pass
    (stated tolerance: |dflow| <= 5e-3 normalised on unit-variance features).
    """
    return _global_match(feat0, feat1, precision, False, algo)[0]


def corr_volume(feat0, feat1, precision=0):
    """GFNet.corr_volume (model/network.py:415-428); returns a lazy handle, see LazyCorrVolume."""
    return LazyCorrVolume(feat0, feat1, precision)


def pos_embed(corr):
    """GFNet.pos_embed (model/network.py:430-440) on a LazyCorrVolume or a real [B,H1,W1,H0,W0] tensor."""
    if isinstance(corr, LazyCorrVolume):
        return coarse_match(corr.feat0, corr.feat1, corr.precision)
    v = require_cuda_f32("corr_volume", corr)
    if v.dim() != 5:
        raise ValueError("corr_volume must be [B,H1,W1,H0,W0]")
    B, H1, W1, H0, W0 = (int(s) for s in v.shape)
    flow = torch.empty((B, 2, H0, W0), device=v.device, dtype=torch.float32)
    with torch.cuda.device(v.device):
        check(lib.gfb_pos_embed_f32(ptr(v), ptr(flow), B, H0, W0, H1, W1, stream_ptr(v.device)), "pos_embed")
    return flow


def local_correlation_launches(B, c, hs, ws, G, r, calls=1):
    """Kernels `calls` default-mode local_correlation calls on the same features launch (bilinear, zeros, window = f1);
    with calls > 1 the tcgen05 shapes use ``local_correlation_prepare`` once."""
    if (r, c) in _TC2_SHAPES:
        groups = int(lib.gfb_local_corr_tc2_groups(B, c, hs, ws, G, 0))
        if (r, c) in _MMA_AUTO:
            return calls * (2 if ws % 4 else 1)                               # mma.sync kernel (+ pad_rows)
        if calls > 1 and groups == 1:
            return 1 + 2 * calls                                              # pre-pass once, then plan + main per flow
        return 2 * groups * calls                                             # fused pre-pass + plan, main kernel
    if _tc2_slice_channels(r, c):
        return 2 * (c // 64) * calls
    return calls * (2 if (ws % 4 and (r, c) in _PT_AUTO) else 1)              # (+ pad_rows)


def local_correlation_bytes(B, c, hs, ws, G, r):
    """Algorithmic HBM bytes of one call: read f0, f1, flow once, write corr once."""
    return 4 * B * (c * G * G + c * hs * ws + 2 * G * G + (2 * r + 1) ** 2 * G * G)


def global_match_flops(B, C, N0, N1):
    return 2 * B * N0 * N1 * C


__all__ = ["local_correlation", "kde", "coarse_match", "corr_volume", "pos_embed", "LazyCorrVolume",
           "local_correlation_bytes", "refiner_input", "refiner_input_launches", "local_correlation_launches", "local_correlation_prepare", "PreparedFeatures",
           "global_match_flops", "ALGO_AUTO", "ALGO_GENERIC", "ALGO_PT", "ALGO_TC2", "ALGO_MMA"]
