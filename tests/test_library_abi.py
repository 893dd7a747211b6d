"""CPU-side checks of the boundary: the shared library loads and exports every symbol the header
declares; host-side argument validation returns the documented codes without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gfnet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gfb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from gfnet_b200 import _lib
    names = _declared()
    assert len(names) >= 17
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/gfnet_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert _lib.lib.gfb_abi_version() == 2


def test_error_strings_and_host_side_validation():
    from gfnet_b200 import _lib
    lib = _lib.lib
    assert lib.gfb_strerror(0) == b"ok"
    assert b"invalid" in lib.gfb_strerror(-1) and b"not supported" in lib.gfb_strerror(-2)
    null = ctypes.c_void_p(0)
    assert lib.gfb_kde_f32(null, null, 1, 10, 4, 1, 0.1, null) == _lib.GFB_EINVAL
    assert lib.gfb_local_corr_f32(null, null, null, null, *([1] * 13), null) == _lib.GFB_EINVAL
    assert lib.gfb_topk_workspace_bytes(2, 1000, 100) == 2 * 2 * 5120 * 4
    assert lib.gfb_topk_workspace_bytes(1, 204800, 20000) == 2 * 20480 * 4
    assert lib.gfb_homography_workspace_bytes(3, 5000, 512) >= 3 * 8 + 3 * 512 * 72
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(NotImplementedError):
        _lib.check(-2, "x")


def test_ops_refuse_cpu_tensors():
    import torch
    import gfnet_b200 as gf
    x = torch.rand(10, 4)
    with pytest.raises(RuntimeError):
        gf.kde(x, half=False)
    with pytest.raises(RuntimeError):
        gf.coarse_match(torch.rand(1, 8, 4, 4), torch.rand(1, 8, 4, 4))
    with pytest.raises(RuntimeError):
        gf.match_postprocess(torch.rand(2, 2, 4, 4), torch.rand(2, 1, 4, 4))


def test_synthetic_batch_shapes_and_algorithmic_bytes():
    import gfnet_b200 as gf
    from gfnet_b200 import synth
    cfg = synth.pyramid_config(448)
    assert cfg == [(16, 64, 32, 32, 7), (8, 64, 56, 32, 6), (4, 32, 112, 64, 4), (2, 16, 224, 128, 2)]
    assert synth.pyramid_config(448, upsample_res=560) == [(8, 64, 70, 40, 6), (4, 32, 140, 80, 4), (2, 16, 280, 160, 2)]
    assert synth.final_grid(448, 560) == 320
    # 12.739 MB per batch element for pass 1, 17.632 MB for the 560 pass
    p1 = sum(gf.local_correlation_bytes(1, c, hs, hs, g, r) for (_, c, hs, g, r) in cfg)
    p2 = sum(gf.local_correlation_bytes(1, c, hs, hs, g, r) for (_, c, hs, g, r) in synth.pyramid_config(448, upsample_res=560))
    assert abs(p1 / 1e6 - 12.739) < 0.01 and abs(p2 / 1e6 - 17.632) < 0.01
    b = synth.PairBatch(1, device="cpu")
    assert b.final_flow.shape == (2, 2, 320, 320) and len(b.passes) == 2


def test_launch_accounting_and_workspace_queries():
    """Host-only entry points: launch count of the default local_correlation path and workspace sizes."""
    from gfnet_b200 import synth
    from gfnet_b200._lib import lib
    from gfnet_b200.ops import local_correlation_launches
    from gfnet_b200.pipeline import HotPath
    # C = 64: fused pre-pass + main kernel, one pair per call; C = 32 without a hoisted pre-pass: the mma.sync kernel;
    # C = 16: one persistent kernel
    assert local_correlation_launches(64, 64, 32, 32, 32, 7) == 2
    assert local_correlation_launches(64, 32, 140, 140, 80, 4) == 1
    assert local_correlation_launches(64, 16, 224, 224, 128, 2) == 1
    assert lib.gfb_local_corr_tc2_groups(64, 32, 140, 140, 80, 16) == 4        # explicit groups of 16 elements
    assert lib.gfb_local_corr_tc2_workspace_bytes(64, 32, 140, 140, 80, 4, 0) > 64 * (140 * 140 + 80 * 80) * 32 * 4
    assert lib.gfb_kde_sym_workspace_bytes(32, 20000) >= 32 * 20000 * (8 + 4 + 8 + 16)
    batch = synth.PairBatch(1, num_itr=2, device="cpu")
    assert local_correlation_launches(64, 64, 70, 70, 40, 6, calls=2) == 5      # pre-pass hoisted: 1 + 2 x (plan, main)
    assert local_correlation_launches(64, 32, 140, 140, 80, 4, calls=2) == 2    # mma.sync kernel: one launch per call
    assert HotPath().kernel_launches(batch) == 50          # 34 refiner-input launches (14 assemble, 20 correlation) + 16; cv2-faithful solver: one RANSAC kernel
    assert HotPath(n_hyp=512).kernel_launches(batch) == 51
