"""ctypes binding of the C-ABI shared library (include/gfnet_b200.h).

The product path has no fallback: if ``gfnet_b200/_C/libgfnet_b200.so`` is missing the import
fails loudly with the build command.  PyTorch is used for device memory and streams only.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libgfnet_b200.so")

EXPORTS = [
    "gfb_abi_version", "gfb_strerror", "gfb_device_info", "gfb_local_corr_f32", "gfb_avg_pool2_f32",
    "gfb_pad_rows_f32", "gfb_local_corr_pt_f32",
    "gfb_local_corr_tc2_workspace_bytes", "gfb_local_corr_tc2_f32", "gfb_local_corr_tc2_slice_f32", "gfb_local_corr_tc2_groups",
    "gfb_local_corr_tc2_prepare_f32", "gfb_local_corr_tc2_run_f32",
    "gfb_local_corr_mma_f32", "gfb_refiner_assemble_f32", "gfb_local_corr_cat_f32", "gfb_debug_local_corr_mma_counters",
    "gfb_debug_local_corr_v2_counters", "gfb_debug_local_corr_tc2_f32", "gfb_debug_local_corr_pt_f32",
    "gfb_global_match_f32", "gfb_pos_embed_f32", "gfb_kde_f32", "gfb_kde_sym_workspace_bytes", "gfb_kde_sym_f32", "gfb_match_postprocess_f32",
    "gfb_sample_keys_f32", "gfb_balance_keys_f32", "gfb_gather_matches_f32", "gfb_topk_workspace_bytes",
    "gfb_debug_refiner_dw5_planar_f16", "gfb_refiner_pack_f16", "gfb_refiner_dw5_f16", "gfb_refiner_pw_f16", "gfb_refiner_out_f32", "gfb_refiner_blocks_weight_bytes",
    "gfb_refiner_blocks_chunk", "gfb_refiner_blocks_workspace_bytes", "gfb_refiner_blocks_f16", "gfb_flow_update_f32",
    "gfb_upsample_bilinear_f32",
    "gfb_topk_desc_f32", "gfb_homography_workspace_bytes", "gfb_homography_f32", "gfb_homography_cv_f32", "gfb_corner_error_f64",
]

GFB_EINVAL, GFB_EUNSUPPORTED, GFB_EALIGN, GFB_EWORKSPACE, GFB_ENODEVICE = -1, -2, -3, -4, -5


class GfbError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension is not built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C gfnet_b200/csrc` "
            "(needs nvcc >= 12.8, target sm_100a). gfnet_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, f32, i64, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong, ctypes.c_uint
    sz = ctypes.c_size_t
    lib.gfb_abi_version.restype = i32
    lib.gfb_strerror.restype = ctypes.c_char_p
    lib.gfb_strerror.argtypes = [i32]
    lib.gfb_device_info.argtypes = [ctypes.POINTER(i32)] * 3
    lib.gfb_local_corr_f32.argtypes = [vp, vp, vp, vp] + [i32] * 13 + [vp]
    lib.gfb_local_corr_pt_f32.argtypes = [vp, vp, vp, vp] + [i32] * 10 + [vp]
    lib.gfb_debug_local_corr_pt_f32.argtypes = [vp, vp, vp, vp] + [i32] * 11 + [vp]
    lib.gfb_local_corr_tc2_slice_f32.argtypes = [vp, vp, vp, vp] + [i32] * 12 + [vp, sz, vp]
    lib.gfb_debug_local_corr_tc2_f32.argtypes = [vp, vp, vp, vp] + [i32] * 11 + [vp, sz, vp]
    lib.gfb_local_corr_tc2_workspace_bytes.restype = sz
    lib.gfb_local_corr_tc2_workspace_bytes.argtypes = [i32] * 7
    lib.gfb_local_corr_tc2_groups.argtypes = [i32] * 6
    lib.gfb_local_corr_tc2_prepare_f32.argtypes = [vp, vp] + [i32] * 7 + [vp, sz, vp]
    lib.gfb_local_corr_tc2_run_f32.argtypes = [vp, vp, vp, vp] + [i32] * 9 + [vp, sz, vp]
    lib.gfb_local_corr_tc2_f32.argtypes = [vp, vp, vp, vp] + [i32] * 10 + [vp, sz, vp]
    lib.gfb_local_corr_mma_f32.argtypes = [vp, vp, vp, vp] + [i32] * 13 + [vp]
    lib.gfb_refiner_assemble_f32.argtypes = [vp] * 6 + [i32] * 8 + [f32, i32, vp]
    lib.gfb_local_corr_cat_f32.argtypes = [vp, i32, vp, vp] + [i32] * 9 + [vp, sz, vp]
    lib.gfb_refiner_pack_f16.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.gfb_refiner_dw5_f16.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    lib.gfb_debug_refiner_dw5_planar_f16.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.gfb_refiner_pw_f16.argtypes = [vp, vp, vp, vp, i64, i32, i32, vp]
    lib.gfb_refiner_out_f32.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp]
    lib.gfb_refiner_blocks_weight_bytes.restype = sz
    lib.gfb_refiner_blocks_weight_bytes.argtypes = [i32, i32, i32]
    lib.gfb_refiner_blocks_chunk.argtypes = [i32, i32, i32]
    lib.gfb_refiner_blocks_workspace_bytes.restype = sz
    lib.gfb_refiner_blocks_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.gfb_refiner_blocks_f16.argtypes = [vp, vp, vp] + [i32] * 5 + [vp, sz, i32, i32, vp]
    lib.gfb_flow_update_f32.argtypes = [vp, vp, vp, vp] + [i32] * 6 + [vp]
    lib.gfb_upsample_bilinear_f32.argtypes = [vp, vp] + [i32] * 5 + [vp]
    lib.gfb_debug_local_corr_mma_counters.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), i32]
    lib.gfb_debug_local_corr_v2_counters.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), i32]
    lib.gfb_pad_rows_f32.argtypes = [vp, vp, i64, i32, i32, vp]
    lib.gfb_avg_pool2_f32.argtypes = [vp, vp, i32, i32, i32, vp]
    lib.gfb_global_match_f32.argtypes = [vp, vp, vp, vp] + [i32] * 8 + [vp]
    lib.gfb_pos_embed_f32.argtypes = [vp, vp] + [i32] * 5 + [vp]
    lib.gfb_kde_f32.argtypes = [vp, vp, i32, i32, i32, i32, f32, vp]
    lib.gfb_kde_sym_workspace_bytes.restype = sz
    lib.gfb_kde_sym_workspace_bytes.argtypes = [i32, i32]
    lib.gfb_kde_sym_f32.argtypes = [vp, vp, i32, i32, f32, f32, vp, sz, vp]
    lib.gfb_match_postprocess_f32.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp]
    lib.gfb_sample_keys_f32.argtypes = [vp, vp, vp, i64, f32, vp]
    lib.gfb_balance_keys_f32.argtypes = [vp, vp, vp, i64, f32, vp]
    lib.gfb_gather_matches_f32.argtypes = [vp, vp, vp, vp, vp, i32, i64, i32, f32, vp]
    lib.gfb_topk_workspace_bytes.restype = sz
    lib.gfb_topk_workspace_bytes.argtypes = [i32, i64, i32]
    lib.gfb_topk_desc_f32.argtypes = [vp, vp, i32, i64, i32, vp, sz, vp]
    lib.gfb_homography_workspace_bytes.restype = sz
    lib.gfb_homography_workspace_bytes.argtypes = [i32, i32, i32]
    lib.gfb_homography_f32.argtypes = [vp, vp, i32, i32, f32, f32, f32, f32, i32, f32, i32, u32, vp, vp, vp, vp, vp, sz, vp]
    lib.gfb_homography_cv_f32.argtypes = [vp, i32, i32, f32, f32, f32, f32, f32, i32, ctypes.c_double, i32, vp, vp, vp, vp, vp, vp, sz, vp]
    lib.gfb_corner_error_f64.argtypes = [vp, vp, vp, i32, f32, f32, f32, vp]
    for name in EXPORTS:   # getattr raises AttributeError if the library lacks a declared symbol
        if name not in ("gfb_strerror", "gfb_topk_workspace_bytes", "gfb_homography_workspace_bytes",
                        "gfb_local_corr_tc2_workspace_bytes",
                        "gfb_kde_sym_workspace_bytes", "gfb_refiner_blocks_weight_bytes",
                        "gfb_refiner_blocks_workspace_bytes"):
            getattr(lib, name).restype = i32
    if lib.gfb_abi_version() != 2:
        raise ImportError("libgfnet_b200.so ABI version mismatch; rebuild with `make -C gfnet_b200/csrc`")
    return lib


_device_checked = False


def require_sm100():
    """The kernels are sm_100a only (tcgen05 / TMEM / TMA): fail with a clear message on any other GPU instead of a late
    'no kernel image is available' (checked once, at the first op call that has a CUDA tensor in hand)."""
    global _device_checked
    if _device_checked:
        return
    sm, major, minor = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    rc = lib.gfb_device_info(ctypes.byref(sm), ctypes.byref(major), ctypes.byref(minor))
    if rc != 0:
        raise GfbError(f"gfb_device_info: {lib.gfb_strerror(rc).decode()} (code {rc})")
    if major.value != 10:
        raise GfbError(f"gfnet_b200 needs a Blackwell sm_100 GPU (B200); the current device is sm_{major.value}{minor.value}")
    _device_checked = True


lib = _load()


def check(rc, what):
    """0 -> ok; <0 -> ValueError / NotImplementedError (no fallback); >0 -> CUDA error."""
    if rc == 0:
        return
    msg = f"{what}: {lib.gfb_strerror(rc).decode()} (code {rc})"
    if rc == GFB_EUNSUPPORTED:
        raise NotImplementedError(msg)
    if rc < 0:
        raise ValueError(msg)
    raise GfbError(msg)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stream_ptr(device=None):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda_f32(name, t, dtype=None):
    import torch
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if t.device.type != "cuda":
        raise RuntimeError(f"{name} is on {t.device}: gfnet_b200 runs on CUDA (sm_100a) only and has no CPU path")
    require_sm100()
    want = dtype or torch.float32
    if t.dtype != want:
        raise TypeError(f"{name} must be {want}, got {t.dtype}")
    return t.contiguous()
