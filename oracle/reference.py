"""Import the REFERENCE's own hot-path functions (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

The reference is imported in place, unmodified, from ``/root/reference`` (build container) or from the staged
copy ``baseline/_ref`` (GPU box; ``tools/stage_reference.py``, git-ignored).  ``model/transformer/__init__.py:5``
imports ``romatch`` (not installed anywhere in the reference), so a stub module exposing the reference's own
``utils/utils.py:294,306`` helpers is registered first.

What can be called: ``local_correlation`` (utils/local_correlation.py:4), ``kde`` (utils/kde.py:4), the unbound
``GFNet.corr_volume / pos_embed / sample`` (model/network.py:415, 430, 385) and ``ConvRefiner`` (:444).
``GFNet.__init__`` (downloads DINOv2 weights) and ``estimation.py`` (imports kornia) cannot be used.
"""
import logging
import os
import sys
import types

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = (os.environ.get("GFNET_REFERENCE", "/root/reference"), os.path.join(_ROOT, "baseline", "_ref"))
_cache = None


def reference_root():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "utils", "local_correlation.py")) and \
                os.path.isfile(os.path.join(c, "model", "network.py")):
            return c
    return None


def available():
    return reference_root() is not None


class Reference(types.SimpleNamespace):
    """network (module), local_correlation, kde, GFNet, ConvRefiner, utils_local_correlation, utils_kde, root"""


def load_reference():
    """Returns a ``Reference`` namespace, or raises ``FileNotFoundError`` when neither location exists."""
    global _cache
    if _cache is not None:
        return _cache
    root = reference_root()
    if root is None:
        raise FileNotFoundError("reference not found: neither %s nor %s (run tools/stage_reference.py in the build "
                                "container)" % _CANDIDATES)
    if root not in sys.path:
        sys.path.insert(0, root)
    import utils.utils as U                                  # the reference's own helpers
    for n in ("romatch", "romatch.utils", "romatch.utils.utils"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["romatch.utils.utils"].get_grid = U.get_grid
    sys.modules["romatch.utils.utils"].get_autocast_params = U.get_autocast_params
    prev = logging.root.manager.disable
    logging.disable(logging.WARNING)                         # "xFormers not available"
    try:
        import model.network as N
    finally:
        logging.disable(prev)
    import utils.local_correlation as ULC
    import utils.kde as UK
    _cache = Reference(network=N, local_correlation=ULC.local_correlation, kde=UK.kde, GFNet=N.GFNet,
                       ConvRefiner=N.ConvRefiner, utils_local_correlation=ULC, utils_kde=UK, root=root)
    return _cache


class SampleSelf:
    """Stand-in for ``self`` in the unbound ``GFNet.sample`` (model/network.py:20, 41 defaults)."""
    sample_mode = "threshold_balanced"
    sample_thresh = 0.05


def make_conv_refiner(ref, scale, feat_chs=(64, 32, 16, 8), radius=(7, 6, 4, 2, 0), displacement_dim=None):
    """The reference's ConvRefiner of one scale with the constructor arguments of model/network.py:76-155."""
    import json
    if displacement_dim is None:
        with open(os.path.join(ref.root, "gfnet_configs", "basic.json")) as f:
            displacement_dim = json.load(f)["matcher"]["displacement_dim"]
    idx = {"16": 0, "8": 1, "4": 2, "2": 3, "1": 4}[str(scale)]
    fdim = feat_chs[0] if idx <= 1 else feat_chs[idx - 1]
    kk = (2 * radius[idx] + 1) ** 2 if idx < 4 else 0
    dim = 2 * fdim + displacement_dim[idx] + kk
    return ref.ConvRefiner(dim, dim, 3, kernel_size=5, dw=True, hidden_blocks=8, displacement_emb="linear",
                           displacement_emb_dim=displacement_dim[idx], local_corr_num=radius[idx],
                           corr_in_other=idx < 4, amp=True, disable_local_corr_grad=True, bn_momentum=0.01)
