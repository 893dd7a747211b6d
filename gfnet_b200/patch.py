"""Rebind the reference's hot-path names to the B200 kernels (drop-in switch).

The reference binds with ``from utils.kde import kde`` / ``from utils.local_correlation import
local_correlation`` (model/network.py:10-11), so the *module attributes of model.network* must be
replaced, not only utils.*; ``GFNet.corr_volume`` / ``pos_embed`` / ``sample`` are methods.
"""
from . import ops, matcher


def patch(network_module, utils_local_correlation=None, utils_kde=None):
    """``patch(model.network)``; returns a dict of the originals so ``unpatch`` can restore them."""
    saved = {
        "local_correlation": network_module.local_correlation,
        "kde": network_module.kde,
        "corr_volume": network_module.GFNet.corr_volume,
        "pos_embed": network_module.GFNet.pos_embed,
        "sample": network_module.GFNet.sample,
    }
    network_module.local_correlation = ops.local_correlation
    network_module.kde = ops.kde
    network_module.GFNet.corr_volume = lambda self, feat0, feat1: ops.corr_volume(feat0, feat1)
    network_module.GFNet.pos_embed = lambda self, corr_volume: ops.pos_embed(corr_volume)

    def _sample(self, matches, certainty, num=5000):
        return matcher.sample(matches, certainty, num, sample_mode=self.sample_mode, sample_thresh=self.sample_thresh)
    network_module.GFNet.sample = _sample
    if utils_local_correlation is not None:
        utils_local_correlation.local_correlation = ops.local_correlation
    if utils_kde is not None:
        utils_kde.kde = ops.kde
    return saved


def unpatch(network_module, saved):
    network_module.local_correlation = saved["local_correlation"]
    network_module.kde = saved["kde"]
    network_module.GFNet.corr_volume = saved["corr_volume"]
    network_module.GFNet.pos_embed = saved["pos_embed"]
    network_module.GFNet.sample = saved["sample"]
