"""Import shim (test infrastructure): einops_exts.rearrange_many, used by the reference at
unet_model.py:6,286,348. Pure reshape, no arithmetic."""
from einops import rearrange


def rearrange_many(tensors, pattern, **kw):
    return tuple(rearrange(t, pattern, **kw) for t in tensors)
