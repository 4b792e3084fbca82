"""Denoiser wrappers handed to the sampler -- mirror of guided_diffusion/condition_functions.py:17-42.

model_fn / dc_model_fn keep the reference's signature (used through functools.partial, scripts/sample_rule.py:132-136):
class-conditional dispatch, classifier-free guidance as two forwards, and the `rule=` keyword swallowed.  The null
label tensor is cached per (device, batch) instead of being rebuilt on every call (:20, :34).  The classifier-gradient
and DPS hooks (:46-174) are user callables that need autograd through stock-PyTorch classifiers; they plug into
`cond_fn` unchanged and are not reimplemented here.
"""
import torch as th

_null_cache = {}


def _y_null(num_classes, n, device):
    key = (num_classes, n, str(device))
    t = _null_cache.get(key)
    if t is None:
        t = th.full((n,), num_classes, device=device, dtype=th.long)
        _null_cache[key] = t
    return t


def model_fn(x, t, y=None, rule=None, model=None, num_classes=3, class_cond=True, cfg=False, w=0.):
    if not class_cond:
        return model(x, t, _y_null(num_classes, x.shape[0], x.device))
    if cfg:
        return (1 + w) * model(x, t, y) - w * model(x, t, _y_null(num_classes, x.shape[0], x.device))
    return model(x, t, y)


def dc_model_fn(x, t, y=None, rule=None, model=None, num_classes=3, class_cond=True, cfg=False, w=0.):
    """Diff-collage workers score latents laid out [4, pitch, time]; the sampler's are [4, time, pitch]."""
    xp = x.permute(0, 1, 3, 2)
    if not class_cond:
        out = model(xp, t, _y_null(num_classes, x.shape[0], x.device))
    elif cfg:
        out = (1 + w) * model(xp, t, y) - w * model(xp, t, _y_null(num_classes, x.shape[0], x.device))
    else:
        out = model(xp, t, y)
    return out.permute(0, 1, 3, 2)
