"""Denoiser wrappers handed to the sampler -- mirror of guided_diffusion/condition_functions.py:17-42.

model_fn / dc_model_fn keep the reference's signature (used through functools.partial, scripts/sample_rule.py:132-136):
class-conditional dispatch, classifier-free guidance as two forwards, and the `rule=` keyword swallowed.  The null
label tensor is cached per (device, batch) instead of being rebuilt on every call (:20, :34).  The classifier-gradient
hooks (:45-86, `composite_nn_zt` :161-167) differentiate a stock-PyTorch classifier the caller supplies with respect to
the noisy latent; they are host-side autograd plumbing around that classifier (nothing of this package is
differentiated), kept here so that scripts/sample_rule.py:30-31, 109-112 imports and runs unchanged.  The DPS hooks
(`nn_z0_*`, `rule_x0_*`, `composite_rule` :88-174) need gradients THROUGH the denoiser and the VAE decoder, which this
inference-only path does not provide: they raise.
"""
import torch as th
import torch.nn.functional as F

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


# ---- classifier guidance: gradient of the classifier's log-probability of the target rule w.r.t. x_t -----------------
def _input_gradient(x, log_prob_of):
    """d/dx of sum_b log_prob_of(x)[b], evaluated at x (autograd through the caller's classifier only)."""
    with th.enable_grad():
        x_in = x.detach().requires_grad_(True)
        return th.autograd.grad(log_prob_of(x_in).sum(), x_in)[0]


def grad_nn_zt_xentropy(x, y=None, rule=None, classifier=None):
    """Class-label rules (:45-55): log-softmax probability of the target class; the classifier is queried at t = 0."""
    assert rule is not None
    t0 = th.zeros(x.shape[0], device=x.device)

    def log_prob(x_in):
        lp = F.log_softmax(classifier(x_in, t0), dim=-1)
        return lp.gather(1, rule.view(-1, 1).long())

    return _input_gradient(x, log_prob)


def grad_nn_zt_mse(x, t, y=None, rule=None, classifier_scale=10., classifier=None):
    """Regression rules (:58-64): -sum of squared errors between the classifier's prediction and the target."""
    assert rule is not None
    return _input_gradient(x, lambda x_in: -((classifier(x_in, t) - rule) ** 2).sum(dim=-1)) * classifier_scale


def grad_nn_zt_chord(x, t, y=None, rule=None, classifier_scale=10., classifier=None, both=False):
    """Chord rules (:67-86): cross-entropy of the per-window chord logits (and of the key logits when `both`)."""
    assert rule is not None

    def log_prob(x_in):
        key_logits, chord_logits = classifier(x_in, t)
        flat = chord_logits.reshape(-1, chord_logits.shape[-1])
        if not both:
            return -F.cross_entropy(flat, rule.reshape(-1), reduction="none")
        key_lp = -F.cross_entropy(key_logits, rule[:, :1], reduction="none")
        chord_lp = -F.cross_entropy(flat, rule[:, 1:].reshape(-1), reduction="none")
        return key_lp + chord_lp.reshape(x_in.shape[0], -1).mean(dim=-1)

    return _input_gradient(x, log_prob) * classifier_scale


def _needs_denoiser_gradient(name):
    def fn(*args, **kwargs):
        raise NotImplementedError(f"{name} is a DPS hook: it needs gradients through the denoiser / VAE decoder, which "
                                  "the inference-only B200 path does not provide (DESIGN.md section 7)")
    fn.__name__ = name
    return fn


nn_z0_chord_dummy = _needs_denoiser_gradient("nn_z0_chord_dummy")
nn_z0_mse_dummy = _needs_denoiser_gradient("nn_z0_mse_dummy")
nn_z0_mse = _needs_denoiser_gradient("nn_z0_mse")
rule_x0_mse_dummy = _needs_denoiser_gradient("rule_x0_mse_dummy")
rule_x0_mse = _needs_denoiser_gradient("rule_x0_mse")

function_map = {
    "grad_nn_zt_xentropy": grad_nn_zt_xentropy, "grad_nn_zt_mse": grad_nn_zt_mse, "grad_nn_zt_chord": grad_nn_zt_chord,
    "nn_z0_chord_dummy": nn_z0_chord_dummy, "nn_z0_mse_dummy": nn_z0_mse_dummy, "nn_z0_mse": nn_z0_mse,
    "rule_x0_mse_dummy": rule_x0_mse_dummy, "rule_x0_mse": rule_x0_mse,
}


def composite_nn_zt(x, t, y=None, rule=None, fns=None, classifier_scales=None, classifiers=None, rule_names=None):
    """Sum of the classifier gradients of several rules (:161-167): fns[i] names the hook for rule_names[i]."""
    out = 0
    for fn_name, scale, clf, name in zip(fns, classifier_scales, classifiers, rule_names):
        out = out + function_map[fn_name](x, t, y=y, rule=rule[name], classifier_scale=scale, classifier=clf)
    return out


def composite_rule(x, t, y=None, rule=None, fns=None, classifier_scales=None, rule_names=None):
    """DPS on differentiable rules (:170-174): needs gradients through the decoder -- every hook it can name raises."""
    out = 0
    for fn_name, scale, name in zip(fns, classifier_scales, rule_names):
        out = out + function_map[fn_name](x, t, y=y, rule=rule[name], rule_name=name) * scale
    return out

