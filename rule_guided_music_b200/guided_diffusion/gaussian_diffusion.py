"""Sampling half of GaussianDiffusion with Stochastic Control Guidance -- host-side mirror of
guided_diffusion/gaussian_diffusion.py of the reference, driving the CUDA kernels of librgm_b200.so.

API kept (reference file:line): get_named_beta_schedule :31-62, ModelMeanType / ModelVarType / LossType :85-120,
GaussianDiffusion.__init__ tables :138-189, q_sample :207-226, q_posterior_mean_variance :228-250, p_mean_variance
:252-357, _predict_xstart_from_eps :359-364, _predict_eps_from_xstart :376-380, condition_mean :387-465 (classifier
branch), condition_score :467-489, scg_sample :491-633, p_sample :635-735, p_sample_loop(_progressive) :737-879,
ddim_sample :881-976, ddim_sample_loop(_progressive) :1016-1143, and the module helpers _extract_into_tensor :1331,
_decode :1347, _extract_rule :1361, guide_schedule :1398.  _encode :1382 (VAE encoder, for scripts/edit.py).  Not on this path
(raise NotImplementedError): training losses, bpd evaluation, DDIM reverse sampling and the DPS branch of
condition_mean, which needs autograd through the denoiser.

What differs from the reference is only HOW a step runs:
  * the schedule tables are cast to fp32 once per device (same float64 -> index -> .float() values as
    _extract_into_tensor) and gathered on the device, instead of one host->device copy per lookup;
  * guidance decisions are made from the host-side timestep index the loops already know, so a step has no
    device->host synchronisation;
  * scg_sample is the fused path: fan-out kernel, ONE batched denoiser call over N*B candidates, x0 kernel, fused
    _decode (re-tiling + VAE decoder writing channel 0 of the roll directly), rule reduction kernels, loss/weight
    accumulation and first-max argmax + gather on the device (csrc/rules.cu).  Rules or losses the user registered
    that the kernels do not know are called as Python callables on the materialised roll, like the reference does.
Quirks of the reference are preserved on purpose: candidates are scored at timestep t; p_sample
passes the unwrapped model to scg_sample while ddim_sample wraps it; classifier guidance applies at every step when SCG
is on; DDPM masks noise with t > t_end, DDIM with t != t_end; rules mutate the roll in place in dict order.
"""
import collections
import collections.abc
import enum
import functools
import math
import types

import numpy as np
import torch as th

from .. import _lib
from . import dist_util
from ..music_rule_guidance import music_rules as _mr
from ..music_rule_guidance.rule_maps import FUNC_DICT, LOSS_DICT, NATIVE_LOSS_KIND


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps):
    """reference :31-62."""
    if schedule_name == "linear":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    if schedule_name == "stable-diffusion":
        scale = 1000 / num_diffusion_timesteps
        return np.linspace(scale * math.sqrt(0.00085), scale * math.sqrt(0.012), num_diffusion_timesteps,
                           dtype=np.float64) ** 2
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """reference :65-82."""
    out = []
    for i in range(num_diffusion_timesteps):
        t1, t2 = i / num_diffusion_timesteps, (i + 1) / num_diffusion_timesteps
        out.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return np.array(out)


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    """reference :1331-1344 (kept for callers that index arbitrary arrays; the sampler itself uses device tables)."""
    res = th.from_numpy(np.asarray(arr)).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)


def guide_schedule(t, t_start=750, t_end=0, interval=1):
    """reference :1398-1400; `t` may be the step tensor (reads t[0], a device sync) or a host integer."""
    t0 = int(t[0]) if hasattr(t, "__getitem__") else int(t)
    return bool(t_start > t0 >= t_end and (t0 + 1) % interval == 0)


def _decode(pred_zstart, embed_model, scale_factor=1., threshold=False):
    """reference :1347-1358.  Uses the fused decoder when embed_model is the native AutoencoderKL."""
    if hasattr(embed_model, "decode_latents"):
        roll = embed_model.decode_latents(pred_zstart, scale_factor)
    else:
        h, w = pred_zstart.shape[-2], pred_zstart.shape[-1]
        s = (pred_zstart / scale_factor).permute(0, 1, 3, 2)
        s = th.concat(th.chunk(s, h // w, dim=-1), dim=0)
        s = embed_model.decode(s)
        roll = th.concat(th.chunk(s, h // w, dim=0), dim=-1)
    if threshold:
        roll[roll <= -0.95] = -1.
    return roll


def _extract_rule(rule_name, pred_xstart):
    """reference :1361-1379 (the chord rule's process pool is the user's callable's business here)."""
    return FUNC_DICT[rule_name](pred_xstart)


def _encode(pred_xstart, embed_model, scale_factor=1.):
    """reference :1382-1395: piano roll [B, 3, 128, L] -> latent [B, 4, L/8, 16] (posterior mean * scale_factor)."""
    h, w = pred_xstart.shape[-2], pred_xstart.shape[-1]
    seq_len = w // h
    micro = th.concat(th.chunk(pred_xstart, seq_len, dim=-1), dim=0)  # 1st second for all batch, 2nd second, ...
    micro = embed_model.encode_save(micro, range_fix=False)
    z = th.chunk(micro, 2, dim=1)[0] if micro.shape[1] == 8 else micro
    z = th.concat(th.chunk(z, seq_len, dim=0), dim=-1)
    return z.permute(0, 1, 3, 2) * scale_factor


_TABLES = ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
           "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_variance",
           "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2")


class _StepGraph:
    """One captured step: static input buffers, the graph, the tensors it leaves its results in, and strong references
    to every object the signature identified by `id()` (a collected object's id can be reused by a different one, and a
    graph replays with the addresses and Python constants it was captured with)."""
    __slots__ = ("graph", "x", "t", "out", "launches", "refs")

    def __init__(self, refs=None):
        self.graph = self.x = self.t = self.out = self.launches = None
        self.refs = refs


class GaussianDiffusion:
    """reference :123-189 (constructor and tables) + the sampling methods."""

    def __init__(self, *, betas, model_mean_type, model_var_type, loss_type, rescale_timesteps=False):
        self.model_mean_type = model_mean_type
        self.model_var_type = model_var_type
        self.loss_type = loss_type
        self.rescale_timesteps = rescale_timesteps
        betas = np.array(betas, dtype=np.float64)
        self.betas = betas
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self.t_end = 0
        self._dev_tables = {}
        # record=True bookkeeping (reference :778-787)
        self.log_probs, self.each_loss = [], {}
        # debugging / test hook: when set to a list, every SCG decision appends (total_log_prob [N, B], chosen index [B])
        self._trace = None
        # CUDA-graph replay of whole steps (enable_cuda_graphs): signature -> _StepGraph
        self._graphs_on = False
        self._graphs = collections.OrderedDict()

    # ---- device-resident tables ----------------------------------------------------------------------------------
    def _tables(self, device):
        key = str(device)
        tab = self._dev_tables.get(key)
        if tab is None:
            tab = {n: th.from_numpy(getattr(self, n)).to(device).float() for n in _TABLES}
            fl = np.append(self.posterior_variance[1], self.betas[1:])
            tab["fixed_large_variance"] = th.from_numpy(fl).to(device).float()
            tab["fixed_large_log_variance"] = th.from_numpy(np.log(fl)).to(device).float()
            tab["log_betas"] = th.from_numpy(np.log(self.betas)).to(device).float()
            self._dev_tables[key] = tab
        return tab

    def _coef(self, name, t, ndim):
        v = self._tables(t.device)[name][t]
        return v.view(-1, *([1] * (ndim - 1)))

    def _scale_timesteps(self, t):
        if self.rescale_timesteps:
            return t.float() * (1000.0 / self.num_timesteps)
        return t

    def _wrap_model(self, model):
        return model

    # ---- closed forms (reference :207-250, :359-380) ----------------------------------------------------------
    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = th.randn_like(x_start)
        return (self._coef("sqrt_alphas_cumprod", t, x_start.dim()) * x_start
                + self._coef("sqrt_one_minus_alphas_cumprod", t, x_start.dim()) * noise)

    def q_posterior_mean_variance(self, x_start, x_t, t):
        n = x_t.dim()
        mean = self._coef("posterior_mean_coef1", t, n) * x_start + self._coef("posterior_mean_coef2", t, n) * x_t
        var = self._coef("posterior_variance", t, n).expand(x_t.shape)
        logvar = self._coef("posterior_log_variance_clipped", t, n).expand(x_t.shape)
        return mean, var, logvar

    def _predict_xstart_from_eps(self, x_t, t, eps):
        assert x_t.shape == eps.shape
        n = x_t.dim()
        return self._coef("sqrt_recip_alphas_cumprod", t, n) * x_t - self._coef("sqrt_recipm1_alphas_cumprod", t, n) * eps

    def _predict_eps_from_xstart(self, x_t, t, pred_xstart):
        n = x_t.dim()
        return ((self._coef("sqrt_recip_alphas_cumprod", t, n) * x_t - pred_xstart)
                / self._coef("sqrt_recipm1_alphas_cumprod", t, n))

    def _predict_xstart_from_xprev(self, x_t, t, xprev):
        n = x_t.dim()
        c1 = self._coef("posterior_mean_coef1", t, n)
        return (1.0 / c1) * xprev - (self._coef("posterior_mean_coef2", t, n) / c1) * x_t

    # ---- p(x_{t-1} | x_t) (reference :252-357) -------------------------------------------------------------------
    def p_mean_variance(self, model, x, t, clip_denoised=True, denoised_fn=None, model_kwargs=None, cond_fn=None,
                        embed_model=None, edit_kwargs=None):
        if model_kwargs is None:
            model_kwargs = {}
        B, C = x.shape[:2]
        assert t.shape == (B,)

        def process_xstart(v):
            if denoised_fn is not None:
                v = denoised_fn(v)
            return v.clamp(-1, 1) if clip_denoised else v

        model_output = model(x, self._scale_timesteps(t), **model_kwargs)
        if edit_kwargs is not None:
            x0 = process_xstart(self._predict_xstart_from_eps(x, t, model_output))
            replaced = edit_kwargs["mask"] * edit_kwargs["gt"] + (1 - edit_kwargs["mask"]) * x0
            model_output = self._predict_eps_from_xstart(x, t, replaced)

        n = x.dim()
        if self.model_var_type in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE):
            assert model_output.shape == (B, C * 2, *x.shape[2:])
            model_output, var_values = th.split(model_output, C, dim=1)
            if self.model_var_type == ModelVarType.LEARNED:
                log_variance = var_values
            else:
                min_log = self._coef("posterior_log_variance_clipped", t, n)
                max_log = self._coef("log_betas", t, n)
                frac = (var_values + 1) / 2
                log_variance = frac * max_log + (1 - frac) * min_log
            variance = th.exp(log_variance)
        elif self.model_var_type == ModelVarType.FIXED_LARGE:
            variance = self._coef("fixed_large_variance", t, n).expand(x.shape)
            log_variance = self._coef("fixed_large_log_variance", t, n).expand(x.shape)
        else:
            variance = self._coef("posterior_variance", t, n).expand(x.shape)
            log_variance = self._coef("posterior_log_variance_clipped", t, n).expand(x.shape)

        if self.model_mean_type == ModelMeanType.PREVIOUS_X:
            pred_xstart = process_xstart(self._predict_xstart_from_xprev(x, t, model_output))
            mean = model_output
        elif self.model_mean_type in (ModelMeanType.START_X, ModelMeanType.EPSILON):
            if self.model_mean_type == ModelMeanType.START_X:
                pred_xstart = process_xstart(model_output)
            else:
                pred_xstart = process_xstart(self._predict_xstart_from_eps(x, t, model_output))
            mean, _, _ = self.q_posterior_mean_variance(pred_xstart, x, t)
        else:
            raise NotImplementedError(self.model_mean_type)
        assert mean.shape == log_variance.shape == pred_xstart.shape == x.shape
        return {"mean": mean, "variance": variance, "log_variance": log_variance, "pred_xstart": pred_xstart}

    # ---- classifier guidance hooks (reference :387-489) ------------------------------------------------------------
    def condition_mean(self, cond_fn, p_mean_var, x, t, model_kwargs=None, guidance_kwargs=None, model=None,
                       embed_model=None, edit_kwargs=None, scale_factor=1., record=False):
        if guidance_kwargs is not None and getattr(guidance_kwargs, "method", None) == "dps":
            raise NotImplementedError("DPS guidance differentiates through the denoiser; the B200 denoiser is "
                                      "inference-only (SURVEY.md section 8f)")
        model_kwargs = model_kwargs or {}
        if edit_kwargs is None:
            gradient = cond_fn(x, self._scale_timesteps(t), **model_kwargs)
            return p_mean_var["mean"].float() + p_mean_var["variance"] * gradient.float()
        ls, le = edit_kwargs["l_start"], edit_kwargs["l_end"]
        gradient = cond_fn(x[:, :, ls:le, :], self._scale_timesteps(t), **model_kwargs)
        new_mean = p_mean_var["mean"].float().clone()
        new_mean[:, :, ls:le, :] += p_mean_var["variance"][:, :, ls:le, :] * gradient.float()
        return new_mean

    def condition_score(self, cond_fn, p_mean_var, x, t, model_kwargs=None):
        model_kwargs = model_kwargs or {}
        alpha_bar = self._coef("alphas_cumprod", t, x.dim())
        eps = self._predict_eps_from_xstart(x, t, p_mean_var["pred_xstart"])
        eps = eps - (1 - alpha_bar).sqrt() * cond_fn(x, self._scale_timesteps(t), **model_kwargs)
        out = dict(p_mean_var)
        out["pred_xstart"] = self._predict_xstart_from_eps(x, t, eps)
        out["mean"], _, _ = self.q_posterior_mean_variance(out["pred_xstart"], x, t)
        return out

    # ---- stochastic control guidance (reference :491-633) ---------------------------------------------------------
    def _score_candidates(self, roll, model_kwargs, scg_kwargs, num_samples, B, seg=None):
        """Σ_rules weight * -loss over the candidates' rolls, in dict order (reference :531-538).  Returns
        total_log_prob [N*B] (fp32, device) and the per-rule losses (for record)."""
        dev = roll.device
        total = th.zeros(roll.shape[0], device=dev, dtype=th.float32)
        each = {}
        stream = _lib.stream_ptr()
        for rule_name, rule_target in model_kwargs["rule"].items():
            if seg is not None:
                rule_target = seg(rule_name, rule_target)
            func = FUNC_DICT[rule_name]
            loss_fn = LOSS_DICT[rule_name]
            weight = float(scg_kwargs.get(rule_name, 1.))
            gen = func(roll)  # native rules launch their reduction kernels; user rules run as given
            kind = NATIVE_LOSS_KIND.get(loss_fn)
            if kind == 1 and gen.dim() == 2 and gen.is_cuda and not gen.is_floating_point():
                gen = gen.to(th.float32)  # class indices (note_density_class): exact in fp32, compared with !=
            if kind is not None and gen.dim() == 2 and gen.dtype == th.float32 and gen.is_cuda:
                tgt = rule_target.to(dev, th.float32).contiguous()
                gen = gen.contiguous()
                if tuple(tgt.shape) != (B, gen.shape[1]):  # the kernel indexes target[(i % B) * K + k]
                    raise _lib.RgmError(f"rule '{rule_name}': target shape {tuple(tgt.shape)} does not match the rule "
                                        f"output [{B}, {gen.shape[1]}] (one row per sample)")
                _lib.call("rgm_rule_loss_accum", _lib.ptr(gen), _lib.ptr(tgt), _lib.ptr(total), gen.shape[0], B,
                          gen.shape[1], kind, weight, stream)
                each[rule_name] = (gen, tgt, loss_fn)
            else:
                if gen.dim() == 1:
                    gen = gen.unsqueeze(0)
                y_ = rule_target.to(gen.device).repeat(num_samples, 1)
                log_prob = -loss_fn(gen, y_)
                total = total + (log_prob * weight).to(dev, th.float32)
                each[rule_name] = (gen, rule_target, loss_fn)
        return total, each

    def scg_sample(self, model, t, mean_pred, g_coeff, embed_model, scale_factor, model_kwargs=None, scg_kwargs=None,
                   edit_kwargs=None, dc_kwargs=None, record=False, record_freq=100):
        """Fan each sample out to N candidate x_{t-1}, score their decoded x0 with the rules, keep the best.
        With dist_util.shard_candidates() on, this rank fans out, denoises, decodes and scores only its contiguous
        share of the N candidates and the ranks exchange their per-sample winners once (dist_util.first_max_over_ranks)."""
        # g_coeff is exp(0.5*log_variance) or sigma expanded to x's shape: one value per sample for the fixed
        # variances.  A learned (per-element) variance cannot reach this point in the reference either: its scg_sample
        # fails the shape assert of _predict_xstart_from_eps on the 2C-channel model output (:519 -> :360).
        if self.model_var_type in (ModelVarType.LEARNED, ModelVarType.LEARNED_RANGE):
            raise AssertionError("scg_sample needs a fixed-variance diffusion (learn_sigma=False): the reference's "
                                 "scg_sample asserts x_t.shape == eps.shape on a learn_sigma model's output")
        N = int(scg_kwargs["num_samples"])
        B = mean_pred.shape[0]
        dev = mean_pred.device
        elems = mean_pred[0].numel()
        stream = _lib.stream_ptr()
        mean_c = mean_pred.contiguous().float()
        g = g_coeff.reshape(B, -1)[:, 0].contiguous().float()  # fixed variances are one value per sample by construction
        noise = th.randn(N, *mean_pred.shape, device=dev, dtype=th.float32)  # same stream as randn_like(sample)
        shard = dist_util.candidate_sharding()
        n0, n1 = (0, N) if shard is None else dist_util.shard_range(N, shard[0], shard[1])
        Nl = n1 - n0  # candidates per sample on this rank
        if shard is not None:
            noise = noise[n0:n1].contiguous()  # every rank drew the same N rows; this one keeps rows n0..n1
        if Nl == 0:  # more ranks than candidates: nothing to offer, but take part in the exchange
            return self._select(None, None, 0, B, mean_c, n0, shard, record, t, None)
        cand = th.empty(Nl * B, *mean_pred.shape[1:], device=dev, dtype=th.float32)
        _lib.call("rgm_scg_fanout", _lib.ptr(mean_c), _lib.ptr(g), _lib.ptr(noise), _lib.ptr(cand), Nl, B, elems, stream)
        del noise
        t_rep = t.repeat(Nl)
        eps = model(cand, self._scale_timesteps(t_rep), y=model_kwargs["y"].repeat(Nl))
        assert eps.shape == cand.shape, "scg_sample: the denoiser must return eps of the input's shape (reference :360)"
        tab = self._tables(dev)
        a = tab["sqrt_recip_alphas_cumprod"][t_rep].contiguous()
        c = tab["sqrt_recipm1_alphas_cumprod"][t_rep].contiguous()
        x0 = th.empty_like(cand)
        _lib.call("rgm_x0_from_eps", _lib.ptr(cand), _lib.ptr(eps.contiguous()), _lib.ptr(a), _lib.ptr(c), _lib.ptr(x0),
                  Nl * B, elems, 0, stream)
        del eps
        if edit_kwargs is not None:
            x0 = x0[:, :, edit_kwargs["l_start"]:edit_kwargs["l_end"], :].contiguous()
        if embed_model is not None:
            native_rules = all(self._rule_is_native(n) for n in model_kwargs["rule"])
            if hasattr(embed_model, "decode_latents") and native_rules:
                roll = embed_model.decode_latents(x0, scale_factor, channels=1)  # the rules read channel 0 only
            else:
                roll = _decode(x0, embed_model, scale_factor=scale_factor)
        else:
            roll = x0
        del x0

        cand_v = cand.view(Nl, B, *mean_pred.shape[1:])
        if dc_kwargs is None or getattr(dc_kwargs, "base", 0) <= 0:
            total, each = self._score_candidates(roll, model_kwargs, scg_kwargs, Nl, B)
            return self._select(total, cand, Nl, B, mean_c, n0, shard, record, t, each)
        # per-segment selection for long sequences (reference :562-592)
        base = int(dc_kwargs.base)
        total_length = roll.shape[-1]
        rule_base = base // 16
        subs = []
        for i, start in enumerate(range(0, total_length, base * 8)):
            end = min(start + base * 8, total_length)
            roll_cur = roll[:, :, :, start:end].contiguous()

            def seg(rule_name, target, i=i):
                if rule_name == "note_density":
                    half = target.shape[-1] // 2
                    lo, hi = i * rule_base, min((i + 1) * rule_base, half)
                    return th.concat((target[:, :half][:, lo:hi], target[:, half:][:, lo:hi]), dim=-1)
                if "chord" in rule_name:
                    return target[:, i * rule_base: min((i + 1) * rule_base, target.shape[-1])]
                return target

            total, _ = self._score_candidates(roll_cur, model_kwargs, scg_kwargs, Nl, B, seg=seg)
            seg_cand = cand_v[:, :, :, start // 8: end // 8].contiguous()
            like = mean_c[:, :, start // 8: end // 8]
            subs.append(self._select(total, seg_cand, Nl, B, like, n0, shard, False, t, None))
        return th.concat(subs, dim=-2)

    def _select(self, total, cand, Nl, B, like, n0, shard, record, t, each):
        """First-max argmax over the candidates and gather of the winners (reference :539-554): the device kernel over
        this rank's Nl candidates, then -- when the candidates are sharded -- the exchange across ranks.
        `like` gives the shape of one winner per sample ([B, C, h, W])."""
        dev = like.device
        elems = like[0].numel()
        idx = th.zeros(B, device=dev, dtype=th.int64)
        sample = th.zeros(B, *like.shape[1:], device=dev, dtype=th.float32)
        if Nl > 0:
            _lib.call("rgm_scg_select", _lib.ptr(total), _lib.ptr(cand), _lib.ptr(sample), _lib.ptr(idx), Nl, B, elems,
                      _lib.stream_ptr())
        if shard is not None:
            if Nl > 0:
                best = total.view(Nl, B).gather(0, idx.view(1, B)).view(B)
            else:
                best = th.full((B,), float("-inf"), device=dev)
            sample, gidx = dist_util.first_max_over_ranks(best, idx + n0, sample, group=shard[2])
        else:
            gidx = idx
        if self._trace is not None:
            self._trace.append((None if total is None else total.view(Nl, B).clone(), gidx.clone()))
        if record and shard is None:
            self._record(t, total, idx, each, Nl, B)
        return sample

    @staticmethod
    def _rule_is_native(name):
        f = FUNC_DICT.get(name)
        f = getattr(f, "func", f)
        return f in (_mr.total_pitch_class_histogram, _mr.note_density, _mr.note_density_class)

    def _record(self, t, total, idx, each, N, B):
        """record=True bookkeeping (reference :594-632, without the matplotlib output): synchronises."""
        t0 = int(t[0])
        tot = total.view(N, B)
        self.log_probs.append((t0, tot[idx, th.arange(B, device=tot.device)][0].item()))
        for name, (gen, tgt, loss_fn) in each.items():
            loss = loss_fn(gen, tgt.to(gen.device).repeat(N, 1)).view(N, B)
            self.each_loss.setdefault(name, []).append((t0, loss[idx, th.arange(B, device=loss.device)][0].item()))

    # ---- whole-step CUDA graphs ------------------------------------------------------------------------------------
    def enable_cuda_graphs(self, on=True):
        """Capture each distinct kind of step (sampler, guidance on/off, last step, shapes, rule set) into a CUDA graph
        the second time it occurs and replay it afterwards: one graph launch instead of ~3000 kernel launches and the
        Python between them.  The timestep is a device tensor and every schedule coefficient is gathered on the device,
        so one graph serves all timesteps of its kind.  Steps that call back into user Python (cond_fn, denoised_fn,
        rules or losses the kernels do not know, record=True, the SCG trace hook) always run eagerly.  Results are
        bit-identical to eager execution: the same kernels run in the same order and torch's Philox generator advances
        by the same offsets."""
        self._graphs_on = bool(on)
        if not on:
            self._graphs = collections.OrderedDict()
        return self

    def _graphable(self, x, kw):
        if not x.is_cuda or kw["cond_fn"] is not None or kw["denoised_fn"] is not None or kw["record"]:
            return False
        if self._trace is not None:
            return False
        mk = kw["model_kwargs"] or {}
        if kw["scg_kwargs"] is not None:
            sh = dist_util.candidate_sharding()
            if sh is not None and dist_util.dist.get_backend(sh[2]) != "nccl":
                return False  # the gloo exchange stages through the host
            rules = mk.get("rule", {})
            if not all(self._rule_is_native(n) and LOSS_DICT.get(n) in NATIVE_LOSS_KIND for n in rules):
                return False
            # a target on the host would be copied in every step: a pageable copy is illegal inside a capture
            if not all(isinstance(v, th.Tensor) and v.is_cuda for v in rules.values()):
                return False
            em = kw["embed_model"]
            if em is not None and not hasattr(em, "decode_latents"):
                return False
        return True

    @staticmethod
    def _sig(v, refs):
        """Hashable signature of a step argument.  Tensors by storage identity (a graph bakes their addresses in);
        mappings, namespaces and functools.partial objects by CONTENT (a new partial with another cfg weight, or a
        config object mutated in place, is a different step); everything else by identity, with a strong reference
        appended to `refs` so that the id cannot be recycled while the cache entry lives."""
        sig = GaussianDiffusion._sig
        if isinstance(v, th.Tensor):
            refs.append(v)
            return ("T", v.data_ptr(), tuple(v.shape), str(v.dtype))
        if isinstance(v, (int, float, str, bool)) or v is None:
            return v
        if isinstance(v, collections.abc.Mapping):
            return ("M",) + tuple((str(k), sig(x, refs)) for k, x in v.items())
        if isinstance(v, types.SimpleNamespace):
            return ("N",) + tuple((k, sig(x, refs)) for k, x in sorted(vars(v).items()))
        if isinstance(v, functools.partial):
            return ("P", sig(v.func, refs), tuple(sig(a, refs) for a in v.args),
                    tuple((k, sig(x, refs)) for k, x in sorted(v.keywords.items())))
        if isinstance(v, (tuple, list)):
            return ("L",) + tuple(sig(x, refs) for x in v)
        refs.append(v)
        return ("O", id(v))  # models, decoders, plain functions: by identity

    MAX_STEP_GRAPHS = 8  # captured graphs kept (each pins its own memory pool); least recently used is released

    def _graphed_step(self, name, eager, model, x, t, t0, kw):
        # of guidance_kwargs a step reads the on/off decision (host-side, from t0) and the per-segment base length
        g = kw["guidance_kwargs"]
        dc_base = getattr(getattr(g, "dc", None), "base", 0)
        sh = dist_util.candidate_sharding()
        refs = []
        key = (name, self._sig(model, refs), tuple(x.shape), str(x.device), self._use_guidance(t0, g), dc_base,
               None if sh is None else sh[:2], t0 > self.t_end, self.t_end,
               tuple((k, self._sig(v, refs)) for k, v in kw.items() if k != "guidance_kwargs"))
        graphs = self._graphs
        ent = graphs.get(key)
        if ent is None:
            # first occurrence: eager (it also grows the library's workspaces; growing under capture is refused).  The
            # placeholder keeps the references too, so the key cannot be matched by recycled ids.
            graphs[key] = _StepGraph(refs)
            self._evict_graphs()
            return eager(model, x, t, t0, **kw)
        graphs.move_to_end(key)
        if ent.graph is None:
            ent.x = x.clone()
            ent.t = t.clone()
            th.cuda.synchronize(x.device)
            graph = th.cuda.CUDAGraph()
            with th.cuda.graph(graph):
                ent.out = eager(model, ent.x, ent.t, t0, **kw)
            ent.graph = graph
            self._evict_graphs()
        ent.x.copy_(x)
        ent.t.copy_(t)
        ent.graph.replay()
        return {k: v.clone() for k, v in ent.out.items()}

    def captured_graphs(self):
        """Number of step kinds currently held as captured CUDA graphs."""
        return sum(1 for e in self._graphs.values() if e.graph is not None)

    def _evict_graphs(self):
        """Bound the cache: at most MAX_STEP_GRAPHS captured graphs and 64 entries in total (a caller that hands in
        freshly allocated tensors every step never repeats a signature)."""
        graphs = self._graphs
        captured = [k for k, e in graphs.items() if e.graph is not None]
        while len(captured) > self.MAX_STEP_GRAPHS:
            graphs.pop(captured.pop(0))  # dropping the entry releases the CUDAGraph and its private pool
        while len(graphs) > 64:
            graphs.popitem(last=False)

    # ---- one ancestral step (reference :635-735) -----------------------------------------------------------------
    @staticmethod
    def _use_guidance(t0, guidance_kwargs):
        if guidance_kwargs is None:
            return False
        if guidance_kwargs.schedule:
            return guide_schedule(t0, guidance_kwargs.t_start, guidance_kwargs.t_end, guidance_kwargs.interval)
        return True

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                 embed_model=None, scale_factor=1., guidance_kwargs=None, scg_kwargs=None, edit_kwargs=None,
                 record=False, _t_host=None):
        t0 = int(t[0]) if _t_host is None else _t_host  # the loops pass the index they built t from: no sync
        kw = dict(clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs,
                  embed_model=embed_model, scale_factor=scale_factor, guidance_kwargs=guidance_kwargs,
                  scg_kwargs=scg_kwargs, edit_kwargs=edit_kwargs, record=record)
        if self._graphs_on and self._graphable(x, kw):
            return self._graphed_step("p_sample", self._p_sample_eager, model, x, t, t0, kw)
        return self._p_sample_eager(model, x, t, t0, **kw)

    def _p_sample_eager(self, model, x, t, t0, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                        embed_model=None, scale_factor=1., guidance_kwargs=None, scg_kwargs=None, edit_kwargs=None,
                        record=False):
        use_guidance = self._use_guidance(t0, guidance_kwargs)
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs, cond_fn=cond_fn, embed_model=embed_model,
                                   edit_kwargs=edit_kwargs)
        if cond_fn is not None and (use_guidance or scg_kwargs is not None):
            out["mean"] = self.condition_mean(cond_fn, out, x, t, model_kwargs=model_kwargs,
                                              guidance_kwargs=guidance_kwargs, model=model, embed_model=embed_model,
                                              edit_kwargs=edit_kwargs, scale_factor=scale_factor)
        if scg_kwargs is None:
            noise = th.randn_like(x)
            nonzero_mask = (t > self.t_end).float().view(-1, *([1] * (x.dim() - 1)))
            sample = out["mean"] + nonzero_mask * th.exp(0.5 * out["log_variance"]) * noise
        elif t0 > self.t_end:
            g_coeff = th.exp(0.5 * out["log_variance"])
            if use_guidance:
                sample = self.scg_sample(model, t, out["mean"], g_coeff, embed_model, scale_factor,
                                         model_kwargs=model_kwargs, scg_kwargs=scg_kwargs, edit_kwargs=edit_kwargs,
                                         dc_kwargs=getattr(guidance_kwargs, "dc", None), record=record)
            else:
                sample = out["mean"] + g_coeff * th.randn_like(x)
        else:
            sample = out["mean"]
        return {"sample": sample, "pred_xstart": out["pred_xstart"]}

    # ---- one DDIM step (reference :881-976) -------------------------------------------------------------------------
    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0,
                    embed_model=None, scale_factor=1., guidance_kwargs=None, scg_kwargs=None, edit_kwargs=None,
                    record=False, _t_host=None):
        t0 = int(t[0]) if _t_host is None else _t_host
        kw = dict(clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn, model_kwargs=model_kwargs,
                  eta=eta, embed_model=embed_model, scale_factor=scale_factor, guidance_kwargs=guidance_kwargs,
                  scg_kwargs=scg_kwargs, edit_kwargs=edit_kwargs, record=record)
        if self._graphs_on and self._graphable(x, kw):
            return self._graphed_step("ddim_sample", self._ddim_sample_eager, model, x, t, t0, kw)
        return self._ddim_sample_eager(model, x, t, t0, **kw)

    def _ddim_sample_eager(self, model, x, t, t0, clip_denoised=True, denoised_fn=None, cond_fn=None,
                           model_kwargs=None, eta=0.0, embed_model=None, scale_factor=1., guidance_kwargs=None,
                           scg_kwargs=None, edit_kwargs=None, record=False):
        use_guidance = self._use_guidance(t0, guidance_kwargs)
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, denoised_fn=denoised_fn,
                                   model_kwargs=model_kwargs, cond_fn=cond_fn, embed_model=embed_model,
                                   edit_kwargs=edit_kwargs)
        if cond_fn is not None and use_guidance:
            out = self.condition_score(cond_fn, out, x, t, model_kwargs=model_kwargs)
        eps = self._predict_eps_from_xstart(x, t, out["pred_xstart"])
        n = x.dim()
        alpha_bar = self._coef("alphas_cumprod", t, n)
        alpha_bar_prev = self._coef("alphas_cumprod_prev", t, n)
        sigma = eta * th.sqrt((1 - alpha_bar_prev) / (1 - alpha_bar)) * th.sqrt(1 - alpha_bar / alpha_bar_prev)
        mean_pred = out["pred_xstart"] * th.sqrt(alpha_bar_prev) + th.sqrt(1 - alpha_bar_prev - sigma ** 2) * eps
        if scg_kwargs is None:
            nonzero_mask = (t != self.t_end).float().view(-1, *([1] * (n - 1)))
            sample = mean_pred + nonzero_mask * sigma * th.randn_like(x)
        elif t0 > self.t_end:
            g_coeff = sigma.expand(x.shape)
            if use_guidance:
                sample = self.scg_sample(self._wrap_model(model), t, mean_pred, g_coeff, embed_model, scale_factor,
                                         model_kwargs=model_kwargs, scg_kwargs=scg_kwargs, edit_kwargs=edit_kwargs,
                                         dc_kwargs=getattr(guidance_kwargs, "dc", None), record=record, record_freq=10)
            else:
                sample = mean_pred + g_coeff * th.randn_like(x)
        else:
            sample = mean_pred
        return {"sample": sample, "pred_xstart": out["pred_xstart"]}

    # ---- loops (reference :737-879, :1016-1143) ---------------------------------------------------------------------
    def _loop(self, step, model, shape, noise, t_end, device, progress, edit_kwargs, **kw):
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        if noise is not None:
            img = noise
        elif edit_kwargs is not None:
            t = th.full((shape[0],), edit_kwargs["noise_level"] - 1, device=device, dtype=th.long)
            ac = self._coef("alphas_cumprod", t, len(shape))
            img = th.sqrt(ac) * edit_kwargs["gt"] + th.sqrt(1 - ac) * th.randn(*shape, device=device)
        else:
            img = th.randn(*shape, device=device)
        indices = list(range(self.num_timesteps))[::-1]
        if t_end:
            indices = indices[:-t_end]
        if edit_kwargs is not None:
            indices = indices[self.num_timesteps - edit_kwargs["noise_level"]:]
        if progress:
            from tqdm.auto import tqdm
            indices = tqdm(indices)
        for i in indices:
            t = th.full((shape[0],), i, device=device, dtype=th.long)
            with th.no_grad():
                out = step(model, img, t, edit_kwargs=edit_kwargs, _t_host=i, **kw)
            yield out
            img = out["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, t_end=0,
                                  cond_fn=None, model_kwargs=None, device=None, progress=False, embed_model=None,
                                  scale_factor=1., guidance_kwargs=None, scg_kwargs=None, edit_kwargs=None,
                                  record=False):
        yield from self._loop(self.p_sample, model, shape, noise, t_end, device, progress, edit_kwargs,
                              clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                              model_kwargs=model_kwargs, embed_model=embed_model, scale_factor=scale_factor,
                              guidance_kwargs=guidance_kwargs, scg_kwargs=scg_kwargs, record=record)

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, t_end=0, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, embed_model=None, scale_factor=1.,
                      guidance_kwargs=None, scg_kwargs=None, edit_kwargs=None, record=False):
        self.t_end = t_end
        if record:
            self.log_probs, self.each_loss = [], {}
        final = None
        for sample in self.p_sample_loop_progressive(
                model, shape, noise=noise, clip_denoised=clip_denoised, denoised_fn=denoised_fn, t_end=t_end,
                cond_fn=cond_fn, model_kwargs=model_kwargs, device=device, progress=progress, embed_model=embed_model,
                scale_factor=scale_factor, guidance_kwargs=guidance_kwargs, scg_kwargs=scg_kwargs,
                edit_kwargs=edit_kwargs, record=record):
            final = sample
        return final["sample"]

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, t_end=0,
                                     cond_fn=None, model_kwargs=None, device=None, progress=False, eta=0.0,
                                     embed_model=None, scale_factor=1., guidance_kwargs=None, scg_kwargs=None,
                                     edit_kwargs=None, record=False):
        yield from self._loop(self.ddim_sample, model, shape, noise, t_end, device, progress, edit_kwargs,
                              clip_denoised=clip_denoised, denoised_fn=denoised_fn, cond_fn=cond_fn,
                              model_kwargs=model_kwargs, eta=eta, embed_model=embed_model, scale_factor=scale_factor,
                              guidance_kwargs=guidance_kwargs, scg_kwargs=scg_kwargs, record=record)

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, t_end=0, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, embed_model=None, scale_factor=1.,
                         guidance_kwargs=None, scg_kwargs=None, edit_kwargs=None, record=False):
        self.t_end = t_end
        if record:
            self.log_probs, self.each_loss = [], {}
        final = None
        for sample in self.ddim_sample_loop_progressive(
                model, shape, noise=noise, clip_denoised=clip_denoised, denoised_fn=denoised_fn, t_end=t_end,
                cond_fn=cond_fn, model_kwargs=model_kwargs, device=device, progress=progress, eta=eta,
                embed_model=embed_model, scale_factor=scale_factor, guidance_kwargs=guidance_kwargs,
                scg_kwargs=scg_kwargs, edit_kwargs=edit_kwargs, record=record):
            final = sample
        return final["sample"]

    # ---- not on the sampling path --------------------------------------------------------------------------------
    def training_losses(self, *a, **k):
        raise NotImplementedError("training is out of scope of the B200 sampling path")

    calc_bpd_loop = ddim_reverse_sample = training_losses
