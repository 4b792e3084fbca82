"""Sampler restatement: p_mean_variance / p_sample / ddim_sample / scg_sample and the two loops, torch fp32 on CPU.

Follows guided_diffusion/gaussian_diffusion.py: _extract_into_tensor :1331-1344 (float64 table -> index -> .float()),
p_mean_variance :252-357, _predict_xstart_from_eps :359-364, q_posterior_mean_variance :228-250,
_predict_eps_from_xstart :376-380, condition_mean :402-407 (classifier-guidance branch), condition_score :467-489,
scg_sample :491-554, p_sample :635-735, p_sample_loop_progressive :809-879, ddim_sample :881-976,
ddim_sample_loop_progressive :1073-1143, _extract_rule :1361-1379; and guided_diffusion/respace.py:63-128
(SpacedDiffusion / _WrappedModel) plus script_util.py:462-500 (EPSILON, FIXED_LARGE unless learn_sigma).

Reference quirks kept on purpose: SCG candidates are scored at timestep t, not t-1; p_sample
hands the UNWRAPPED model to scg_sample (:711) while ddim_sample wraps it (:954); SCG needs "y" in model_kwargs;
argmax takes the first maximal candidate; DDPM masks noise with t > t_end, DDIM with t != t_end.
Not restated: edit_kwargs, DPS (autograd), record/plot branches, per-segment dc selection.
"""
import numpy as np
import torch

from . import rules as orules
from . import schedule as osched


def _extract(arr, t, shape):
    res = torch.from_numpy(np.asarray(arr, dtype=np.float64)).to(t.device)[t].float()
    while res.dim() < len(shape):
        res = res[..., None]
    return res.expand(shape)


class _Wrapped:
    def __init__(self, model, tmap, rescale, orig_steps):
        self.model, self.tmap, self.rescale, self.orig = model, tmap, rescale, orig_steps

    def __call__(self, x, ts, **kw):
        new_ts = torch.tensor(self.tmap, dtype=ts.dtype, device=ts.device)[ts]
        if self.rescale:
            new_ts = new_ts.float() * (1000.0 / self.orig)
        return self.model(x, new_ts, **kw)


class OracleDiffusion:
    def __init__(self, steps=1000, noise_schedule="linear", timestep_respacing="", learn_sigma=False,
                 rescale_timesteps=False, randn=None):
        base = osched.named_beta_schedule(noise_schedule, steps)
        if not timestep_respacing:
            timestep_respacing = [steps]
        use = osched.space_timesteps(steps, timestep_respacing)
        betas, self.timestep_map = osched.spaced_betas(base, use)
        self.original_num_steps = steps
        self.tab = osched.diffusion_tables(betas)
        self.num_timesteps = len(betas)
        self.learn_sigma = learn_sigma
        self.rescale_timesteps = rescale_timesteps
        self.t_end = 0
        self._randn = randn or (lambda shape: torch.randn(*shape))

    def randn_like(self, x):
        return self._randn(tuple(x.shape))

    def wrap(self, model):
        if isinstance(model, _Wrapped):
            return model
        return _Wrapped(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    # ---- :359-380, :228-250
    def predict_xstart_from_eps(self, x_t, t, eps):
        return (_extract(self.tab["sqrt_recip_alphas_cumprod"], t, x_t.shape) * x_t
                - _extract(self.tab["sqrt_recipm1_alphas_cumprod"], t, x_t.shape) * eps)

    def predict_eps_from_xstart(self, x_t, t, x0):
        return ((_extract(self.tab["sqrt_recip_alphas_cumprod"], t, x_t.shape) * x_t - x0)
                / _extract(self.tab["sqrt_recipm1_alphas_cumprod"], t, x_t.shape))

    def q_posterior_mean(self, x0, x_t, t):
        return (_extract(self.tab["posterior_mean_coef1"], t, x_t.shape) * x0
                + _extract(self.tab["posterior_mean_coef2"], t, x_t.shape) * x_t)

    # ---- :252-357 (SpacedDiffusion wraps the model first, respace.py:88-91)
    def p_mean_variance(self, model, x, t, clip_denoised=True, model_kwargs=None):
        model_kwargs = model_kwargs or {}
        out = self.wrap(model)(x, t, **model_kwargs)
        C = x.shape[1]
        if self.learn_sigma:
            out, var_values = torch.split(out, C, dim=1)
            min_log = _extract(self.tab["posterior_log_variance_clipped"], t, x.shape)
            max_log = _extract(np.log(self.tab["betas"]), t, x.shape)
            frac = (var_values + 1) / 2
            log_var = frac * max_log + (1 - frac) * min_log
            var = torch.exp(log_var)
        else:
            var = _extract(self.tab["fixed_large_variance"], t, x.shape)
            log_var = _extract(self.tab["fixed_large_log_variance"], t, x.shape)
        x0 = self.predict_xstart_from_eps(x, t, out)
        if clip_denoised:
            x0 = x0.clamp(-1, 1)
        return {"mean": self.q_posterior_mean(x0, x, t), "variance": var, "log_variance": log_var, "pred_xstart": x0}

    # ---- :491-554
    def scg_sample(self, model, t, mean_pred, g_coeff, decode_fn, model_kwargs, scg_kwargs, trace=None):
        n = scg_kwargs["num_samples"]
        sample = mean_pred.unsqueeze(0).expand(n, *mean_pred.shape).contiguous()
        noise = self.randn_like(sample)
        sample = (sample + g_coeff * noise).view(-1, *mean_pred.shape[1:])
        tt = t.repeat(n)
        eps = model(sample, tt, y=model_kwargs["y"].repeat(n))
        x0 = self.predict_xstart_from_eps(sample, tt, eps)
        if decode_fn is not None:
            x0 = decode_fn(x0)
        total = 0
        for name, target in model_kwargs["rule"].items():
            gen = orules.FUNC_DICT[name](x0)
            log_prob = -orules.LOSS_DICT[name](gen, target.repeat(n, 1))
            total = total + log_prob * scg_kwargs.get(name, 1.0)
        total = total.view(n, -1)
        max_ind = total.argmax(dim=0)
        sample = sample.view(n, *mean_pred.shape)
        chosen = sample[max_ind, torch.arange(mean_pred.shape[0], device=max_ind.device)]
        if trace is not None:
            trace.append({"candidates": sample.clone(), "eps": eps.clone(), "roll": x0.clone(),
                          "total_log_prob": total.clone(), "max_ind": max_ind.clone()})
        return chosen

    def _use_guidance(self, t, guidance_kwargs):
        if guidance_kwargs is None:
            return False
        if guidance_kwargs.schedule:
            return osched.guide_schedule(int(t[0]), guidance_kwargs.t_start, guidance_kwargs.t_end,
                                         guidance_kwargs.interval)
        return True

    # ---- :635-735
    def p_sample(self, model, x, t, clip_denoised=True, cond_fn=None, model_kwargs=None, decode_fn=None,
                 guidance_kwargs=None, scg_kwargs=None, trace=None):
        use_guidance = self._use_guidance(t, guidance_kwargs)
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, model_kwargs=model_kwargs)
        if cond_fn is not None and (use_guidance or scg_kwargs is not None):
            grad = self.wrap(cond_fn)(x, t, **model_kwargs)
            out["mean"] = out["mean"].float() + out["variance"] * grad.float()
        if scg_kwargs is None:
            noise = self.randn_like(x)
            mask = (t > self.t_end).float().view(-1, *([1] * (x.dim() - 1)))
            sample = out["mean"] + mask * torch.exp(0.5 * out["log_variance"]) * noise
        elif int(t[0]) > self.t_end:
            g = torch.exp(0.5 * out["log_variance"])
            if use_guidance:
                sample = self.scg_sample(model, t, out["mean"], g, decode_fn, model_kwargs, scg_kwargs, trace)
            else:
                sample = out["mean"] + g * self.randn_like(x)
        else:
            sample = out["mean"]
        return {"sample": sample, "pred_xstart": out["pred_xstart"], "mean": out["mean"]}

    # ---- :881-976
    def ddim_sample(self, model, x, t, clip_denoised=True, cond_fn=None, model_kwargs=None, eta=0.0, decode_fn=None,
                    guidance_kwargs=None, scg_kwargs=None, trace=None):
        use_guidance = self._use_guidance(t, guidance_kwargs)
        out = self.p_mean_variance(model, x, t, clip_denoised=clip_denoised, model_kwargs=model_kwargs)
        if cond_fn is not None and use_guidance:
            ab = _extract(self.tab["alphas_cumprod"], t, x.shape)
            eps = self.predict_eps_from_xstart(x, t, out["pred_xstart"])
            eps = eps - (1 - ab).sqrt() * self.wrap(cond_fn)(x, t, **model_kwargs)
            out["pred_xstart"] = self.predict_xstart_from_eps(x, t, eps)
            out["mean"] = self.q_posterior_mean(out["pred_xstart"], x, t)
        eps = self.predict_eps_from_xstart(x, t, out["pred_xstart"])
        ab = _extract(self.tab["alphas_cumprod"], t, x.shape)
        ab_prev = _extract(self.tab["alphas_cumprod_prev"], t, x.shape)
        sigma = eta * torch.sqrt((1 - ab_prev) / (1 - ab)) * torch.sqrt(1 - ab / ab_prev)
        mean_pred = out["pred_xstart"] * torch.sqrt(ab_prev) + torch.sqrt(1 - ab_prev - sigma ** 2) * eps
        if scg_kwargs is None:
            mask = (t != self.t_end).float().view(-1, *([1] * (x.dim() - 1)))
            sample = mean_pred + mask * sigma * self.randn_like(x)
        elif int(t[0]) > self.t_end:
            if use_guidance:
                sample = self.scg_sample(self.wrap(model), t, mean_pred, sigma, decode_fn, model_kwargs, scg_kwargs,
                                         trace)
            else:
                sample = mean_pred + sigma * self.randn_like(x)
        else:
            sample = mean_pred
        return {"sample": sample, "pred_xstart": out["pred_xstart"], "mean": mean_pred}

    # ---- :809-879 / :1073-1143
    def _loop(self, step_fn, model, shape, noise, t_end, **kw):
        self.t_end = t_end
        img = noise if noise is not None else self._randn(tuple(shape))
        indices = list(range(self.num_timesteps))[::-1]
        if t_end:
            indices = indices[:-t_end]
        for i in indices:
            t = torch.tensor([i] * shape[0])
            with torch.no_grad():
                out = step_fn(model, img, t, **kw)
            yield out
            img = out["sample"]

    def p_sample_loop(self, model, shape, noise=None, t_end=0, **kw):
        final = None
        for final in self._loop(self.p_sample, model, shape, noise, t_end, **kw):
            pass
        return final["sample"]

    def ddim_sample_loop(self, model, shape, noise=None, t_end=0, **kw):
        final = None
        for final in self._loop(self.ddim_sample, model, shape, noise, t_end, **kw):
            pass
        return final["sample"]
