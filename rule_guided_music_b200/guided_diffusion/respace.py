"""Timestep respacing -- mirror of guided_diffusion/respace.py:7-128 of the reference.

space_timesteps :7-60, SpacedDiffusion :63-113 (betas of the kept steps re-derived from the cumulative products,
model / cond_fn wrapped so they see ORIGINAL timestep numbers), _WrappedModel :116-128.  The one difference: the
spaced->original map lives on the device once instead of being rebuilt from a Python list on every call (:124).
"""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Which original timesteps a shortened process keeps.  "ddimN": the integer stride that yields exactly N steps
    (ValueError when none does); "a,b,c" or a list: split the range into equal sections and take a, b, c evenly
    spaced steps (rounded) from each."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    kept, start = [], 0
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += stride
        start += size
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """A diffusion process that skips steps of a base process (respace.py:63-113)."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(kwargs["betas"])
        base = GaussianDiffusion(**kwargs)
        last = 1.0
        new_betas = []
        for i, ac in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - ac / last)
                last = ac
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)
        # device copies of timestep_map, shared by every wrapper this diffusion hands out: a wrapper is built per call,
        # and a fresh host->device copy per step would be a pageable transfer (illegal inside a graph capture)
        self._dev_maps = {}

    def p_mean_variance(self, model, *args, **kwargs):
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def condition_mean(self, cond_fn, *args, **kwargs):
        return super().condition_mean(self._wrap_model(cond_fn), *args, **kwargs)

    def condition_score(self, cond_fn, *args, **kwargs):
        return super().condition_score(self._wrap_model(cond_fn), *args, **kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps,
                             maps=self._dev_maps)

    def _scale_timesteps(self, t):
        return t  # scaling is done by the wrapped model (respace.py:111-113)


class _WrappedModel:
    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps, maps=None):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self._maps = {} if maps is None else maps

    def __call__(self, x, ts, **kwargs):
        key = (str(ts.device), ts.dtype)
        m = self._maps.get(key)
        if m is None:
            m = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
            self._maps[key] = m
        new_ts = m[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)

    def parameters(self):
        return self.model.parameters()
