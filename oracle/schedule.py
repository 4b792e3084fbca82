"""Diffusion schedules and timestep respacing (float64 numpy, exactly as the reference).

Follows guided_diffusion/gaussian_diffusion.py:31-62 (beta schedules), :138-189 (derived tables), :316-329
(FIXED_LARGE variance) and guided_diffusion/respace.py:7-60 (space_timesteps), :63-86 (re-derived betas).
"""
import math

import numpy as np


def named_beta_schedule(name, num_steps):
    """gaussian_diffusion.py:31-62."""
    if name == "linear":
        scale = 1000 / num_steps
        return np.linspace(scale * 0.0001, scale * 0.02, num_steps, dtype=np.float64)
    if name == "cosine":
        f = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        betas = []
        for i in range(num_steps):
            t1, t2 = i / num_steps, (i + 1) / num_steps
            betas.append(min(1 - f(t2) / f(t1), 0.999))
        return np.array(betas)
    if name == "stable-diffusion":
        scale = 1000 / num_steps
        return np.linspace(scale * math.sqrt(0.00085), scale * math.sqrt(0.012), num_steps, dtype=np.float64) ** 2
    raise NotImplementedError(name)


def diffusion_tables(betas):
    """All per-timestep tables GaussianDiffusion.__init__ derives (gaussian_diffusion.py:152-189) plus the
    FIXED_LARGE variance pair used at sampling time (:316-329).  Everything float64."""
    betas = np.array(betas, dtype=np.float64)
    assert betas.ndim == 1 and (betas > 0).all() and (betas <= 1).all()
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    ac_next = np.append(ac[1:], 0.0)
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    t = {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "alphas_cumprod_next": ac_next,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": np.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": np.log(np.append(post_var[1], post_var[1:])),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }
    fixed_large = np.append(post_var[1], betas[1:])
    t["fixed_large_variance"] = fixed_large
    t["fixed_large_log_variance"] = np.log(fixed_large)
    return t


def space_timesteps(num_timesteps, section_counts):
    """respace.py:7-60."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == desired:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start = 0
    steps = []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        frac = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            steps.append(start + round(cur))
            cur += frac
        start += size
    return set(steps)


def spaced_betas(base_betas, use_timesteps):
    """respace.py:72-86: betas of the sub-sampled process and the map spaced index -> original timestep."""
    ac = np.cumprod(1.0 - np.array(base_betas, dtype=np.float64))
    use = set(use_timesteps)
    last = 1.0
    new_betas, tmap = [], []
    for i, a in enumerate(ac):
        if i in use:
            new_betas.append(1 - a / last)
            last = a
            tmap.append(i)
    return np.array(new_betas), tmap


def guide_schedule(t0, t_start=750, t_end=0, interval=1):
    """gaussian_diffusion.py:1398-1400 (t0 = t[0])."""
    return bool(t_start > t0 >= t_end and (t0 + 1) % interval == 0)
