"""Seeded inputs shared by the golden-vector generator (tests/golden/make_golden.py, runs the unmodified reference),
the CPU oracle tests and the GPU parity tests.  Only reference OUTPUTS are stored under tests/golden/; inputs are
regenerated here from seeds with torch's CPU generator, which is identical in the build container and on the GPU box.
"""
import torch

SCALE_FACTOR = 1.2465  # reference README.md:59
VAE_SEED = 1

SCHEDULE_KEYS = [
    "betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_alphas_cumprod",
    "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
    "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
    "posterior_mean_coef2",
]

RULE_NAMES = ["pitch_hist", "note_density", "note_density_hr_1", "note_density_hr_2", "note_density_class",
              "note_density_pixel"]

RULE_QUANT = [2, 4]  # note_density(quantize_factor=...) cases

# ---------------------------------------------------------------------------------------------------------------
# DiT
# ---------------------------------------------------------------------------------------------------------------
DIT_CASES = {
    # shallow model with the XL width and head geometry (head_dim 72, rotary 36): every kernel path, quickly
    "small": dict(preset=None, input_size=[128, 16], batch=3, half_tile=True,
                  weights=dict(seed=11, depth=2, hidden=1152, patch=8, heads=16, num_classes=3)),
    # head_dim 64 (DiTRotary_B geometry), patch 16 (one token per time step)
    "small_hd64": dict(preset=None, input_size=[128, 16], batch=2, half_tile=False,
                       weights=dict(seed=12, depth=2, hidden=384, patch=16, heads=6, num_classes=3)),
    # the flagship: DiTRotary_XL_8 (depth 28, hidden 1152, 16 heads), reference dit.py:902
    "xl8": dict(preset="DiTRotary_XL_8", input_size=[128, 16], batch=2, half_tile=True,
                weights=dict(seed=0, depth=28, hidden=1152, patch=8, heads=16, num_classes=3)),
}


# DiTRotaryClassifier (dit.py:735-831; `DiTRotary-XS/8-cls`, dit.py:951): the classifier-guidance model of SURVEY.md 8(f)
CLASSIFIER_CASE = dict(input_size=[128, 16], batch=2,
                       weights=dict(seed=21, depth=4, hidden=384, patch=8, heads=6, num_classes=9))


def classifier_inputs(cfg=CLASSIFIER_CASE):
    g = torch.Generator(device="cpu").manual_seed(2000 + cfg["weights"]["seed"])
    B = cfg["batch"]
    H, W = cfg["input_size"]
    x = torch.randn(B, 4, H, W, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    labels = torch.randint(0, cfg["weights"]["num_classes"], (B, 1), generator=g)
    return x, t, labels


def dit_inputs(cfg):
    g = torch.Generator(device="cpu").manual_seed(1000 + cfg["weights"]["seed"])
    B = cfg["batch"]
    H, W = cfg["input_size"]
    x = torch.randn(B, 4, H, W, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    y = torch.randint(0, cfg["weights"]["num_classes"], (B,), generator=g)
    return x, t, y


# ---------------------------------------------------------------------------------------------------------------
# rules
# ---------------------------------------------------------------------------------------------------------------
def rule_rolls():
    """Hand-built and random piano rolls [B,3,128,1024] in [-1,1]."""
    rolls = {}
    pr = -torch.ones(2, 3, 128, 1024)
    pr[0, 0, 60, 0:128] = 1
    pr[1, 0, 61, 100:300] = 0.5
    pr[1, 0, 73, 100:300] = 0.5
    pr[1, 0, 10, :] = 1
    pr[1, 0, 110, :] = 1
    pr[1, 0, 64, 512] = -0.95
    pr[1, 0, 65, 513] = -0.951
    rolls["kat"] = pr
    o = -torch.ones(2, 3, 128, 1024)
    o[0, 0, 60, :] = -1.2
    o[0, 0, 62, :10] = 1
    o[1, 0, 40, 5:900] = 0.3
    rolls["order"] = o
    g = torch.Generator(device="cpu").manual_seed(77)
    # sparse random notes: values near the -0.95 threshold, out-of-range pitches, values beyond [-1, 1]
    r = -torch.ones(4, 3, 128, 1024)
    mask = torch.rand(4, 128, 1024, generator=g) < 0.03
    vals = torch.rand(4, 128, 1024, generator=g) * 2.4 - 1.2
    r[:, 0] = torch.where(mask, vals, r[:, 0])
    near = torch.rand(4, 128, 1024, generator=g) < 0.01
    r[:, 0] = torch.where(near, -0.95 + (torch.rand(4, 128, 1024, generator=g) - 0.5) * 1e-3, r[:, 0])
    rolls["random"] = r
    one = -torch.ones(1, 3, 128, 1024)
    one[0, 0, 50:55, 200:600] = 0.7
    rolls["single"] = one  # B == 1: the reference squeezes the batch dimension
    return rolls


def loss_pairs():
    g = torch.Generator(device="cpu").manual_seed(5)
    return torch.rand(6, 16, generator=g) * 8, torch.rand(6, 16, generator=g) * 8


# ---------------------------------------------------------------------------------------------------------------
# VAE
# ---------------------------------------------------------------------------------------------------------------
def vae_tiles(n=2):
    g = torch.Generator(device="cpu").manual_seed(21)
    return torch.randn(n, 4, 16, 16, generator=g)


VAE_ENC_SEED = 2


def vae_rolls(B=2, L=256):
    """Piano-roll-like encoder input [B,3,128,L] in [-1,1]: silence (-1) with sparse held notes / onsets / pedal."""
    g = torch.Generator(device="cpu").manual_seed(23)
    r = -torch.ones(B, 3, 128, L)
    for b in range(B):
        for _ in range(40):
            p = int(torch.randint(21, 109, (1,), generator=g))
            t0 = int(torch.randint(0, L - 8, (1,), generator=g))
            d = int(torch.randint(2, 40, (1,), generator=g))
            v = float(torch.rand(1, generator=g)) * 1.6 - 0.6
            r[b, 0, p, t0:t0 + d] = v
            r[b, 1, p, t0] = v
        r[b, 2, :, 32 * b:32 * b + 64] = 0.3
    return r + 0.01 * torch.randn(B, 3, 128, L, generator=g)


def vae_latents(B=2, H=32):
    g = torch.Generator(device="cpu").manual_seed(22)
    return torch.randn(B, 4, H, 16, generator=g) * SCALE_FACTOR


# ---------------------------------------------------------------------------------------------------------------
# sampler trajectories (teacher data: every intermediate x_t of a short loop through the reference)
# ---------------------------------------------------------------------------------------------------------------
_GUIDE_ON = dict(schedule=False, t_start=750, t_end=0, interval=1, method="scg", step_size=1.0, nn=False)
_GUIDE_SCHED = dict(schedule=True, t_start=750, t_end=0, interval=2, method="scg", step_size=1.0, nn=False)

SAMPLER_CASES = {
    # plain ancestral sampling, 4 respaced steps, no guidance
    "ddpm_plain": dict(dit="small", respacing="4", ddim=False, shape=(2, 4, 128, 16), seed=3, scg=None,
                       guidance=None),
    # DDIM eta=0 / eta=1 without guidance
    "ddim_plain": dict(dit="small", respacing="ddim4", ddim=True, eta=0.0, shape=(2, 4, 128, 16), seed=4, scg=None,
                       guidance=None),
    # DDIM(eta=1) + SCG, pitch histogram, N=3 (config-3 shape in miniature); H=64 latents = 4 VAE tiles
    "ddim_scg_pitch": dict(dit="small", respacing="4", ddim=True, eta=1.0, shape=(2, 4, 64, 16), seed=5,
                           scg=dict(num_samples=3, pitch_hist=1.0), guidance=_GUIDE_ON, rules=["pitch_hist"]),
    # DDPM + SCG with two rules (order dependent in-place masking) and a scheduled guidance window
    "ddpm_scg_two": dict(dit="small", respacing="6", ddim=False, shape=(2, 4, 64, 16), seed=6,
                         scg=dict(num_samples=2, note_density=0.5, pitch_hist=2.0), guidance=_GUIDE_SCHED,
                         rules=["pitch_hist", "note_density"]),
}


def rule_targets(B, L, names):
    """Targets in the layout sample_rule.py builds (scripts/sample_rule.py:139-198); L = roll length."""
    nwin = L // 128
    out = {}
    for n in names:
        if n == "pitch_hist":
            out[n] = torch.tensor([0.5, 0, 0, 0, 0.25, 0, 0, 0.25, 0, 0, 0, 0]).repeat(B, 1)
        elif n.startswith("note_density"):
            v = torch.tensor([1., 1, 2, 3, 3, 2, 1, 1])[:nwin] if nwin <= 8 else torch.ones(nwin)
            h = torch.tensor([5., 5, 10, 15, 15, 10, 5, 5])[:nwin] / 5 if nwin <= 8 else torch.ones(nwin)
            out[n] = torch.cat([v, h]).repeat(B, 1)
        else:
            raise KeyError(n)
    return out


def sampler_model_kwargs(cfg):
    B, _, H, _ = cfg["shape"]
    kw = {"y": torch.ones(B, dtype=torch.long)}
    if cfg.get("rules"):
        kw["rule"] = rule_targets(B, H * 8, cfg["rules"])
    return kw


# ---------------------------------------------------------------------------------------------------------------
# extended sampler cases: classifier-guidance hook, replacement editing, DiffCollage long sequences (+ per-segment
# selection), final decode to the uint8 roll.  Goldens in tests/golden/sampler_ext.npz
# ---------------------------------------------------------------------------------------------------------------
def analytic_cond_fn(x, t, y=None, rule=None):
    """Stand-in for a classifier-gradient hook (condition_functions.py:58-85): deterministic, no autograd."""
    return -(x - 0.3) * 0.5


def edit_inputs(device="cpu"):
    g = torch.Generator(device="cpu").manual_seed(31)
    gt = torch.randn(1, 4, 128, 16, generator=g) * 0.7
    mask = torch.zeros(1, 1, 128, 1)
    mask[:, :, :64] = 1.0
    return {"gt": gt.to(device), "mask": mask.to(device), "noise_level": 3, "l_start": 64, "l_end": 128}


EXT_CASES = {
    # DDPM + SCG + classifier guidance applied at every step (gaussian_diffusion.py:691)
    "ddpm_scg_cond": dict(dit="small", respacing="4", ddim=False, shape=(2, 4, 64, 16), seed=7,
                          scg=dict(num_samples=2, pitch_hist=1.0), guidance=_GUIDE_ON, rules=["pitch_hist"],
                          cond=True),
    # replacement-based editing of the second half (edit_kwargs, :293-298, :520-522, :841-852)
    "ddpm_scg_edit": dict(dit="small", respacing="4", ddim=False, shape=(1, 4, 128, 16), seed=8,
                          scg=dict(num_samples=2, pitch_hist=1.0), guidance=_GUIDE_ON, rules=["pitch_hist"],
                          edit=True, rule_len=512),
    # DiffCollage CondIndSimple, 3 windows -> latent length 256, per-segment selection dc.base = 128 (:562-592)
    "dc_simple_base": dict(dit="small", respacing="3", ddim=False, shape=(1, 4, 256, 16), seed=9,
                           scg=dict(num_samples=2, pitch_hist=1.0, note_density=0.5),
                           guidance=dict(_GUIDE_ON, dc=dict(base=128)), rules=["pitch_hist", "note_density"],
                           dc=dict(type="simple", num_img=3)),
    # DiffCollage CondIndCircle (num_img + 1 windows, scripts/sample_rule.py:127), DDIM eta = 1
    "dc_circle": dict(dit="small", respacing="3", ddim=True, eta=1.0, shape=(1, 4, 256, 16), seed=10,
                      scg=dict(num_samples=2, pitch_hist=1.0), guidance=_GUIDE_ON, rules=["pitch_hist"],
                      dc=dict(type="circle", num_img=3)),
}


def ext_model_kwargs(cfg):
    B, _, H, _ = cfg["shape"]
    kw = {"y": torch.ones(B, dtype=torch.long)}
    kw["rule"] = rule_targets(B, cfg.get("rule_len", H * 8), cfg["rules"])
    return kw


# ---------------------------------------------------------------------------------------------------------------
# DiffCollage workers on the CPU with an analytic denoiser (host logic of SURVEY.md section 8 row a12)
# ---------------------------------------------------------------------------------------------------------------
def collage_eps_fn(x, t, y=None):
    """A cheap stand-in for the denoiser with the worker-side signature: any window width, per-row t and y."""
    out = torch.sin(x * (1.0 + 0.01 * t.view(-1, 1, 1, 1).float())) + 0.05 * x.flip(-1)
    if y is not None:
        out = out + 0.1 * y.view(-1, 1, 1, 1).float()
    return out


def collage_inputs(num_img, circle, B=2):
    g = torch.Generator(device="cpu").manual_seed(60 + num_img + (7 if circle else 0))
    W = 128 * num_img - 64 * (num_img if circle else num_img - 1)
    return torch.randn(B, 4, 4, W, generator=g), torch.tensor([17, 803][:B]), torch.tensor([1, 2][:B])


# ---------------------------------------------------------------------------------------------------------------
# host-logic branches of the sampler on the CPU with an analytic denoiser: classifier
# guidance with and without a schedule, DDIM score conditioning, replacement editing, learned-range variances,
# rescaled timesteps, t_end.  No SCG here (its fan-out / decode / select are kernels, covered by the GPU tests).
# ---------------------------------------------------------------------------------------------------------------
def host_model(x, t, y=None, rule=None, learn_sigma=False):
    tt = t.float().view(-1, 1, 1, 1)
    eps = torch.tanh(x) * 0.5 + 0.0007 * tt + 0.02 * x.roll(1, dims=2)
    if y is not None:
        eps = eps + 0.05 * y.float().view(-1, 1, 1, 1)
    if learn_sigma:
        return torch.cat([eps, torch.sin(x + 0.001 * tt)], dim=1)   # second half: variance interpolation in [-1, 1]
    return eps


def host_edit_inputs():
    g = torch.Generator(device="cpu").manual_seed(71)
    gt = torch.randn(2, 4, 32, 16, generator=g) * 0.6
    mask = torch.zeros(2, 1, 32, 1)
    mask[:, :, :16] = 1.0
    return {"gt": gt, "mask": mask, "noise_level": 4, "l_start": 16, "l_end": 32}


_SCHED = dict(schedule=True, t_start=800, t_end=100, interval=2, method="classifier", step_size=1.0, nn=False)
_ALWAYS = dict(schedule=False, t_start=750, t_end=0, interval=1, method="classifier", step_size=1.0, nn=False)
HOST_CASES = {
    "ddpm_cond_sched": dict(respacing="6", ddim=False, seed=81, guidance=_SCHED, cond=True),
    "ddim_cond_score": dict(respacing="5", ddim=True, eta=0.0, seed=82, guidance=_ALWAYS, cond=True),
    "ddim_eta1_tend": dict(respacing="6", ddim=True, eta=1.0, seed=83, t_end=2),
    # (replacement editing together with a classifier-gradient cond_fn raises a shape error in the reference itself,
    # gaussian_diffusion.py:409-413 multiplies the full-length variance with the cropped gradient: not a golden case)
    "ddpm_edit": dict(respacing="6", ddim=False, seed=84, edit=True),
    "ddim_edit": dict(respacing="6", ddim=True, eta=0.5, seed=88, edit=True),
    "ddpm_learn_sigma": dict(respacing="5", ddim=False, seed=85, learn_sigma=True),
    "ddpm_rescaled_t": dict(respacing="ddim4", ddim=False, seed=86, rescale=True),
    "ddpm_no_clip": dict(respacing="4", ddim=False, seed=87, clip=False),
}
HOST_SHAPE = (2, 4, 32, 16)


# ---------------------------------------------------------------------------------------------------------------
# flagship step (BASELINE.json config 3's model and guidance at B = 1): DiTRotary_XL_8, DDIM(eta=1), respacing "256",
# SCG N = 16, pitch histogram.  ONE teacher-forced ddim_sample call per case through the unmodified reference with
# every stage of scg_sample captured (tests/golden/flagship.npz); the GPU test replays the same noise and compares
# stage by stage.
# ---------------------------------------------------------------------------------------------------------------
FLAGSHIP = dict(dit="xl8", respacing="256", eta=1.0, N=16, rules=["pitch_hist"],
                scg=dict(num_samples=16, pitch_hist=1.0), guidance=_GUIDE_ON,
                # spaced timestep index (of 256) -> original timestep 4 * idx (respace.py:38-60): t = 900 and t = 200
                cases={"t900": dict(t_index=225, seed=41), "t200": dict(t_index=50, seed=42)})


def flagship_inputs(case, alphas_cumprod):
    """x_t [1,4,128,16] for the spaced step index of `case`: a seeded latent-like x0 noised to that level
    (`alphas_cumprod` = the SPACED diffusion's table, float64 numpy)."""
    c = FLAGSHIP["cases"][case]
    g = torch.Generator(device="cpu").manual_seed(c["seed"])
    x0 = torch.randn(1, 4, 128, 16, generator=g) * 0.8
    ab = float(alphas_cumprod[c["t_index"]])
    x_t = ab ** 0.5 * x0 + (1 - ab) ** 0.5 * torch.randn(1, 4, 128, 16, generator=g)
    return x_t.float(), torch.tensor([c["t_index"]], dtype=torch.long)


# ---------------------------------------------------------------------------------------------------------------
# chord rule, reference-owned parts (docs/CHORD_SPEC.md): goldens in tests/golden/chords.npz
# ---------------------------------------------------------------------------------------------------------------
def chord_rolls(B=3, L=1024):
    """Rolls with held chords, repeated notes, overlapping voices, a note running to the last frame, out-of-range pitches,
    values around the -0.95 threshold, and low-level 'background' noise in the out-of-range rows."""
    g = torch.Generator(device="cpu").manual_seed(91)
    r = -torch.ones(B, 3, 128, L)
    for b in range(B):
        t = 0
        while t < L - 40:
            d = int(torch.randint(20, 160, (1,), generator=g))
            root = int(torch.randint(40, 80, (1,), generator=g))
            for iv in (0, 4, 7) if b != 1 else (0, 3, 7, 10):
                v = float(torch.rand(1, generator=g)) * 1.4 - 0.5
                r[b, 0, root + iv, t:min(t + d, L)] = v
            t += d + int(torch.randint(0, 12, (1,), generator=g))
    r[0, 0, 60, L - 30:] = 0.9            # still sounding at the end
    r[1, 0, 10, :] = 0.5                  # below the piano range: masked
    r[1, 0, 115, 100:200] = 0.5           # above
    r[2, 0, 64, 300:340] = -0.95          # exactly the threshold: kept (velocity 3)
    r[2, 0, 65, 300:340] = -0.951         # below: removed
    r[2, 0, 66, 400:401] = 1.3            # beyond the range: clamped to 127, one frame long
    return r + 0.0 * torch.randn(B, 3, 128, L, generator=g)


CHORD_FIGURES = ["I", "i6", "V7", "v", "IV64", "iv", "vii/o7", "VII", "VI", "vi6", "iii+64", "#iii6b42", "II", "bII6",
                 "ii/o42", "null", "N6", "Ger65", "It6", "Fr43", "V/V", "viio7/V", "", "IV/vi"]


def chord_vote_cases():
    """(chords [[duration, offset, figure]], end_time, window, total_time) for get_longest_chords."""
    g = torch.Generator(device="cpu").manual_seed(92)
    out = {}
    for k, (win, total) in enumerate(((1.28, 10.24), (1.6, 10.24), (1.28, 5.12))):
        chords, t = [], 0.0
        while t < total * 0.9:
            d = float(torch.rand(1, generator=g)) * 2.0 + 0.05
            fig = CHORD_FIGURES[int(torch.randint(0, 15, (1,), generator=g))]
            chords.append([d, t, fig])
            t += d + (0.7 if k == 1 and len(chords) % 3 == 0 else 0.0)   # gaps -> 'null' windows
        out[f"case{k}"] = (chords, min(t, total), win, total)
    out["short"] = ([[0.5, 0.0, "I"], [0.3, 0.5, "V"]], 0.8, 1.28, 10.24)   # music ends early: padded with 'null'
    return out
