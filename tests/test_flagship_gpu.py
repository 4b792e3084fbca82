"""The flagship step -- BASELINE.json config 3's model and guidance at B = 1 -- on the B200 against the UNMODIFIED
reference, stage by stage (tests/golden/flagship.npz, written by tests/golden/make_golden.py flagship):

    ddim_sample (gaussian_diffusion.py:881-976): DiTRotary_XL_8 on x_t -> pred_xstart, mean_pred, sigma
    scg_sample  (:491-554): fan-out to N = 16 candidates -> DiTRotary_XL_8 on 16 candidates -> x0 -> _decode (128 VAE
                tiles) -> pitch histogram -> -MSE -> argmax -> winner

Teacher-forced: x_t is regenerated from the same seed, the 16 noise draws come from the same CPU generator stream
(noise tape).  Every stage's relative L2 error is recorded (conftest `parity`) and asserted:
  * everything that feeds the sampled latent (eps, mean_pred, candidates, candidate eps and x0, the chosen x_(t-1))
    within 1e-3 -- BASELINE.json's tolerance; the B-pass pred_xstart within 1e-3 times its conditioning (see below);
  * the decoded roll within 5e-3 (fp16 activations through 30 GroupNorm layers; it only feeds the scores);
  * candidate scores within 2e-3 relative; the argmax index equal to the reference's wherever the reference's margin
    between best and runner-up exceeds twice the measured score error (index work is exact given equal scores:
    tests/test_rules_gpu.py).
"""
import os
from functools import partial
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import golden_inputs as gi
import gpu_util
from rule_guided_music_b200.guided_diffusion import gaussian_diffusion as gd
from rule_guided_music_b200.guided_diffusion.condition_functions import model_fn
from rule_guided_music_b200.guided_diffusion.script_util import create_diffusion
from rule_guided_music_b200.music_rule_guidance import music_rules

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flagship.npz"))
LATENT_TOL = 1e-3   # BASELINE.json north_star: sampled latents within 1e-3 rel-fp of the reference
ROLL_TOL = 5e-3
SCORE_TOL = 2e-3


class _SpyDecoder:
    """Forwards to the native AutoencoderKL and keeps what went in and came out of the fused _decode."""

    def __init__(self, vae):
        self.vae, self.x0, self.roll = vae, None, None

    def decode_latents(self, latents, scale_factor=1.0, channels=None):
        self.x0 = latents.clone()
        roll = self.vae.decode_latents(latents, scale_factor, channels=channels)
        self.roll = roll.clone()  # before the rule kernels mask it in place
        return roll


@pytest.fixture(scope="module")
def flagship_models(cuda):
    model, _ = gpu_util.native_dit(gi.DIT_CASES[gi.FLAGSHIP["dit"]], cuda)
    vae, _ = gpu_util.native_vae(cuda)
    return model, vae


@pytest.mark.parametrize("case", list(gi.FLAGSHIP["cases"]))
def test_flagship_step_matches_reference(cuda, flagship_models, parity, case):
    cfg = gi.FLAGSHIP
    c = cfg["cases"][case]
    model, vae = flagship_models
    N = cfg["N"]
    diffusion = create_diffusion(timestep_respacing=cfg["respacing"])
    x_t, t = gi.flagship_inputs(case, diffusion.alphas_cumprod)
    cap = {}

    def spy_model(x, t_, y=None):
        o = model(x, t_, y=y)
        if x.shape[0] > 1:
            cap["cand"], cap["eps"], cap["t_model"] = x.clone(), o.clone(), t_.clone()
        else:
            cap["eps_b"] = o.clone()
        return o

    spy_vae = _SpyDecoder(vae)
    real_scg = diffusion.scg_sample

    def spy_scg(m, t_, mean_pred, g_coeff, *a, **k):
        cap["mean_pred"], cap["g"] = mean_pred.clone(), g_coeff.reshape(mean_pred.shape[0], -1)[:, 0].clone()
        return real_scg(m, t_, mean_pred, g_coeff, *a, **k)

    diffusion.scg_sample = spy_scg
    diffusion._trace = []
    fn = partial(model_fn, model=spy_model, num_classes=3, class_cond=True, cfg=False, w=0.0)
    kwargs = {"y": torch.ones(1, dtype=torch.long, device=cuda),
              "rule": {n: v.to(cuda) for n, v in gi.rule_targets(1, 1024, cfg["rules"]).items()}}
    diffusion.t_end = 0
    with gpu_util.cpu_noise_tape(gd.th, c["seed"]), torch.no_grad():
        out = diffusion.ddim_sample(fn, x_t.to(cuda), t.to(cuda), model_kwargs=kwargs, eta=cfg["eta"],
                                    embed_model=spy_vae, scale_factor=gi.SCALE_FACTOR,
                                    guidance_kwargs=SimpleNamespace(**cfg["guidance"]), scg_kwargs=dict(cfg["scg"]),
                                    _t_host=int(t[0]))
    torch.cuda.synchronize()
    G = lambda k: torch.from_numpy(GOLD[f"{case}__{k}"])  # noqa: E731

    # the denoiser saw the reference's original-scale timestep (respace.py:116-128), exactly
    assert torch.equal(cap["t_model"].cpu().long(), G("t_model").long())
    lat = {}
    lat["eps (B=1 pass)"] = gpu_util.rel_l2(cap["eps_b"].cpu(), G("eps_b"))
    lat["mean_pred"] = gpu_util.rel_l2(cap["mean_pred"].cpu(), G("mean_pred"))
    lat["sigma"] = gpu_util.rel_l2(cap["g"].cpu(), G("g"))
    lat["candidates x_(t-1)"] = gpu_util.rel_l2(cap["cand"][:, :, ::2].cpu(), G("cand_sub"))
    lat["candidate eps (N=16 pass)"] = gpu_util.rel_l2(cap["eps"][:, :, ::2].cpu(), G("eps_sub"))
    lat["candidate x0"] = gpu_util.rel_l2(spy_vae.x0[:, :, ::2].cpu(), G("x0_sub"))
    for k, v in lat.items():
        parity(f"flagship {case}: {k}", v, LATENT_TOL)
    # pred_xstart = clip(sqrt_recip * x - sqrt_recipm1 * eps) is not a sampled latent and is ill-conditioned at high noise
    # (sqrt_recipm1[t = 900] ~ 40: the un-clipped entries amplify the denoiser's error by that factor): its bar is the
    # latent tolerance times that conditioning, c_t * |eps| / |x0|, and never below the latent tolerance itself
    c_t = float(diffusion.sqrt_recipm1_alphas_cumprod[int(t[0])])
    cond = max(1.0, c_t * G("eps_b").norm().item() / G("pred_xstart").norm().item())
    x0_err = parity(f"flagship {case}: pred_xstart (conditioning x{cond:.0f})",
                    gpu_util.rel_l2(out["pred_xstart"].cpu(), G("pred_xstart")), LATENT_TOL * cond)
    assert x0_err < LATENT_TOL * cond

    roll = spy_vae.roll
    assert roll.shape == (N, 1, 128, 1024)
    roll_err = parity(f"flagship {case}: decoded roll (channel 0, sub-sampled)",
                      gpu_util.rel_l2(roll[:, 0, ::4, ::8].cpu(), G("roll_sub")), ROLL_TOL)
    stats = G("roll_stats")
    parity(f"flagship {case}: fraction of roll pixels above the -0.95 note threshold (abs diff)",
           abs((roll[:, 0] > -0.95).float().mean().item() - stats[2].item()), 5e-3, "abs")
    hist = music_rules.total_pitch_class_histogram(roll.clone())
    hist_err = parity(f"flagship {case}: pitch histograms of the 16 candidates", gpu_util.rel_l2(hist.cpu(), G("hist")),
                      SCORE_TOL)

    total, idx = diffusion._trace[0]
    ref_total = G("total")
    dev = (total.cpu() - ref_total).abs().max().item()
    score_err = parity(f"flagship {case}: candidate scores total_log_prob (max abs / max |ref|)",
                       dev / ref_total.abs().max().item(), SCORE_TOL, "rel-max")
    top = ref_total.view(-1).topk(2).values
    gap = (top[0] - top[1]).item()
    parity(f"flagship {case}: reference margin best vs runner-up / measured score error", gap / max(dev, 1e-30), 2.0,
           "ratio>=bar")
    same = bool(torch.equal(idx.cpu(), G("max_ind")))
    if gap > 2 * dev:
        assert same, (case, "argmax", idx.cpu(), G("max_ind"), gap, dev)
    if same:  # the same candidate chosen: the sampled latent is the reference's within the latent tolerance
        lat["chosen x_(t-1)"] = parity(f"flagship {case}: chosen x_(t-1)",
                                       gpu_util.rel_l2(out["sample"].cpu(), G("sample")), LATENT_TOL)
    else:     # a near-tie decided differently: the chosen latent must still be ONE of the reference's candidates
        mine = cap["cand"].view(N, *out["sample"].shape)[int(idx[0]), 0]
        assert torch.equal(out["sample"][0], mine)
    assert max(lat.values()) < LATENT_TOL, lat
    assert roll_err < ROLL_TOL and hist_err < SCORE_TOL and score_err < SCORE_TOL
