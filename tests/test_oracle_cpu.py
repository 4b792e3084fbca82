"""Pins the CPU oracle (oracle/) against outputs of the UNMODIFIED reference (tests/golden/*.npz, produced by
tests/golden/make_golden.py in the build container).  Tolerances: schedule tables are float64 and must match to 1e-15
relative; integer/threshold work must be exact; fp32 network outputs are the same torch-CPU ops in the same order, so
1e-5 absolute on O(1) values."""
import os
from functools import partial
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import dit as odit
from oracle import rules as orules
from oracle import sampler as osampler
from oracle import schedule as osched
from oracle import vae as ovae
from oracle import weights as ow

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def test_schedule_tables():
    g = _load("schedule")
    # create_diffusion always builds a SpacedDiffusion, which re-derives the betas from the cumulative products even
    # when every timestep is kept (respace.py:72-86) -- a 1e-16 rounding difference the oracle follows
    base = osched.named_beta_schedule("linear", 1000)
    tab = osched.diffusion_tables(osched.spaced_betas(base, range(1000))[0])
    for k in gi.SCHEDULE_KEYS:
        np.testing.assert_allclose(tab[k], g["full_" + k], rtol=1e-15, atol=0, err_msg=k)
    np.testing.assert_allclose(tab["fixed_large_variance"], g["full_fixed_large_variance"], rtol=1e-15)
    # SURVEY appendix C known answers
    assert abs(tab["betas"][0] - 1e-4) < 1e-15 and abs(tab["betas"][999] - 2e-2) < 1e-15
    assert abs(tab["alphas_cumprod"][500] - 0.07779665836502389) < 1e-15


@pytest.mark.parametrize("resp", ["256", "ddim50", "4", "ddim25", "10,15,20"])
def test_respacing(resp):
    g = _load("schedule")
    base = osched.named_beta_schedule("linear", 1000)
    betas, tmap = osched.spaced_betas(base, osched.space_timesteps(1000, resp))
    tag = resp.replace(",", "_")
    np.testing.assert_array_equal(np.array(tmap), g[f"resp_{tag}_map"])
    np.testing.assert_allclose(betas, g[f"resp_{tag}_betas"], rtol=1e-14)


def test_respacing_errors_and_cosine():
    g = _load("schedule")
    assert int(g["ddim256_raises"]) == 1
    with pytest.raises(ValueError):
        osched.space_timesteps(1000, "ddim256")
    cos = osched.spaced_betas(osched.named_beta_schedule("cosine", 100), range(100))[0]
    np.testing.assert_allclose(cos, g["cosine100_betas"], rtol=1e-14)
    assert osched.guide_schedule(749) and not osched.guide_schedule(750) and osched.guide_schedule(0)
    assert not osched.guide_schedule(10, 750, 0, 2) and osched.guide_schedule(11, 750, 0, 2)


def test_rules_against_reference():
    g = _load("rules")
    for case, roll in gi.rule_rolls().items():
        for name in gi.RULE_NAMES:
            key = f"{case}__{name}"
            if key + "__raises" in g.files:
                with pytest.raises(IndexError):
                    orules.FUNC_DICT[name](roll.clone())
                continue
            got = orules.FUNC_DICT[name](roll.clone()).numpy()
            ref = g[key]
            assert got.shape == ref.shape, key
            if got.dtype.kind in "iu":
                np.testing.assert_array_equal(got, ref, err_msg=key)
            else:
                np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-7, err_msg=key)


def test_rules_order_dependence_and_losses():
    g = _load("rules")
    r = gi.rule_rolls()["order"]
    orules.FUNC_DICT["note_density"](r)
    np.testing.assert_allclose(orules.FUNC_DICT["pitch_hist"](r).numpy(), g["order__pitch_hist_after_nd"], rtol=1e-6,
                               atol=1e-7)
    np.testing.assert_array_equal(r[:, 0, 55:70, :16].numpy(), g["order__roll_after"])
    a, b = gi.loss_pairs()
    np.testing.assert_allclose(orules.LOSS_DICT["pitch_hist"](a, b).numpy(), g["loss_mse"], rtol=1e-6)
    np.testing.assert_array_equal(orules.LOSS_DICT["note_density_class"](a.round().long(), b.round().long()).numpy(),
                                  g["loss_zero_one"])


def test_rule_known_answers():
    """SURVEY.md appendix C."""
    rolls = gi.rule_rolls()
    ph = orules.FUNC_DICT["pitch_hist"](rolls["kat"].clone())
    assert ph[0].argmax() == 0 and abs(ph[0, 0] - 1) < 1e-6
    nd = orules.FUNC_DICT["note_density"](rolls["kat"].clone())
    np.testing.assert_allclose(nd[0].numpy(), [1, 0, 0, 0, 0, 0, 0, 0, 0.2, 0, 0, 0, 0, 0, 0, 0], atol=1e-7)
    np.testing.assert_allclose(nd[1].numpy(), [0.4375, 2.0, 0.6875, 0, 0.0078125, 0, 0, 0, 0.2, 0, 0, 0, 0.2, 0, 0, 0],
                               atol=1e-7)
    assert torch.tensor([[0, 1], [0, 1], [-1, 1]]).argmax(0).tolist() == [0, 0]


def _dit_kw(cfg):
    w = cfg["weights"]
    return dict(heads=w["heads"], patch=w["patch"])


@pytest.mark.parametrize("tag", ["small", "small_hd64", "xl8"])
def test_dit_against_reference(tag):
    g = _load("dit")
    cfg = gi.DIT_CASES[tag]
    sd = ow.make_dit_state_dict(**cfg["weights"])
    x, t, y = gi.dit_inputs(cfg)
    with torch.no_grad():
        out = odit.dit_forward(sd, x, t, y, **_dit_kw(cfg)).numpy()
    assert np.abs(g[tag]).max() > 1e-2  # the golden is not the all-zero output of an adaLN-zero init
    np.testing.assert_allclose(out, g[tag], atol=2e-5, rtol=1e-4)
    if cfg.get("half_tile"):
        with torch.no_grad():
            outh = odit.dit_forward(sd, x[:, :, :64].contiguous(), t, y, **_dit_kw(cfg)).numpy()
        np.testing.assert_allclose(outh, g[tag + "__half"], atol=2e-5, rtol=1e-4)


def test_classifier_against_reference():
    """DiTRotaryClassifier forward (T = 257 with the class token) and the input-gradient of its log-probability that
    classifier guidance uses (condition_functions.py:45-55), against the unmodified reference."""
    g = _load("classifier")
    cfg = gi.CLASSIFIER_CASE
    sd = ow.make_classifier_state_dict(**cfg["weights"])
    x, t, labels = gi.classifier_inputs(cfg)
    kw = _dit_kw(cfg)
    with torch.no_grad():
        logits = odit.classifier_forward(sd, x, t, **kw).numpy()
        logits_t0 = odit.classifier_forward(sd, x, torch.zeros(x.shape[0]), **kw).numpy()
    assert np.abs(g["logits"]).max() > 1e-2 and np.abs(g["logits"] - g["logits_t0"]).max() > 1e-4  # t matters
    np.testing.assert_allclose(logits, g["logits"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(logits_t0, g["logits_t0"], atol=2e-5, rtol=1e-4)
    grad = odit.classifier_xentropy_grad(sd, x, labels, **kw).numpy()
    assert np.abs(g["xentropy_grad"]).max() > 1e-4
    np.testing.assert_allclose(grad, g["xentropy_grad"], atol=2e-6 + 1e-3 * np.abs(g["xentropy_grad"]).max(), rtol=0)


def test_vae_against_reference():
    g = _load("vae")
    sd = ow.make_vae_state_dict(seed=gi.VAE_SEED)
    with torch.no_grad():
        tiles = ovae.vae_decode(sd, gi.vae_tiles()).numpy()
        roll = ovae.decode_latents(sd, gi.vae_latents(), gi.SCALE_FACTOR)
    np.testing.assert_allclose(tiles, g["tiles"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(roll[:, :, ::4, ::4].numpy(), g["decode_latents_sub4"], atol=2e-5, rtol=1e-4)


def test_vae_encoder_against_reference():
    """Encoder + quant_conv and _encode against the unmodified reference's outputs."""
    g = _load("vae_enc")
    sd = ow.make_vae_encoder_state_dict(seed=gi.VAE_ENC_SEED)
    rolls = gi.vae_rolls()
    tiles = torch.cat(torch.chunk(rolls, rolls.shape[-1] // 128, dim=-1), dim=0)
    with torch.no_grad():
        moments = ovae.vae_encode(sd, tiles).numpy()
        lat = ovae.encode_rolls(sd, rolls, gi.SCALE_FACTOR).numpy()
    assert moments.shape == (4, 8, 16, 16) and lat.shape == (2, 4, 32, 16)
    np.testing.assert_allclose(moments, g["moments"], atol=2e-5, rtol=1e-4)
    np.testing.assert_allclose(lat, g["encode_latents"], atol=2e-5, rtol=1e-4)


def oracle_model_fn(sd, cfg):
    kw = _dit_kw(cfg)

    def fn(x, t, y=None, rule=None):  # condition_functions.py:17-28 with class_cond=True, cfg=False
        return odit.dit_forward(sd, x, t, y, **kw)

    return fn


@pytest.mark.parametrize("tag", list(gi.SAMPLER_CASES))
def test_sampler_against_reference(tag):
    g = _load("sampler")
    cfg = gi.SAMPLER_CASES[tag]
    dcfg = gi.DIT_CASES[cfg["dit"]]
    sd = ow.make_dit_state_dict(**dcfg["weights"])
    vsd = ow.make_vae_state_dict(seed=gi.VAE_SEED)
    diff = osampler.OracleDiffusion(timestep_respacing=cfg["respacing"])
    kwargs = gi.sampler_model_kwargs(cfg)
    guidance = SimpleNamespace(**cfg["guidance"]) if cfg.get("guidance") else None
    step = diff.ddim_sample if cfg["ddim"] else diff.p_sample
    extra = {"eta": cfg["eta"]} if cfg["ddim"] else {}
    decode = partial(ovae.decode_latents, vsd, scale_factor=gi.SCALE_FACTOR) if cfg["scg"] else None
    torch.manual_seed(cfg["seed"])
    steps = [o["sample"].numpy().copy() for o in
             diff._loop(step, oracle_model_fn(sd, dcfg), cfg["shape"], None, cfg.get("t_end", 0), model_kwargs=kwargs,
                        decode_fn=decode, guidance_kwargs=guidance,
                        scg_kwargs=dict(cfg["scg"]) if cfg["scg"] else None, **extra)]
    ref = g[tag]
    assert len(steps) == ref.shape[0]
    np.testing.assert_allclose(np.stack(steps), ref, atol=5e-5, rtol=1e-4)
