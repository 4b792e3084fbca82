"""CPU-side checks of the host mirror (no GPU, no compute calls into the library): schedule tables and respacing vs
the reference's goldens, the plain sampling loops (pure torch host logic) vs the reference trajectories with the
oracle's denoiser as the model callable, config/registry surface, and the C ABI's exported symbols."""
import ctypes
import os
import re
from functools import partial

import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import dit as odit
from oracle import weights as ow
from rule_guided_music_b200 import _lib
from rule_guided_music_b200.guided_diffusion import respace
from rule_guided_music_b200.guided_diffusion.condition_functions import dc_model_fn, model_fn
from rule_guided_music_b200.guided_diffusion.script_util import create_diffusion

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "rgm_b200.h")).read()
    declared = set(re.findall(r"\b(rgm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert lib.rgm_version() >= 100


def test_no_cpu_fallback():
    """Compute entry points refuse to run without an sm_100 device instead of falling back."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rule_guided_music_b200.guided_diffusion.dit import DiT_models
    from rule_guided_music_b200.music_rule_guidance.rule_maps import FUNC_DICT

    m = DiT_models["DiTRotary_XL_8"](input_size=[128, 16], in_channels=4, num_classes=3, learn_sigma=False)
    with pytest.raises(_lib.RgmError):
        m.to("cpu")
    with pytest.raises(_lib.RgmError):
        FUNC_DICT["pitch_hist"](torch.zeros(1, 3, 128, 1024))
    assert _lib.lib().rgm_check_device() != 0
    assert b"no CUDA device" in _lib.lib().rgm_last_error() or b"sm_" in _lib.lib().rgm_last_error()


def test_schedule_tables_match_reference():
    g = np.load(os.path.join(GOLD, "schedule.npz"))
    d = create_diffusion()
    for k in gi.SCHEDULE_KEYS:
        np.testing.assert_allclose(getattr(d, k), g["full_" + k], rtol=1e-15, atol=0, err_msg=k)
    for resp in ("256", "ddim50", "4", "ddim25", "10,15,20"):
        s = create_diffusion(timestep_respacing=resp)
        tag = resp.replace(",", "_")
        np.testing.assert_array_equal(np.array(s.timestep_map), g[f"resp_{tag}_map"])
        np.testing.assert_allclose(s.betas, g[f"resp_{tag}_betas"], rtol=1e-14)
    with pytest.raises(ValueError):
        respace.space_timesteps(1000, "ddim256")
    c = create_diffusion(diffusion_steps=100, noise_schedule="cosine")
    np.testing.assert_allclose(c.betas, g["cosine100_betas"], rtol=1e-14)


@pytest.mark.parametrize("tag", ["ddpm_plain", "ddim_plain"])
def test_plain_loops_match_reference_on_cpu(tag):
    """Host logic of p_sample / ddim_sample / loops, device tables and the wrapped timestep map, with the oracle DiT."""
    g = np.load(os.path.join(GOLD, "sampler.npz"))
    cfg = gi.SAMPLER_CASES[tag]
    dcfg = gi.DIT_CASES[cfg["dit"]]
    sd = ow.make_dit_state_dict(**dcfg["weights"])
    w = dcfg["weights"]

    class Oracle:
        def __call__(self, x, t, y=None):
            return odit.dit_forward(sd, x, t, y, heads=w["heads"], patch=w["patch"])

        def parameters(self):
            yield torch.zeros(1)

    fn = partial(model_fn, model=Oracle(), num_classes=3, class_cond=True, cfg=False, w=0.0)
    diffusion = create_diffusion(timestep_respacing=cfg["respacing"])
    loop = diffusion.ddim_sample_loop_progressive if cfg["ddim"] else diffusion.p_sample_loop_progressive
    extra = {"eta": cfg["eta"]} if cfg["ddim"] else {}
    torch.manual_seed(cfg["seed"])
    steps = [o["sample"].numpy().copy() for o in
             loop(fn, cfg["shape"], model_kwargs=gi.sampler_model_kwargs(cfg), device="cpu", **extra)]
    np.testing.assert_allclose(np.stack(steps), g[tag], atol=5e-5, rtol=1e-4)


def test_model_fn_dispatch():
    calls = []

    def fake(x, t, y):
        calls.append(y.clone())
        return x * 0 + y.view(-1, 1, 1, 1).float()

    x = torch.zeros(2, 4, 8, 16)
    t = torch.zeros(2, dtype=torch.long)
    y = torch.tensor([0, 2])
    assert torch.equal(model_fn(x, t, y=y, rule={"a": 1}, model=fake)[:, 0, 0, 0], torch.tensor([0., 2.]))
    assert torch.equal(model_fn(x, t, y=y, model=fake, class_cond=False)[:, 0, 0, 0], torch.tensor([3., 3.]))
    out = model_fn(x, t, y=y, model=fake, cfg=True, w=2.0)  # (1+w) m(y) - w m(null)
    assert torch.equal(out[:, 0, 0, 0], torch.tensor([3 * 0. - 2 * 3, 3 * 2. - 2 * 3]))
    xd = torch.arange(2 * 4 * 16 * 8, dtype=torch.float32).view(2, 4, 16, 8)
    seen = {}

    def fake2(x, t, y):
        seen["shape"] = tuple(x.shape)
        return x

    assert torch.equal(dc_model_fn(xd, t, y=y, model=fake2), xd) and seen["shape"] == (2, 4, 8, 16)


def test_registry_and_config_surface(tmp_path):
    from rule_guided_music_b200.guided_diffusion.dit import DiT_models
    from rule_guided_music_b200.guided_diffusion.midi_util import load_config
    from rule_guided_music_b200.music_rule_guidance.rule_maps import FUNC_DICT, LOSS_DICT

    m = DiT_models["DiTRotary_XL_8"](input_size=[128, 16], in_channels=4, num_classes=3, learn_sigma=False)
    assert (m.depth, m.hidden_size, m.num_heads, m.patch_size, m.label_rows, m.out_channels) == (28, 1152, 16, 8, 4, 4)
    assert set(FUNC_DICT) == set(LOSS_DICT) and "pitch_hist" in FUNC_DICT
    p = tmp_path / "c.yml"
    p.write_text("guidance:\n  schedule: true\n  t_start: 750\nscg:\n  num_samples: 16\n  pitch_hist: 1.0\n")
    c = load_config(str(p))
    assert c.guidance.t_start == 750 and vars(c.scg)["num_samples"] == 16


def test_classifier_guidance_hooks_match_autograd_by_hand():
    """composite_nn_zt (condition_functions.py:161-167 of the reference) with tiny torch classifiers: the hooks are
    host-side autograd around the caller's classifier; check them against gradients written out by hand."""
    from rule_guided_music_b200.guided_diffusion import condition_functions as cf

    torch.manual_seed(0)
    x = torch.randn(3, 4, 8, 16)
    t = torch.tensor([5, 5, 5])
    W = torch.randn(4 * 8 * 16, 6)
    reg = lambda z, tt: z.reshape(z.shape[0], -1) @ W + tt.float().view(-1, 1)   # "regression" classifier
    target = torch.randn(3, 6)
    g = cf.grad_nn_zt_mse(x, t, rule=target, classifier_scale=2.0, classifier=reg)
    pred = x.reshape(3, -1) @ W + t.float().view(-1, 1)
    want = (-2.0 * (pred - target) @ W.t()).reshape(x.shape) * 2.0
    torch.testing.assert_close(g, want, rtol=1e-4, atol=1e-4)
    cls = lambda z, tt: z.reshape(z.shape[0], -1) @ W                            # class logits, queried at t = 0
    labels = torch.tensor([[1], [4], [0]])
    g2 = cf.grad_nn_zt_xentropy(x, rule=labels, classifier=cls)
    logits = x.reshape(3, -1) @ W
    p = torch.softmax(logits, -1)
    onehot = torch.zeros_like(p).scatter_(1, labels, 1.0)
    torch.testing.assert_close(g2, ((onehot - p) @ W.t()).reshape(x.shape), rtol=1e-4, atol=1e-5)
    both = cf.composite_nn_zt(x, t, rule={"a": target, "b": target}, fns=["grad_nn_zt_mse", "grad_nn_zt_mse"],
                              classifier_scales=[2.0, 1.0], classifiers=[reg, reg], rule_names=["a", "b"])
    torch.testing.assert_close(both, want * 1.5, rtol=1e-4, atol=1e-4)
    import pytest
    with pytest.raises(NotImplementedError):
        cf.composite_rule(x, t, rule={"pitch_hist": target}, fns=["rule_x0_mse_dummy"], classifier_scales=[1.0],
                          rule_names=["pitch_hist"])


def test_script_util_names_used_by_sample_rule():
    from rule_guided_music_b200.guided_diffusion import script_util as su

    d = su.model_and_diffusion_defaults()
    assert su.NUM_CLASSES == 3 and d["timestep_respacing"] == "" and d["image_size"] == 128
    import argparse, pytest
    p = argparse.ArgumentParser()
    su.add_dict_to_argparser(p, d)
    assert p.parse_args(["--diffusion_steps", "500"]).diffusion_steps == 500
    with pytest.raises(NotImplementedError):
        su.create_model_and_diffusion()


def test_eval_rule_loss_report_layout():
    """midi_util.eval_rule_loss with a user-registered (pure-torch) rule: one row per sample, the reference's columns."""
    from rule_guided_music_b200.guided_diffusion import midi_util
    from rule_guided_music_b200.music_rule_guidance import rule_maps

    rule_maps.FUNC_DICT["mean_velocity"] = lambda roll: roll[:, 0].mean(dim=(1, 2)).unsqueeze(-1)
    rule_maps.LOSS_DICT["mean_velocity"] = lambda gen, tgt: ((gen - tgt) ** 2).mean(dim=-1)
    try:
        rolls = torch.linspace(-1, 1, 2 * 3 * 128 * 256).reshape(2, 3, 128, 256)
        target = torch.tensor([[0.0], [0.5]])
        df = midi_util.eval_rule_loss(rolls, {"mean_velocity": target})
        assert list(df.columns) == ["mean_velocity.target_rule", "mean_velocity.gen_rule", "mean_velocity.loss"]
        assert len(df) == 2
        want = (rolls[:, 0].mean(dim=(1, 2)) - target[:, 0]) ** 2
        torch.testing.assert_close(torch.tensor(df["mean_velocity.loss"].tolist()), want)
    finally:
        del rule_maps.FUNC_DICT["mean_velocity"], rule_maps.LOSS_DICT["mean_velocity"]


def test_diff_collage_workers_match_reference_on_cpu():
    """CondIndSimple / CondIndCircle and the window split / merge are host logic in
    torch: with an analytic denoiser they run on the CPU and must reproduce the reference's workers
    (tests/golden/collage.npz, made by diff_collage/*.py of the reference with the same inputs)."""
    from rule_guided_music_b200 import diff_collage as dc
    from rule_guided_music_b200.diff_collage.w_img import avg_merge_wimg, split_wimg

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collage.npz"))
    for n in (2, 3, 5):
        for circle in (False, True):
            x, t, y = gi.collage_inputs(n, circle)
            cls = dc.CondIndCircle if circle else dc.CondIndSimple
            worker = cls((4, 4, 128), gi.collage_eps_fn, n, overlap_size=64)
            tag = f"n{n}_{'circle' if circle else 'long'}"
            assert tuple(worker.shape) == (4, 4, x.shape[-1])
            np.testing.assert_allclose(worker.eps_scalar_t_fn(x, t, y=y)[..., ::2].numpy(), gold[tag + "__eps_y"],
                                       rtol=1e-6, atol=1e-6, err_msg=tag)
            np.testing.assert_allclose(worker.eps_scalar_t_fn(x, t)[..., 1::2].numpy(), gold[tag + "__eps_noy"],
                                       rtol=1e-6, atol=1e-6, err_msg=tag)
    x, _, _ = gi.collage_inputs(4, False)
    tiles, ov = split_wimg(x, 4)
    assert ov == int(gold["split4_overlap"])
    np.testing.assert_array_equal(tiles.numpy(), gold["split4"])
    np.testing.assert_allclose(avg_merge_wimg(tiles * 2.0, ov, n=4).numpy(), gold["merge4_avg"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(avg_merge_wimg(tiles, ov, n=4, is_avg=False).numpy(), gold["merge4_sum"], rtol=1e-6,
                               atol=1e-6)


@pytest.mark.parametrize("tag", list(gi.HOST_CASES))
def test_sampler_host_branches_match_reference_on_cpu(tag):
    """Classifier guidance (with / without schedule), DDIM score conditioning, replacement editing, learned-range
    variance, rescaled timesteps, t_end, clip off: host logic of rows a4-a6, run on the CPU with an analytic denoiser
    against trajectories of the unmodified reference (tests/golden/host.npz)."""
    from types import SimpleNamespace

    from rule_guided_music_b200.guided_diffusion.script_util import create_diffusion as full_create

    g = np.load(os.path.join(GOLD, "host.npz"))
    cfg = gi.HOST_CASES[tag]
    d = full_create(learn_sigma=cfg.get("learn_sigma", False), diffusion_steps=1000, noise_schedule="linear",
                    timestep_respacing=cfg["respacing"], use_kl=False, predict_xstart=False,
                    rescale_timesteps=cfg.get("rescale", False), rescale_learned_sigmas=False)
    fn = partial(gi.host_model, learn_sigma=cfg.get("learn_sigma", False))
    guidance = SimpleNamespace(**cfg["guidance"]) if cfg.get("guidance") else None
    loop = d.ddim_sample_loop_progressive if cfg["ddim"] else d.p_sample_loop_progressive
    extra = {"eta": cfg["eta"]} if cfg["ddim"] else {}
    torch.manual_seed(cfg["seed"])
    d.t_end = cfg.get("t_end", 0)
    steps = [o["sample"].numpy().copy() for o in loop(
        fn, gi.HOST_SHAPE, model_kwargs={"y": torch.tensor([1, 2])}, device="cpu", t_end=cfg.get("t_end", 0),
        clip_denoised=cfg.get("clip", True), cond_fn=gi.analytic_cond_fn if cfg.get("cond") else None,
        guidance_kwargs=guidance, edit_kwargs=gi.host_edit_inputs() if cfg.get("edit") else None, **extra)]
    ref = g[tag]
    assert len(steps) == ref.shape[0]
    np.testing.assert_allclose(np.stack(steps), ref, atol=2e-5, rtol=1e-4)


def test_decode_sample_for_midi_generic_branch_on_cpu():
    """midi_util.decode_sample_for_midi with a NON-native embed_model (the oracle decoder on the CPU): the re-tiling,
    the -0.95 threshold and the uint8 quantisation are host logic; against the reference's output (sampler_ext.npz)."""
    from oracle import vae as ovae
    from rule_guided_music_b200.guided_diffusion.midi_util import decode_sample_for_midi

    vsd = ow.make_vae_state_dict(seed=gi.VAE_SEED)

    class Embed:
        @staticmethod
        def decode(z):
            with torch.no_grad():
                return ovae.vae_decode(vsd, z)

    got = decode_sample_for_midi(gi.vae_latents(), Embed, gi.SCALE_FACTOR, threshold=-0.95).numpy()
    ref = np.load(os.path.join(GOLD, "sampler_ext.npz"))["midi_roll"]
    assert got.shape == ref.shape and got.dtype == ref.dtype == np.uint8
    # the oracle decoder agrees with the reference's to ~2e-5, so a value may land on the other side of an integer
    # boundary of the uint8 quantisation once in a while; never by more than one step
    diff = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3


def test_vae_gn_timeouts_without_a_device_handle():
    """The diagnostic accessor of the in-epilogue GroupNorm path must not need a device (or the library) before the
    model has been moved to one."""
    from rule_guided_music_b200.taming.models.klvae_pedal import AutoencoderKL
    from oracle import weights as ow

    v = AutoencoderKL(ddconfig=ow.VAE_DDCONFIG, embed_dim=4)
    assert v.gn_timeouts() == 0
