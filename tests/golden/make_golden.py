"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
The reference needs nine packages that are not installed here; `ref_shims/` holds import stand-ins (no-ops for
matplotlib/mpi4py/blobfile/omegaconf/pytorch_lightning/mido/music21, and behavioural restatements of timm.Mlp and
rotary_embedding_torch.RotaryEmbedding -- see oracle/__init__.py for what that means for pinning).
Inputs are regenerated from seeds by the tests (tests/golden_inputs.py), so only reference OUTPUTS are stored.
"""
import os
import sys
from functools import partial
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "ref_shims"))
sys.path.insert(0, os.environ.get("RGM_REFERENCE", "/root/reference"))

import guided_diffusion.gaussian_diffusion as gd  # noqa: E402
import guided_diffusion.dit as rdit  # noqa: E402
from guided_diffusion import respace as rrespace  # noqa: E402
from guided_diffusion.condition_functions import model_fn  # noqa: E402
from guided_diffusion.script_util import create_diffusion as _create_diffusion  # noqa: E402
from music_rule_guidance.rule_maps import FUNC_DICT, LOSS_DICT  # noqa: E402
from taming.modules.diffusionmodules.model import Decoder, Encoder  # noqa: E402

from oracle import weights as ow  # noqa: E402
import golden_inputs as gi  # noqa: E402

torch.set_grad_enabled(False)


def create_diffusion(diffusion_steps=1000, noise_schedule="linear", timestep_respacing="", learn_sigma=False):
    return _create_diffusion(learn_sigma=learn_sigma, diffusion_steps=diffusion_steps, noise_schedule=noise_schedule,
                             timestep_respacing=timestep_respacing, use_kl=False, predict_xstart=False,
                             rescale_timesteps=False, rescale_learned_sigmas=False)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


# ---------------------------------------------------------------------------------------------------------------
def golden_schedule():
    out = {}
    d = create_diffusion(diffusion_steps=1000, noise_schedule="linear", timestep_respacing="")
    for k in gi.SCHEDULE_KEYS:
        out["full_" + k] = getattr(d, k)
    out["full_fixed_large_variance"] = np.append(d.posterior_variance[1], d.betas[1:])
    for resp in ("256", "ddim50", "4", "ddim25", "10,15,20"):
        s = create_diffusion(diffusion_steps=1000, noise_schedule="linear", timestep_respacing=resp)
        tag = resp.replace(",", "_")
        out[f"resp_{tag}_betas"] = s.betas
        out[f"resp_{tag}_map"] = np.array(s.timestep_map)
    try:
        rrespace.space_timesteps(1000, "ddim256")
        out["ddim256_raises"] = 0
    except ValueError:
        out["ddim256_raises"] = 1
    c = create_diffusion(diffusion_steps=100, noise_schedule="cosine", timestep_respacing="")
    out["cosine100_betas"] = c.betas
    save("schedule", **out)


def golden_rules():
    out = {}
    for case, roll in gi.rule_rolls().items():
        for name in gi.RULE_NAMES:
            try:
                out[f"{case}__{name}"] = FUNC_DICT[name](roll.clone()).numpy()
            except IndexError:  # note_density_class on a B == 1 roll: the reference indexes a squeezed tensor
                out[f"{case}__{name}__raises"] = 1
    # quantize_factor != 1 (music_rules.py:59-61): nearest resampling of the time axis first, input left untouched
    from music_rule_guidance.music_rules import note_density as ref_note_density
    for q in gi.RULE_QUANT:
        r = gi.rule_rolls()["random"]
        out[f"random__note_density_q{q}"] = ref_note_density(r, quantize_factor=q).numpy()
        out[f"random__note_density_q{q}__input_untouched"] = int(torch.equal(r, gi.rule_rolls()["random"]))
    # order dependence: pitch_hist evaluated after note_density on the SAME tensor (in-place threshold)
    r = gi.rule_rolls()["order"]
    FUNC_DICT["note_density"](r)
    out["order__pitch_hist_after_nd"] = FUNC_DICT["pitch_hist"](r).numpy()
    out["order__roll_after"] = r[:, 0, 55:70, :16].numpy()
    # losses
    g, t = gi.loss_pairs()
    out["loss_mse"] = LOSS_DICT["pitch_hist"](g, t).numpy()
    out["loss_zero_one"] = LOSS_DICT["note_density_class"](g.round().long(), t.round().long()).numpy()
    save("rules", **out)


def build_ref_dit(cfg):
    sd = ow.make_dit_state_dict(**cfg["weights"])
    if cfg.get("preset"):
        model = rdit.DiT_models[cfg["preset"]](input_size=cfg["input_size"], in_channels=4,
                                               num_classes=cfg["weights"].get("num_classes", 3),
                                               learn_sigma=cfg["weights"].get("learn_sigma", False))
    else:
        w = cfg["weights"]
        model = rdit.DiTRotary(input_size=cfg["input_size"], patch_size=w["patch"], in_channels=4,
                               hidden_size=w["hidden"], depth=w["depth"], num_heads=w["heads"],
                               num_classes=w.get("num_classes", 3), learn_sigma=w.get("learn_sigma", False))
    missing, unexpected = model.load_state_dict(sd, strict=True), None
    del missing, unexpected
    return model.eval(), sd


def golden_dit():
    out = {}
    for tag, cfg in gi.DIT_CASES.items():
        model, _ = build_ref_dit(cfg)
        x, t, y = gi.dit_inputs(cfg)
        out[tag] = model(x, t, y).numpy()
        if cfg.get("half_tile"):
            xh = x[:, :, :64].contiguous()
            out[tag + "__half"] = model(xh, t, y).numpy()
        del model
    save("dit", **out)


def golden_classifier():
    """The unmodified DiTRotaryClassifier (dit.py:735-831) and the input-gradient classifier guidance takes of it
    (condition_functions.grad_nn_zt_xentropy :45-55): groundwork for SURVEY.md 8(f) rank 1."""
    from guided_diffusion.condition_functions import grad_nn_zt_xentropy

    cfg = gi.CLASSIFIER_CASE
    w = cfg["weights"]
    sd = ow.make_classifier_state_dict(**w)
    model = rdit.DiTRotaryClassifier(input_size=cfg["input_size"], patch_size=w["patch"], in_channels=4,
                                     hidden_size=w["hidden"], depth=w["depth"], num_heads=w["heads"],
                                     num_classes=w["num_classes"])
    model.load_state_dict(sd, strict=True)
    model.eval()
    x, t, labels = gi.classifier_inputs(cfg)
    logits = model(x, t)
    logits_t0 = model(x, torch.zeros(x.shape[0]))
    grad = grad_nn_zt_xentropy(x, rule=labels, classifier=model)
    save("classifier", logits=logits.numpy(), logits_t0=logits_t0.numpy(), xentropy_grad=grad.numpy())


def build_ref_vae():
    sd = ow.make_vae_state_dict(seed=gi.VAE_SEED)
    dec = Decoder(**ow.VAE_DDCONFIG)
    pq = torch.nn.Conv2d(4, 4, 1)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    pq.load_state_dict({k[len("post_quant_conv."):]: v for k, v in sd.items() if k.startswith("post_quant_conv.")},
                       strict=True)

    class Embed:  # what AutoencoderKL.decode does (klvae_pedal.py:80-85)
        @staticmethod
        def decode(z):
            return dec(pq(z))

    return Embed


def golden_vae():
    embed = build_ref_vae()
    z = gi.vae_tiles()
    out = {"tiles": embed.decode(z).numpy().astype(np.float32)}
    lat = gi.vae_latents()
    roll = gd._decode(lat, embed, scale_factor=gi.SCALE_FACTOR)
    out["decode_latents_sub4"] = roll[:, :, ::4, ::4].numpy()
    out["decode_latents_ch0_stats"] = np.array([roll[:, 0].mean().item(), roll[:, 0].std().item()])
    save("vae", **out)


def golden_vae_enc():
    """Encoder + quant_conv of the reference (model.py:342-433, klvae_pedal.py:60-68) and gaussian_diffusion._encode."""
    sd = ow.make_vae_encoder_state_dict(seed=gi.VAE_ENC_SEED)
    enc = Encoder(**ow.VAE_DDCONFIG)
    qc = torch.nn.Conv2d(8, 8, 1)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    qc.load_state_dict({k[len("quant_conv."):]: v for k, v in sd.items() if k.startswith("quant_conv.")}, strict=True)

    class Embed:  # AutoencoderKL.encode_save with range_fix=False (klvae_pedal.py:60-68)
        @staticmethod
        def encode_save(x, range_fix=False):
            assert not range_fix
            return qc(enc(x))

    rolls = gi.vae_rolls()
    tiles = torch.cat(torch.chunk(rolls, rolls.shape[-1] // 128, dim=-1), dim=0)
    out = {"moments": Embed.encode_save(tiles).numpy(),
           "encode_latents": gd._encode(rolls, Embed, scale_factor=gi.SCALE_FACTOR).numpy()}
    save("vae_enc", **out)


def golden_collage():
    """DiffCollage workers and window split / merge of the reference (diff_collage/*.py) with an analytic denoiser."""
    import diff_collage as dc
    from diff_collage.w_img import avg_merge_wimg, split_wimg

    out = {}
    for n in (2, 3, 5):
        for circle in (False, True):
            x, t, y = gi.collage_inputs(n, circle)
            cls = dc.CondIndCircle if circle else dc.CondIndSimple
            worker = cls((4, 4, 128), gi.collage_eps_fn, n, overlap_size=64)
            tag = f"n{n}_{'circle' if circle else 'long'}"
            assert tuple(worker.shape) == (4, 4, x.shape[-1])
            out[tag + "__eps_y"] = worker.eps_scalar_t_fn(x, t, y=y)[..., ::2].numpy()  # every other column: fixture size
            out[tag + "__eps_noy"] = worker.eps_scalar_t_fn(x, t)[..., 1::2].numpy()
    x, _, _ = gi.collage_inputs(4, False)
    tiles, ov = split_wimg(x, 4)
    out["split4"] = tiles.numpy()
    out["split4_overlap"] = np.array(ov)
    out["merge4_avg"] = avg_merge_wimg(tiles * 2.0, ov, n=4).numpy()
    out["merge4_sum"] = avg_merge_wimg(tiles, ov, n=4, is_avg=False).numpy()
    save("collage", **out)


def golden_host():
    """Host-logic branches of the reference's sampler with an analytic denoiser (no networks, no SCG)."""
    out = {}
    for tag, cfg in gi.HOST_CASES.items():
        d = _create_diffusion(learn_sigma=cfg.get("learn_sigma", False), diffusion_steps=1000, noise_schedule="linear",
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
        out[tag] = np.stack(steps)
    save("host", **out)


def golden_sampler():
    """Short guided trajectories through the reference's own loops (seeded torch CPU RNG = shared noise stream)."""
    embed = build_ref_vae()
    out = {}
    for tag, cfg in gi.SAMPLER_CASES.items():
        model, _ = build_ref_dit(gi.DIT_CASES[cfg["dit"]])
        diffusion = create_diffusion(diffusion_steps=1000, noise_schedule="linear",
                                     timestep_respacing=cfg["respacing"])
        fn = partial(model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)
        kwargs = gi.sampler_model_kwargs(cfg)
        guidance = SimpleNamespace(**cfg["guidance"]) if cfg.get("guidance") else None
        loop = diffusion.ddim_sample_loop_progressive if cfg["ddim"] else diffusion.p_sample_loop_progressive
        extra = {"eta": cfg["eta"]} if cfg["ddim"] else {}
        diffusion.t_end = cfg.get("t_end", 0)
        torch.manual_seed(cfg["seed"])
        steps = []
        for o in loop(fn, cfg["shape"], model_kwargs=kwargs, device="cpu", embed_model=embed if cfg["scg"] else None,
                      scale_factor=gi.SCALE_FACTOR, guidance_kwargs=guidance,
                      scg_kwargs=dict(cfg["scg"]) if cfg["scg"] else None, t_end=cfg.get("t_end", 0), **extra):
            steps.append(o["sample"].numpy().copy())
        out[tag] = np.stack(steps)
        del model
    save("sampler", **out)


def golden_sampler_ext():
    """Classifier-guidance hook, replacement editing, DiffCollage workers, per-segment selection, final uint8 decode --
    all through the reference's own code."""
    import diff_collage as dc  # the reference's package
    from guided_diffusion.condition_functions import dc_model_fn
    from guided_diffusion.midi_util import decode_sample_for_midi

    embed = build_ref_vae()
    out = {}
    # spy on the losses so the goldens also carry, per step, the smallest relative margin between the best and the
    # runner-up candidate: a test may only insist on the same choice where the reference's own decision is not a
    # near-tie (random-weight models give nearly identical candidates)
    calls = []
    real_losses = dict(gd.LOSS_DICT)

    def spy(name):
        def f(gen, y):
            loss = real_losses[name](gen, y)
            calls.append((name, loss.clone()))
            return loss
        return f

    for name in list(gd.LOSS_DICT):
        gd.LOSS_DICT[name] = spy(name)

    def step_totals(cfg):
        """total_log_prob [N, B] of every argmax decision made since the last call (gaussian_diffusion.py:531-540)."""
        rules = cfg["rules"]
        n = cfg["scg"]["num_samples"]
        totals = []
        for i in range(0, len(calls), len(rules)):
            total = 0
            for name, loss in calls[i:i + len(rules)]:
                total = total + (-loss) * cfg["scg"].get(name, 1.0)
            totals.append(total.view(n, -1).numpy().copy())
        calls.clear()
        return totals

    for tag, cfg in gi.EXT_CASES.items():
        model, _ = build_ref_dit(gi.DIT_CASES[cfg["dit"]])
        diffusion = create_diffusion(timestep_respacing=cfg["respacing"])
        if cfg.get("dc"):
            def eps_fn(x, t, y=None, model=model):
                return model(x.permute(0, 1, 3, 2), t, y=y).permute(0, 1, 3, 2)
            if cfg["dc"]["type"] == "circle":
                worker = dc.CondIndCircle((4, 16, 128), eps_fn, cfg["dc"]["num_img"] + 1, overlap_size=64)
            else:
                worker = dc.CondIndSimple((4, 16, 128), eps_fn, cfg["dc"]["num_img"], overlap_size=64)
            assert (cfg["shape"][2], cfg["shape"][3]) == (worker.shape[2], worker.shape[1])
            fn = partial(dc_model_fn, model=worker.eps_scalar_t_fn, num_classes=3, class_cond=True, cfg=False, w=0.0)
        else:
            fn = partial(model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)
        g = dict(cfg["guidance"])
        if "dc" in g:
            g["dc"] = SimpleNamespace(**g["dc"])
        guidance = SimpleNamespace(**g)
        loop = diffusion.ddim_sample_loop_progressive if cfg["ddim"] else diffusion.p_sample_loop_progressive
        extra = {"eta": cfg["eta"]} if cfg["ddim"] else {}
        diffusion.t_end = 0
        torch.manual_seed(cfg["seed"])
        steps, totals, ndec = [], [], []
        calls.clear()
        for o in loop(fn, cfg["shape"], model_kwargs=gi.ext_model_kwargs(cfg), device="cpu", embed_model=embed,
                      scale_factor=gi.SCALE_FACTOR, guidance_kwargs=guidance, scg_kwargs=dict(cfg["scg"]),
                      cond_fn=gi.analytic_cond_fn if cfg.get("cond") else None,
                      edit_kwargs=gi.edit_inputs() if cfg.get("edit") else None, **extra):
            steps.append(o["sample"].numpy().copy())
            tt = step_totals(cfg)
            totals += tt
            ndec.append(len(tt))
        out[tag] = np.stack(steps)
        out[tag + "__totals"] = np.stack(totals)      # [decisions, N, B], in call order
        out[tag + "__ndec"] = np.array(ndec)           # decisions per step
        del model
    for name, fn_ in real_losses.items():
        gd.LOSS_DICT[name] = fn_
    out["midi_roll"] = decode_sample_for_midi(gi.vae_latents(), embed, gi.SCALE_FACTOR, threshold=-0.95).numpy()
    save("sampler_ext", **out)

def golden_flagship():
    """BASELINE config 3's step at B = 1 through the unmodified reference (gaussian_diffusion.py:881-976 -> :491-554):
    XL/8 denoiser, N = 16 candidates, 128 VAE tiles, pitch-histogram scores, argmax, teacher-forced at two timesteps."""
    cfg = gi.FLAGSHIP
    embed = build_ref_vae()
    model, _ = build_ref_dit(gi.DIT_CASES[cfg["dit"]])
    diffusion = create_diffusion(timestep_respacing=cfg["respacing"])
    cap = {}

    def spy_model(x, t, *a, **kw):
        o = model(x, t, *a, **kw)
        cap["eps" if x.shape[0] > 1 else "eps_b"] = o.clone()
        if x.shape[0] > 1:
            cap["cand"], cap["t_model"] = x.clone(), t.clone()
        return o

    real_decode, real_scg = gd._decode, diffusion.scg_sample
    real_func, real_loss = gd.FUNC_DICT["pitch_hist"], gd.LOSS_DICT["pitch_hist"]

    def spy_decode(z, em, scale_factor=1., threshold=False):
        cap["x0"] = z.clone()
        roll = real_decode(z, em, scale_factor=scale_factor, threshold=threshold)
        cap["roll"] = roll.clone()  # before the rules mask it in place
        return roll

    def spy_scg(m, t, mean_pred, g_coeff, *a, **k):
        cap["mean_pred"], cap["g"] = mean_pred.clone(), g_coeff.reshape(mean_pred.shape[0], -1)[:, 0].clone()
        return real_scg(m, t, mean_pred, g_coeff, *a, **k)

    def spy_func(roll):
        cap["hist"] = real_func(roll).clone()
        return cap["hist"]

    def spy_loss(gen, y):
        cap["loss"] = real_loss(gen, y).clone()
        return cap["loss"]

    gd._decode, diffusion.scg_sample = spy_decode, spy_scg
    gd.FUNC_DICT["pitch_hist"], gd.LOSS_DICT["pitch_hist"] = spy_func, spy_loss
    fn = partial(model_fn, model=spy_model, num_classes=3, class_cond=True, cfg=False, w=0.0)
    out = {}
    try:
        for case, c in cfg["cases"].items():
            x_t, t = gi.flagship_inputs(case, diffusion.alphas_cumprod)
            kwargs = {"y": torch.ones(1, dtype=torch.long), "rule": gi.rule_targets(1, 1024, cfg["rules"])}
            torch.manual_seed(c["seed"])
            diffusion.t_end = 0
            o = diffusion.ddim_sample(fn, x_t, t, model_kwargs=kwargs, eta=cfg["eta"], embed_model=embed,
                                      scale_factor=gi.SCALE_FACTOR, guidance_kwargs=SimpleNamespace(**cfg["guidance"]),
                                      scg_kwargs=dict(cfg["scg"]))
            total = (-cap["loss"] * cfg["scg"]["pitch_hist"]).view(cfg["N"], -1)
            out[case + "__t_model"] = cap["t_model"].numpy()          # original-scale timestep the denoiser saw
            out[case + "__eps_b"] = cap["eps_b"].numpy()              # the B = 1 pass of p_mean_variance
            out[case + "__pred_xstart"] = o["pred_xstart"].numpy()
            out[case + "__mean_pred"] = cap["mean_pred"].numpy()
            out[case + "__g"] = cap["g"].numpy()
            out[case + "__cand_sub"] = cap["cand"][:, :, ::2].numpy()  # every other time step: fixture size
            out[case + "__eps_sub"] = cap["eps"][:, :, ::2].numpy()
            out[case + "__x0_sub"] = cap["x0"][:, :, ::2].numpy()
            out[case + "__roll_sub"] = cap["roll"][:, 0, ::4, ::8].numpy()
            out[case + "__roll_stats"] = np.array([cap["roll"][:, 0].mean().item(), cap["roll"][:, 0].std().item(),
                                                   (cap["roll"][:, 0] > -0.95).float().mean().item()])
            out[case + "__hist"] = cap["hist"].numpy()
            out[case + "__total"] = total.numpy()
            out[case + "__max_ind"] = total.argmax(dim=0).numpy()
            out[case + "__sample"] = o["sample"].numpy()
            print(case, "totals", total.view(-1).numpy(), "argmax", int(total.argmax(dim=0)))
    finally:
        gd._decode, diffusion.scg_sample = real_decode, real_scg
        gd.FUNC_DICT["pitch_hist"], gd.LOSS_DICT["pitch_hist"] = real_func, real_loss
    save("flagship", **out)

def golden_chords():
    """The reference-owned parts of the chord rule (docs/CHORD_SPEC.md parts 1 and 3): the integer roll get_chords hands
    to piano_roll_to_chords, the note list piano_roll_to_pretty_midi extracts from it, the window vote and the degree
    tags.  music21 itself (part 2) is not available here: piano_roll_to_chords is intercepted at its entry."""
    import music_rule_guidance.music_rules as rmr
    import music_rule_guidance.piano_roll_to_chord as rch

    out = {}
    captured = []

    def fake_p2c(piano_roll, given_key=None, fs=100, window_size=1.28, return_key=False):
        captured.append(np.array(piano_roll, copy=True))
        return {"chords": torch.zeros(int(piano_roll.shape[-1] / fs / window_size), dtype=torch.long)}

    real = rmr.piano_roll_to_chords
    rmr.piano_roll_to_chords = fake_p2c
    try:
        roll = gi.chord_rolls()
        keep = roll.clone()
        res = rmr.get_chords(roll)
        out["velocities"] = np.stack(captured)
        out["get_chords_shape"] = np.array(res.shape)
        out["roll_after_ch0_sub"] = roll[:, 0, ::8, ::16].numpy()       # the in-place mask / threshold side effect
        out["roll_changed"] = int(not torch.equal(roll, keep))
    finally:
        rmr.piano_roll_to_chords = real
    for i, v in enumerate(out["velocities"]):
        pm = rch.piano_roll_to_pretty_midi(v.copy(), fs=100)
        notes = [(n.pitch, n.start, n.end, n.velocity) for n in pm.instruments[0].notes]
        out[f"notes_{i}"] = np.array(notes, dtype=np.float64).reshape(-1, 4)
    pm = rch.piano_roll_to_pretty_midi(out["velocities"][0][:, :160].copy(), fs=12.5)
    out["notes_fs12"] = np.array([(n.pitch, n.start, n.end, n.velocity) for n in pm.instruments[0].notes],
                                 dtype=np.float64).reshape(-1, 4)
    for tag, (chords, end_time, win, total) in gi.chord_vote_cases().items():
        out["vote_" + tag] = np.array(rch.get_longest_chords(chords, end_time, window_size=win, total_time=total))
    out["tags"] = np.array([rch.chord_tag_num(f) for f in gi.CHORD_FIGURES])
    save("chords", **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["schedule", "rules", "dit", "classifier", "vae", "vae_enc", "collage", "host", "sampler", "sampler_ext", "flagship", "chords"]
    for w in which:
        globals()["golden_" + w]()
