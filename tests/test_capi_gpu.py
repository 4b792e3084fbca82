"""Round-2 boundary on the B200: one rgm_scg_step call (what a C host would make) against the Python-orchestrated
scg_sample of the same package -- bit for bit --, rgm_ddim_mean against the torch elementwise ops of ddim_sample,
rgm_coeff_tables on the device, and the workspace contract under CUDA graphs (buffers grow by retiring, never by
freeing: ADVICE round 1)."""
import ctypes
import os
from functools import partial
from types import SimpleNamespace

import numpy as np
import pytest
import torch

import golden_inputs as gi
import gpu_util
from oracle import rules as orules
from rule_guided_music_b200 import _lib
from rule_guided_music_b200.guided_diffusion.condition_functions import model_fn
from rule_guided_music_b200.guided_diffusion.script_util import create_diffusion
from rule_guided_music_b200.music_rule_guidance.rule_maps import FUNC_DICT

pytestmark = pytest.mark.gpu
GOLD_RULES = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rules.npz"))
TARGET = [0.5, 0, 0, 0, 0.25, 0, 0, 0.25, 0, 0, 0, 0]


def _tables_on_device(diffusion, cuda):
    T = diffusion.num_timesteps
    betas = np.ascontiguousarray(diffusion.betas, dtype=np.float64)
    tab = torch.empty(len(_lib.COEF_ROWS), T, device=cuda, dtype=torch.float32)
    _lib.call("rgm_coeff_tables", betas.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), T, _lib.ptr(tab),
              _lib.stream_ptr())
    return tab


def test_coeff_tables_on_device_equal_the_python_tables(cuda):
    d = create_diffusion(timestep_respacing="256")
    tab = _tables_on_device(d, cuda)
    mine = d._tables(cuda)
    for i, n in enumerate(_lib.COEF_ROWS):
        assert torch.allclose(tab[i], mine[n], rtol=1.2e-7, atol=0), n
        if "log" not in n:
            assert torch.equal(tab[i], mine[n]), n


def test_ddim_mean_equals_the_torch_ops_of_ddim_sample(cuda):
    d = create_diffusion(timestep_respacing="256")
    tab = _tables_on_device(d, cuda)
    g = torch.Generator(device="cpu").manual_seed(5)
    B = 3
    x = torch.randn(B, 4, 128, 16, generator=g).to(cuda)
    eps = torch.randn(B, 4, 128, 16, generator=g).to(cuda)
    t = torch.tensor([255, 120, 3], device=cuda)
    for eta, clip in ((1.0, True), (0.3, False)):
        out = d._ddim_sample_eager(lambda xx, tt, **k: eps, x, t, 255, clip_denoised=clip, eta=eta)  # scg None: + noise
        # re-derive mean_pred / sigma with the same torch ops (gaussian_diffusion.py ddim_sample)
        n = x.dim()
        x0 = d._predict_xstart_from_eps(x, t, eps)
        x0 = x0.clamp(-1, 1) if clip else x0
        e = d._predict_eps_from_xstart(x, t, x0)
        ab, abp = d._coef("alphas_cumprod", t, n), d._coef("alphas_cumprod_prev", t, n)
        sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
        mean = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * e
        px, mp, sg = torch.empty_like(x), torch.empty_like(x), torch.empty(B, device=cuda)
        _lib.call("rgm_ddim_mean", _lib.ptr(x), _lib.ptr(eps), _lib.ptr(tab), d.num_timesteps, _lib.ptr(t), eta,
                  int(clip), _lib.ptr(px), _lib.ptr(mp), _lib.ptr(sg), B, x[0].numel(), _lib.stream_ptr())
        assert torch.equal(px, x0) and torch.equal(px, out["pred_xstart"])
        assert torch.equal(sg, sigma.view(B))
        assert torch.equal(mp, mean)


@pytest.mark.parametrize("rules", [("pitch_hist",), ("note_density", "pitch_hist"), ("note_density_class",)])
def test_one_call_scg_step_equals_python_orchestration(cuda, rules):
    """rgm_scg_step == GaussianDiffusion.scg_sample of this package: chosen latents, indices and all scores, exactly."""
    cfg = gi.DIT_CASES["small"]
    model, _ = gpu_util.native_dit(cfg, cuda)
    vae, _ = gpu_util.native_vae(cuda)
    B, N, H = 2, 3, 64
    diffusion = create_diffusion(timestep_respacing="8")
    g = torch.Generator(device="cpu").manual_seed(17)
    mean = torch.randn(B, 4, H, 16, generator=g).to(cuda) * 0.7
    t = torch.full((B,), 5, device=cuda, dtype=torch.long)
    sigma = torch.tensor([0.31, 0.27], device=cuda)
    y = torch.tensor([1, 2], device=cuda)
    targets = {"pitch_hist": torch.tensor([TARGET], device=cuda).repeat(B, 1),
               "note_density": torch.tensor([[1., 2, 3, 1, 1, 2, 0.4, 0.2]], device=cuda).repeat(B, 1),
               "note_density_class": torch.tensor([[1., 2, 3, 1, 0, 2, 4, 1]], device=cuda).repeat(B, 1)}
    weights = {"pitch_hist": 1.0, "note_density": 0.5, "note_density_class": 2.0}
    mk = {"y": y, "rule": {n: targets[n] for n in rules}}
    scg = dict(num_samples=N, **{n: weights[n] for n in rules})
    fn = partial(model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)
    diffusion._trace = []
    torch.manual_seed(123)
    with torch.no_grad():
        want = diffusion.scg_sample(diffusion._wrap_model(fn), t, mean, sigma.view(B, 1, 1, 1).expand(mean.shape), vae,
                                    gi.SCALE_FACTOR, model_kwargs=mk, scg_kwargs=scg)
    want_total, want_idx = diffusion._trace[0]
    torch.manual_seed(123)
    noise = torch.randn(N, B, 4, H, 16, device=cuda)

    h = ctypes.c_void_p()
    _lib.call("rgm_scg_create", ctypes.byref(h), model._h, vae._h)
    try:
        _lib.call("rgm_scg_reserve", h, N, B, 4, H, 16)
        tab = diffusion._tables(cuda)
        t_model = torch.tensor(diffusion.timestep_map, device=cuda)[t].float()
        a = tab["sqrt_recip_alphas_cumprod"][t].contiguous()
        c = tab["sqrt_recipm1_alphas_cumprod"][t].contiguous()
        kind = {"pitch_hist": (0, 128, 5.0, 0), "note_density": (1, 128, 5.0, 0), "note_density_class": (2, 128, 1.0, 1)}
        specs = (_lib.RuleSpec * len(rules))()
        for i, n in enumerate(rules):
            k, interval, hs, loss = kind[n]
            specs[i] = _lib.RuleSpec(k, interval, hs, loss, weights[n], targets[n].data_ptr())
        out = torch.empty_like(mean)
        idx = torch.empty(B, device=cuda, dtype=torch.int64)
        scores = torch.empty(N * B, device=cuda)
        _lib.call("rgm_scg_step", h, _lib.ptr(mean), _lib.ptr(sigma), _lib.ptr(noise), _lib.ptr(t_model), _lib.ptr(y),
                  _lib.ptr(a), _lib.ptr(c), gi.SCALE_FACTOR, specs, len(rules), N, B, 4, H, 16, _lib.ptr(out),
                  _lib.ptr(idx), _lib.ptr(scores), _lib.stream_ptr())
        torch.cuda.synchronize()
    finally:
        _lib.call("rgm_scg_destroy", h)
    assert torch.equal(scores.view(N, B), want_total)
    assert torch.equal(idx, want_idx)
    assert torch.equal(out, want)


def test_scheduled_guidance_with_graphs_survives_workspace_growth(cuda):
    """ADVICE (round 1, high): with `schedule: True` the unguided step is captured at batch B while its workspaces are
    small; the first guided step then runs the denoiser on N*B candidates and GROWS them.  Replaying the earlier graph
    afterwards (interval = 2 alternates the two kinds of step) must still be correct: bit-identical to eager."""
    model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], cuda)
    vae, _ = gpu_util.native_vae(cuda)
    B, N = 2, 40  # 80 candidates (> B) and 640 tiles (several chunks): every workspace grows at the first guided step
    fn = partial(model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)
    kwargs = {"y": torch.ones(B, dtype=torch.long, device=cuda),
              "rule": {"pitch_hist": torch.tensor([TARGET], device=cuda).repeat(B, 1),
                       "note_density_class": torch.tensor([[1., 2, 3, 1, 1, 2, 0, 1, 1, 2, 3, 1, 1, 2, 0, 1]],
                                                          device=cuda).repeat(B, 1)}}
    guidance = SimpleNamespace(schedule=True, t_start=9, t_end=0, interval=2, method="scg", step_size=1.0, nn=False)

    def run(graphs):
        m2, _ = gpu_util.native_dit(gi.DIT_CASES["small"], cuda)  # fresh handles: workspaces start empty
        v2, _ = gpu_util.native_vae(cuda)
        f2 = partial(model_fn, model=m2, num_classes=3, class_cond=True, cfg=False, w=0.0)
        d = create_diffusion(timestep_respacing="12").enable_cuda_graphs(graphs)
        torch.manual_seed(7)
        steps = [o["sample"].clone() for o in d.p_sample_loop_progressive(
            f2, (B, 4, 128, 16), model_kwargs=kwargs, device=cuda, embed_model=v2, scale_factor=gi.SCALE_FACTOR,
            guidance_kwargs=guidance, scg_kwargs={"num_samples": N, "pitch_hist": 1.0, "note_density_class": 0.3})]
        torch.cuda.synchronize()
        return steps, d

    del model, vae, fn
    eager, _ = run(False)
    graphed, d = run(True)
    assert d.captured_graphs() >= 2  # the unguided kind and the guided kind (note_density_class included) were captured
    for i, (a, b) in enumerate(zip(eager, graphed)):
        assert torch.equal(a, b), f"step {i}: max abs diff {(a - b).abs().max().item():.3e}"


def test_reserve_then_capture_first_occurrence(cuda):
    """rgm_dit_reserve / rgm_vae_reserve pre-size the workspaces, so a step may be captured the first time it runs."""
    model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], cuda)
    vae, _ = gpu_util.native_vae(cuda)
    _lib.call("rgm_dit_reserve", model._h, 6, 128)
    _lib.call("rgm_vae_reserve", vae._h, 48)
    x = torch.randn(6, 4, 128, 16, device=cuda)
    t = torch.full((6,), 10.0, device=cuda)
    want = vae.decode_latents(model(x, t, None), 1.0, channels=1).clone()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = vae.decode_latents(model(x, t, None), 1.0, channels=1)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    # growing under capture is refused with an error, never silently reallocated
    big = torch.randn(600, 4, 128, 16, device=cuda)
    g2 = torch.cuda.CUDAGraph()
    with pytest.raises(_lib.RgmError):
        with torch.cuda.graph(g2):
            model(big, torch.full((600,), 10.0, device=cuda), None)


def test_note_density_quantize_factor(cuda):
    """note_density(quantize_factor != 1) (music_rules.py:59-61) against the reference's outputs; the caller's roll is
    left untouched (the reference resamples into a new tensor first)."""
    for q in gi.RULE_QUANT:
        r = gi.rule_rolls()["random"].to(cuda)
        keep = r.clone()
        got = FUNC_DICT["note_density"](r, quantize_factor=q).cpu().numpy()
        np.testing.assert_array_equal(got, GOLD_RULES[f"random__note_density_q{q}"])
        assert torch.equal(r, keep) == bool(GOLD_RULES[f"random__note_density_q{q}__input_untouched"])
        ref = orules.note_density(gi.rule_rolls()["random"], quantize_factor=q)
        assert torch.equal(torch.from_numpy(got), ref)


def test_rule_target_shape_is_checked(cuda):
    model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], cuda)
    vae, _ = gpu_util.native_vae(cuda)
    d = create_diffusion(timestep_respacing="8")
    fn = partial(model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)
    B = 2
    mean = torch.zeros(B, 4, 64, 16, device=cuda)
    mk = {"y": torch.ones(B, dtype=torch.long, device=cuda), "rule": {"pitch_hist": torch.zeros(1, 12, device=cuda)}}
    with pytest.raises(_lib.RgmError):
        d.scg_sample(d._wrap_model(fn), torch.full((B,), 3, device=cuda), mean, torch.ones_like(mean) * 0.1, vae, 1.0,
                     model_kwargs=mk, scg_kwargs={"num_samples": 2})


def test_c_host_runs_a_whole_step(cuda, tmp_path):
    """examples/scg_step_host.c: a C99 program (no Python, no torch) runs DiT -> rgm_ddim_mean -> rgm_scg_step on a B200."""
    import subprocess

    from test_capi_cpu import _build_c_host

    exe, env = _build_c_host(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "scg step ok" in out.stdout
