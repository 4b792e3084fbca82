"""Whole-step CUDA graphs (GaussianDiffusion.enable_cuda_graphs): a graph replay launches the same kernels in the same
order and torch's Philox generator advances identically, so a trajectory sampled with graphs must be BIT-IDENTICAL to
the eager one for the same seed -- for the plain ancestral sampler and for DDIM + SCG (two-lane VAE pipelining
included, i.e. cross-stream fork/join inside the capture)."""
from functools import partial
from types import SimpleNamespace

import pytest
import torch

import golden_inputs as gi
import gpu_util
from rule_guided_music_b200.guided_diffusion.condition_functions import model_fn
from rule_guided_music_b200.guided_diffusion.script_util import create_diffusion

pytestmark = pytest.mark.gpu
TARGET = [0.5, 0, 0, 0, 0.25, 0, 0, 0.25, 0, 0, 0, 0]


def _run(cuda, model, vae, graphs, ddim, scg, respacing="8", B=2, N=9):
    # N*B = 18 candidates = 144 VAE tiles = two chunks, so the decoder runs its two-lane (two-stream) pipeline
    diffusion = create_diffusion(timestep_respacing=respacing)
    diffusion.enable_cuda_graphs(graphs)
    fn = partial(model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)
    kwargs = {"y": torch.ones(B, dtype=torch.long, device=cuda),
              "rule": {"pitch_hist": torch.tensor([TARGET], device=cuda).repeat(B, 1),
                       "note_density": torch.tensor([[1, 1, 2, 3, 3, 2, 1, 1, 1, 1, 2, 3, 3, 2, 1, 1.]], device=cuda).repeat(B, 1)}}
    guidance = SimpleNamespace(schedule=False, t_start=750, t_end=0, interval=1, method="scg", step_size=1.0, nn=False)
    loop = diffusion.ddim_sample_loop_progressive if ddim else diffusion.p_sample_loop_progressive
    extra = {"eta": 1.0} if ddim else {}
    torch.manual_seed(77)
    steps = []
    for o in loop(fn, (B, 4, 128, 16), model_kwargs=kwargs, device=cuda, embed_model=vae if scg else None,
                  scale_factor=gi.SCALE_FACTOR, guidance_kwargs=guidance if scg else None,
                  scg_kwargs={"num_samples": N, "pitch_hist": 1.0, "note_density": 0.5} if scg else None, **extra):
        steps.append(o["sample"].clone())
    torch.cuda.synchronize()
    return steps, diffusion


@pytest.mark.parametrize("ddim,scg", [(False, False), (True, True), (False, True)])
def test_graph_replay_is_bit_identical_to_eager(cuda, ddim, scg):
    model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], cuda)
    vae, _ = gpu_util.native_vae(cuda)
    eager, _ = _run(cuda, model, vae, False, ddim, scg)
    graphed, diffusion = _run(cuda, model, vae, True, ddim, scg)
    assert len(eager) == len(graphed) == 8
    # steps 7..1 share one signature (first eager, then captured + replayed), step 0 is its own kind
    assert diffusion.captured_graphs() == 1
    for i, (a, b) in enumerate(zip(eager, graphed)):
        assert torch.equal(a, b), f"step {i}: max abs diff {(a - b).abs().max().item():.3e}"
    assert all(torch.isfinite(s).all() for s in graphed)


def test_user_rule_falls_back_to_eager(cuda):
    """A rule the kernels do not know is a Python callable on the roll: such steps must not be captured."""
    from rule_guided_music_b200.music_rule_guidance import rule_maps

    model, _ = gpu_util.native_dit(gi.DIT_CASES["small"], cuda)
    vae, _ = gpu_util.native_vae(cuda)
    rule_maps.FUNC_DICT["mean_velocity"] = lambda roll: roll[:, 0].mean(dim=(1, 2)).unsqueeze(-1)
    rule_maps.LOSS_DICT["mean_velocity"] = lambda gen, tgt: ((gen - tgt) ** 2).mean(dim=-1)
    try:
        diffusion = create_diffusion(timestep_respacing="4").enable_cuda_graphs(True)
        fn = partial(model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)
        kwargs = {"y": torch.ones(1, dtype=torch.long, device=cuda), "rule": {"mean_velocity": torch.zeros(1, 1, device=cuda)}}
        guidance = SimpleNamespace(schedule=False, t_start=750, t_end=0, interval=1, method="scg", step_size=1.0, nn=False)
        out = diffusion.p_sample_loop(fn, (1, 4, 128, 16), model_kwargs=kwargs, device=cuda, embed_model=vae,
                                      scale_factor=gi.SCALE_FACTOR, guidance_kwargs=guidance,
                                      scg_kwargs={"num_samples": 2, "mean_velocity": 1.0})
        assert torch.isfinite(out).all()
        assert not diffusion._graphs
    finally:
        del rule_maps.FUNC_DICT["mean_velocity"], rule_maps.LOSS_DICT["mean_velocity"]
