"""BASELINE.json config-3 sizes (B=64, N=16: 1088 DiT samples, 1024 candidates = 8192 VAE tiles, rolls of 1024 columns)
checked through size-independent properties, because the CPU oracle cannot run this size in test time:

* the denoiser and the decoder are per-sample functions: a sample's output must not depend on what else is in the batch
  or on where the library cuts its chunks (bit-identical);
* the rule kernels on the full roll against the oracle's rules (integer work exact, histogram 2e-6);
* loss accumulation and the first-max argmax + gather at N=16, B=64 against torch, ties included (bit-exact);
* one full SCG step through the public API: finite, and the recorded choice is the first maximum of the recorded scores.
"""
from functools import partial
from types import SimpleNamespace

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
B, N = 64, 16
TARGET = [0.5, 0, 0, 0, 0.25, 0, 0, 0.25, 0, 0, 0, 0]


@pytest.fixture(scope="module")
def models(cuda):
    model, _ = gpu_util.native_dit(gi.DIT_CASES["xl8"], cuda)
    vae, _ = gpu_util.native_vae(cuda)
    return model, vae


def _first_max(total):
    top = total.max(dim=0).values
    return ((total == top).cumsum(0) == 0).sum(0)


def test_denoiser_is_batch_and_chunk_independent(cuda, models):
    model, _ = models
    g = torch.Generator(device="cpu").manual_seed(41)
    x = torch.randn(B + N * B, 4, 128, 16, generator=g).to(cuda)
    t = torch.full((x.shape[0],), 613, device=cuda)
    y = torch.ones(x.shape[0], dtype=torch.long, device=cuda)
    full = model(x, t, y)                       # 1088 samples: 4 chunks of 256 + one of 64
    assert torch.isfinite(full).all()
    for lo, hi in ((0, 8), (250, 262), (1080, 1088)):   # inside a chunk, across a chunk boundary, the ragged tail
        part = model(x[lo:hi].contiguous(), t[lo:hi], y[lo:hi])
        assert torch.equal(part, full[lo:hi]), (lo, hi)


def test_decoder_rules_and_selection_at_full_size(cuda, models):
    _, vae = models
    g = torch.Generator(device="cpu").manual_seed(42)
    x0 = (torch.randn(N * B, 4, 128, 16, generator=g) * gi.SCALE_FACTOR).to(cuda)
    roll = vae.decode_latents(x0, gi.SCALE_FACTOR, channels=1)      # [1024, 1, 128, 1024]: 64 chunks of 128 tiles
    assert roll.shape == (N * B, 1, 128, 1024) and torch.isfinite(roll).all()
    for i in (0, 517, 1023):                                        # alone, the candidate's 8 tiles form one chunk
        single = vae.decode_latents(x0[i:i + 1].contiguous(), gi.SCALE_FACTOR, channels=1)
        assert torch.equal(single[0], roll[i]), i
    # rules on the full roll vs the oracle (CPU) on a slice of candidates and, for the histogram, on all of them
    cpu_roll = roll.cpu()
    hist = FUNC_DICT["pitch_hist"](roll)
    ref_hist = orules.FUNC_DICT["pitch_hist"](cpu_roll)
    torch.testing.assert_close(hist.cpu(), ref_hist, rtol=2e-6, atol=1e-7)
    assert torch.equal(roll.cpu(), cpu_roll)                        # the same in-place piano mask
    nd = FUNC_DICT["note_density"](roll)
    ref_nd = orules.FUNC_DICT["note_density"](cpu_roll[:96])
    assert torch.equal(nd[:96].cpu(), ref_nd)
    # loss accumulation + selection, N=16, B=64
    target = torch.tensor([TARGET], device=cuda).repeat(B, 1)
    total = torch.zeros(N * B, device=cuda)
    _lib.call("rgm_rule_loss_accum", _lib.ptr(hist), _lib.ptr(target), _lib.ptr(total), N * B, B, 12, 0, 1.0,
              _lib.stream_ptr())
    ref_total = -((hist - target.repeat(N, 1)) ** 2).mean(dim=-1)
    torch.testing.assert_close(total, ref_total, rtol=1e-5, atol=1e-9)
    total = total.view(N, B).clone()
    total[3, 5] = total[9, 5] = total[:, 5].max() + 1.0             # a tie: index 3 must win
    total[:, 7] = 0.25                                              # all equal: index 0
    cand = x0.view(N * B, -1)
    out = torch.empty(B, cand.shape[1], device=cuda)
    idx = torch.empty(B, dtype=torch.int64, device=cuda)
    _lib.call("rgm_scg_select", _lib.ptr(total.contiguous()), _lib.ptr(cand), _lib.ptr(out), _lib.ptr(idx), N, B,
              cand.shape[1], _lib.stream_ptr())
    want = _first_max(total)
    assert torch.equal(idx, want) and idx[5] == 3 and idx[7] == 0
    assert torch.equal(out, cand.view(N, B, -1)[want, torch.arange(B, device=cuda)])


def test_full_step_through_the_public_api(cuda, models):
    model, vae = models
    diffusion = create_diffusion(timestep_respacing="256")
    diffusion._trace = []
    fn = partial(model_fn, model=model, num_classes=3, class_cond=True, cfg=False, w=0.0)
    kwargs = {"y": torch.ones(B, dtype=torch.long, device=cuda),
              "rule": {"pitch_hist": torch.tensor([TARGET], device=cuda).repeat(B, 1)}}
    guidance = SimpleNamespace(schedule=False, t_start=750, t_end=0, interval=1, method="scg", step_size=1.0, nn=False)
    torch.manual_seed(5)
    x = torch.randn(B, 4, 128, 16, device=cuda)
    t = torch.full((B,), 200, device=cuda, dtype=torch.long)
    out = diffusion.ddim_sample(fn, x, t, model_kwargs=kwargs, eta=1.0, embed_model=vae, scale_factor=gi.SCALE_FACTOR,
                                guidance_kwargs=guidance, scg_kwargs={"num_samples": N, "pitch_hist": 1.0}, _t_host=200)
    assert out["sample"].shape == (B, 4, 128, 16) and torch.isfinite(out["sample"]).all()
    (total, idx), = diffusion._trace
    assert total.shape == (N, B) and (total <= 0).all()             # -MSE
    assert torch.equal(idx, _first_max(total))
    assert len(torch.unique(idx)) > 1                               # the guidance really chooses
