"""DiTRotary forward on the B200 vs the reference's own outputs (tests/golden/dit.npz, fp32 CPU) and vs the oracle.

Tolerance: BASELINE.json states 1e-3 relative for sampled latents; the GEMM operands are fp16 (10-bit mantissa, the
class SURVEY.md section 0 measured at 6e-4 for one forward), residual stream / LN / softmax / accumulators fp32.
One forward must be within 1e-3 relative L2 of the fp32 reference, the 28-layer flagship included (measured on B200:
5.1e-4 ... 5.9e-4, exactly what tools/cpu_precision_study.py predicts from the operand roundings alone).
"""
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
import gpu_util
from oracle import dit as odit

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dit.npz"))
TOL = 1e-3


@pytest.mark.parametrize("tag", ["small", "small_hd64", "xl8"])
def test_dit_forward_matches_reference(cuda, tag, parity):
    cfg = gi.DIT_CASES[tag]
    model, _ = gpu_util.native_dit(cfg, cuda)
    x, t, y = gi.dit_inputs(cfg)
    out = model(x.to(cuda), t.to(cuda), y.to(cuda)).cpu()
    ref = torch.from_numpy(GOLD[tag])
    err = parity(f"DiT forward {tag} T=256", gpu_util.rel_l2(out, ref), TOL)
    assert err < TOL, (tag, err)
    if cfg.get("half_tile"):  # T = 128 tokens (diff-collage half tiles, condind_long.py:37)
        outh = model(x[:, :, :64].contiguous().to(cuda), t.to(cuda), y.to(cuda)).cpu()
        errh = parity(f"DiT forward {tag} T=128", gpu_util.rel_l2(outh, torch.from_numpy(GOLD[tag + "__half"])), TOL)
        assert errh < TOL, (tag, "half", errh)


def test_dit_batch_chunking_and_no_label(cuda, monkeypatch, parity):
    """Chunked execution (workspace reuse) gives the same result as one chunk; y=None skips the label embedding."""
    cfg = gi.DIT_CASES["small"]
    x, t, y = gi.dit_inputs(cfg)
    xs = torch.cat([x, x.flip(0), x * 0.5])  # 9 samples
    ts = torch.cat([t, t.flip(0), t])
    ys = torch.cat([y, y.flip(0), y])
    model, sd = gpu_util.native_dit(cfg, cuda)
    full = model(xs.to(cuda), ts.to(cuda), ys.to(cuda)).cpu()
    monkeypatch.setenv("RGM_DIT_CHUNK", "4")
    model2, _ = gpu_util.native_dit(cfg, cuda)
    chunked = model2(xs.to(cuda), ts.to(cuda), ys.to(cuda)).cpu()
    assert torch.equal(full, chunked)
    w = cfg["weights"]
    with torch.no_grad():
        ref = odit.dit_forward(sd, x, t, None, heads=w["heads"], patch=w["patch"])
    out = model(x.to(cuda), t.to(cuda), None).cpu()
    assert parity("DiT forward small, y=None, vs oracle", gpu_util.rel_l2(out, ref), TOL) < TOL


@pytest.mark.parametrize("H", [32, 96])
def test_dit_other_latent_lengths(cuda, H, parity):
    """The reference's DiTRotary accepts any latent length (dit.py:618-634); the native one takes 64 / 128 / 192 / 256
    tokens.  T = 64 and T = 192 (half-full last query tile in the attention kernel) against the oracle."""
    cfg = gi.DIT_CASES["small"]
    model, sd = gpu_util.native_dit(cfg, cuda)
    g = torch.Generator(device="cpu").manual_seed(H)
    x = torch.randn(3, 4, H, 16, generator=g)
    t = torch.tensor([900, 40, 512])
    y = torch.tensor([0, 2, 1])
    w = cfg["weights"]
    with torch.no_grad():
        ref = odit.dit_forward(sd, x, t, y, heads=w["heads"], patch=w["patch"])
    out = model(x.to(cuda), t.to(cuda), y.to(cuda)).cpu()
    assert parity(f"DiT forward small T={2 * H} vs oracle", gpu_util.rel_l2(out, ref), TOL) < TOL


def test_dit_rejects_cpu_and_bad_shapes(cuda):
    from rule_guided_music_b200 import _lib
    from rule_guided_music_b200.guided_diffusion.dit import DiT_models

    m = DiT_models["DiTRotary_B_8"](input_size=[128, 16], in_channels=4, num_classes=3, learn_sigma=False)
    with pytest.raises(_lib.RgmError):
        m.to("cpu")
    m.to(cuda)
    with pytest.raises(_lib.RgmError):
        m(torch.zeros(1, 4, 40, 16, device=cuda), torch.zeros(1, device=cuda), None)  # 80 tokens: unsupported
