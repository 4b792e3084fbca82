"""taming VAE decoder on the B200 vs the reference's outputs (tests/golden/vae.npz) and the oracle.

Activations are stored in fp16 between layers and conv operands are fp16 (fp32 accumulate); the reference runs its
convolutions in TF32 on a GPU and fp32 on the CPU.  Bar: 5e-3 relative L2 of the decoded roll and at most 0.5 % of
pixels on the other side of the rules' -0.95 note threshold."""
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
import gpu_util
from oracle import vae as ovae

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vae.npz"))


def test_decode_tiles_matches_reference(cuda, parity):
    vae, _ = gpu_util.native_vae(cuda)
    out = vae.decode(gi.vae_tiles().to(cuda)).cpu()
    ref = torch.from_numpy(GOLD["tiles"])
    assert out.shape == ref.shape
    err = parity("VAE decode, 2 tiles", gpu_util.rel_l2(out, ref), 5e-3)
    flips = parity("VAE decode, note-threshold flips", ((out >= -0.95) != (ref >= -0.95)).float().mean().item(), 5e-3,
                   "fraction")
    assert err < 5e-3, err
    assert flips < 5e-3, flips
    assert vae.gn_timeouts() == 0


def test_groupnorm_in_epilogue_equals_separate_pass(cuda, monkeypatch, parity):
    """conv1 of every ResnetBlock normalises its own output in its epilogue (default) vs the separate normalise pass
    (RGM_GN_EPI=0): same statistics, but the fused form normalises the fp32 result instead of its fp16 rounding, so the
    two decodes differ like two fp16 pipelines do (each is within 5e-3 of the reference; measured 3.5e-3 apart)."""
    vae, _ = gpu_util.native_vae(cuda)
    g = torch.Generator(device="cpu").manual_seed(5)
    lat = torch.randn(40, 4, 64, 16, generator=g).to(cuda)   # 160 tiles: two chunks on two lanes
    a = vae.decode_latents(lat, 1.1)
    assert vae.gn_timeouts() == 0
    a2 = vae.decode_latents(lat, 1.1)
    assert torch.equal(a, a2), "the in-epilogue statistics are summed in a fixed order: runs must be bit-identical"
    monkeypatch.setenv("RGM_GN_EPI", "0")
    vae0, _ = gpu_util.native_vae(cuda)
    b = vae0.decode_latents(lat, 1.1)
    assert parity("VAE decode, GroupNorm in the epilogue vs separate pass", gpu_util.rel_l2(a, b), 6e-3) < 6e-3


def test_decode_latents_layout_and_chunking(cuda, monkeypatch, parity):
    """_decode's tile-major re-tiling: roll[b, :, :, k*128:(k+1)*128] is tile k of sample b; chunked == unchunked."""
    vae, sd = gpu_util.native_vae(cuda)
    lat = gi.vae_latents()
    roll = vae.decode_latents(lat.to(cuda), gi.SCALE_FACTOR).cpu()
    ref_sub = torch.from_numpy(GOLD["decode_latents_sub4"])
    assert parity("_decode roll (sub-sampled)", gpu_util.rel_l2(roll[:, :, ::4, ::4], ref_sub), 5e-3) < 5e-3
    monkeypatch.setenv("RGM_VAE_CHUNK", "3")
    vae2, _ = gpu_util.native_vae(cuda)
    roll2 = vae2.decode_latents(lat.to(cuda), gi.SCALE_FACTOR).cpu()
    assert torch.equal(roll, roll2)
    ch0 = vae.decode_latents(lat.to(cuda), gi.SCALE_FACTOR, channels=1).cpu()
    assert torch.equal(ch0[:, 0], roll[:, 0])


def test_decode_many_tiles_vs_oracle(cuda, parity):
    """More tiles than one chunk, random latents, against the oracle run on the CPU here."""
    vae, sd = gpu_util.native_vae(cuda)
    g = torch.Generator(device="cpu").manual_seed(9)
    lat = torch.randn(3, 4, 48, 16, generator=g)
    with torch.no_grad():
        ref = ovae.decode_latents(sd, lat, 1.3)
    out = vae.decode_latents(lat.to(cuda), 1.3).cpu()
    assert gpu_util.rel_l2(out, ref) < 5e-3


# ---- encoder (scripts/edit.py: gaussian_diffusion._encode of the ground-truth roll) -------------------------------------
GOLD_ENC = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vae_enc.npz"))


def test_encode_matches_reference(cuda, parity):
    """Encoder + quant_conv (stride-2 implicit-GEMM Downsample, fp32 stem and tail) vs the reference's moments, and
    _encode's re-tiling vs the reference's latents.  Same precision policy and bar as the decoder (5e-3 rel L2)."""
    from rule_guided_music_b200.guided_diffusion.gaussian_diffusion import _encode

    vae, _ = gpu_util.native_vae(cuda, encoder=True)
    rolls = gi.vae_rolls()
    tiles = torch.cat(torch.chunk(rolls, rolls.shape[-1] // 128, dim=-1), dim=0)
    moments = vae.encode_save(tiles.to(cuda)).cpu()
    ref = torch.from_numpy(GOLD_ENC["moments"])
    assert moments.shape == ref.shape
    assert parity("VAE encoder moments", gpu_util.rel_l2(moments, ref), 5e-3) < 5e-3
    lat = _encode(rolls.to(cuda), vae, scale_factor=gi.SCALE_FACTOR).cpu()
    ref_lat = torch.from_numpy(GOLD_ENC["encode_latents"])
    assert lat.shape == ref_lat.shape
    assert parity("_encode latents", gpu_util.rel_l2(lat, ref_lat), 5e-3) < 5e-3
    post = vae.encode(tiles.to(cuda))
    assert torch.equal(post.mode().cpu(), moments[:, :4])
    assert post.sample().shape == (4, 4, 16, 16)


def test_encode_chunked_equals_unchunked_and_oracle(cuda, monkeypatch):
    vae, sd = gpu_util.native_vae(cuda, encoder=True)
    g = torch.Generator(device="cpu").manual_seed(31)
    x = (torch.rand(5, 3, 128, 128, generator=g) * 2 - 1)
    with torch.no_grad():
        ref = ovae.vae_encode(sd, x)
    out = vae.encode_save(x.to(cuda)).cpu()
    assert gpu_util.rel_l2(out, ref) < 5e-3
    monkeypatch.setenv("RGM_VAE_CHUNK", "2")
    vae2, _ = gpu_util.native_vae(cuda, encoder=True)
    assert torch.equal(vae2.encode_save(x.to(cuda)).cpu(), out)


def test_encode_without_encoder_weights_raises(cuda):
    from rule_guided_music_b200 import _lib

    vae, _ = gpu_util.native_vae(cuda)
    with pytest.raises(_lib.RgmError):
        vae.encode_save(torch.zeros(1, 3, 128, 128, device=cuda))
