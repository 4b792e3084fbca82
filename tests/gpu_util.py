"""Helpers shared by the GPU parity tests: build the native models from the oracle's synthetic state dicts, and a
noise tape that makes the CUDA sampler consume the same CPU-generated Gaussian noise as the reference did."""
import contextlib

import torch

import golden_inputs as gi
from oracle import weights as ow


def native_dit(cfg, device):
    from rule_guided_music_b200.guided_diffusion.dit import DiTRotary

    w = cfg["weights"]
    H, W = cfg["input_size"]
    m = DiTRotary(input_size=[H, W], patch_size=w["patch"], in_channels=4, hidden_size=w["hidden"], depth=w["depth"],
                  num_heads=w["heads"], num_classes=w.get("num_classes", 3), learn_sigma=w.get("learn_sigma", False))
    sd = ow.make_dit_state_dict(**w)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing, missing
    m.to(device).eval()
    return m, sd


def native_vae(device, encoder=False):
    from rule_guided_music_b200.taming.models.klvae_pedal import AutoencoderKL

    sd = ow.make_vae_state_dict(seed=gi.VAE_SEED)
    if encoder:
        sd.update(ow.make_vae_encoder_state_dict(seed=gi.VAE_ENC_SEED))
    v = AutoencoderKL(ddconfig=ow.VAE_DDCONFIG, embed_dim=4)
    v.load_state_dict(sd, strict=False)
    v.to(device).eval()
    return v, sd


@contextlib.contextmanager
def cpu_noise_tape(module_th, seed):
    """Patch `randn` / `randn_like` of the torch module object `module_th` (the `th` a product module imported) so
    noise comes from torch's CPU generator -- the stream the reference consumed when the goldens were made."""
    real_randn, real_like = module_th.randn, module_th.randn_like
    torch.manual_seed(seed)

    def randn(*shape, device=None, dtype=None, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        return real_randn(*shape).to(device=device, dtype=dtype or torch.float32)

    def randn_like(x, **kw):
        return real_randn(*x.shape).to(device=x.device, dtype=x.dtype)

    module_th.randn, module_th.randn_like = randn, randn_like
    try:
        yield
    finally:
        module_th.randn, module_th.randn_like = real_randn, real_like


def rel_l2(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()
