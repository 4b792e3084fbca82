"""Config loading and the final latent -> piano-roll decode -- mirror of guided_diffusion/midi_util.py:26-64.
MIDI writing, plotting and evaluation reports are host-side I/O outside the sampling path."""
from types import SimpleNamespace

import torch
import yaml


def dict_to_obj(d):
    if isinstance(d, list):
        return [dict_to_obj(x) if isinstance(x, dict) else x for x in d]
    if not isinstance(d, dict):
        return d
    return SimpleNamespace(**{k: dict_to_obj(v) for k, v in d.items()})


def load_config(filename):
    """YAML -> nested SimpleNamespace (midi_util.py:34-39)."""
    with open(filename, "r") as f:
        return dict_to_obj(yaml.safe_load(f))


@torch.no_grad()
def decode_sample_for_midi(sample, embed_model=None, scale_factor=1., threshold=-0.95):
    """Final latents [B,4,H,16] -> uint8 piano roll [B,128,8H,3] in [0,127] (midi_util.py:42-64)."""
    if embed_model is not None and hasattr(embed_model, "decode_latents") and sample.shape[-2] > sample.shape[-1]:
        roll = embed_model.decode_latents(sample, scale_factor)
    else:
        sample = sample / scale_factor
        if embed_model is not None:
            h, w = sample.shape[-2], sample.shape[-1]
            if h > w:
                sample = sample.permute(0, 1, 3, 2)
            n = sample.shape[-1] // sample.shape[-2]
            if h >= w:
                sample = torch.concat(torch.chunk(sample, n, dim=-1), dim=0)
            sample = embed_model.decode(sample)
            if h >= w:
                sample = torch.concat(torch.chunk(sample, n, dim=0), dim=-1)
        roll = sample
    roll[roll <= threshold] = -1.
    roll = ((roll + 1) * 63.5).clamp(0, 127).to(torch.uint8)
    return roll.permute(0, 2, 3, 1).contiguous()
