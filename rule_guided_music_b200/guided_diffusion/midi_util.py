"""Config loading, the final latent -> piano-roll decode and the per-sample rule report -- mirror of
guided_diffusion/midi_util.py:26-64, 96-124.  MIDI writing and plotting are host-side I/O outside the sampling path."""
from types import SimpleNamespace

import torch
import yaml

from ..music_rule_guidance.rule_maps import FUNC_DICT, LOSS_DICT


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


@torch.no_grad()
def eval_rule_loss(generated_samples, target_rules):
    """Rule report of finished samples (midi_util.py:96-124; scripts/sample_rule.py:241-243): for every target rule
    the rule program is run on the decoded rolls [B,3,128,L] in [-1,1] and compared with the target by the rule's loss.
    Returns a pandas DataFrame with the reference's columns `<rule>.target_rule`, `<rule>.gen_rule`, `<rule>.loss`,
    one row per sample.  The script hands in a CPU tensor; the native rule kernels run on the GPU, so the rolls follow
    the target's device (the reference moves the target to the rolls instead)."""
    import pandas as pd

    results = {}
    B = generated_samples.shape[0]
    # one roll tensor for all rules: they write through their input (piano mask, -0.95 threshold), so a later rule
    # sees what an earlier one left behind, exactly like the reference's loop
    dev = next((t.device for t in target_rules.values() if t.is_cuda), generated_samples.device)
    rolls = generated_samples.to(dev).float()
    for name, target in target_rules.items():
        tl = target.tolist()
        results[name + ".target_rule"] = [tl] if B == 1 else tl
        gen = FUNC_DICT[name](rolls)
        loss = LOSS_DICT[name](gen, target.to(gen.device))
        gl = gen.tolist()
        results[name + ".gen_rule"] = [gl] if B == 1 else gl
        results[name + ".loss"] = loss.reshape(-1).tolist()
    return pd.DataFrame(results)
