"""FUNC_DICT / LOSS_DICT registries -- mirror of music_rule_guidance/rule_maps.py:5-38.

Users extend both dicts exactly as with the reference (README.md:65-68).  The sampler recognises the entries below
by identity and runs them (and their losses) through the fused CUDA path; any other callable is simply called on the
materialised roll tensor.
"""
from functools import partial

import torch.nn.functional as F

from .music_rules import get_chords, note_density, note_density_class, total_pitch_class_histogram


def mse_loss_mean(gen_rule, y_):
    """rule_maps.py:17-18."""
    return F.mse_loss(gen_rule.float(), y_.float(), reduction="none").mean(dim=-1)


def zero_one_loss_mean(gen_rule, y_):
    """rule_maps.py:21-22."""
    return (y_ != gen_rule).float().mean(dim=-1)


FUNC_DICT = {
    "pitch_hist": total_pitch_class_histogram,
    "note_density": note_density,
    "note_density_hr_1": partial(note_density, horizontal_scale=1.0),
    "note_density_hr_2": partial(note_density, horizontal_scale=2.0),
    "note_density_class": note_density_class,
    "note_density_pixel": partial(note_density, interval=16),
    "chord_progression": get_chords,
    "chord_progression_pixel": partial(get_chords, fs=12.5),
}

LOSS_DICT = {
    "pitch_hist": mse_loss_mean,
    "note_density": mse_loss_mean,
    "note_density_hr_1": mse_loss_mean,
    "note_density_hr_2": mse_loss_mean,
    "note_density_class": zero_one_loss_mean,
    "note_density_pixel": mse_loss_mean,
    "chord_progression": zero_one_loss_mean,
    "chord_progression_pixel": zero_one_loss_mean,
}

# losses the fused scoring kernel implements, by identity: callable -> kind of rgm_rule_loss_accum
NATIVE_LOSS_KIND = {mse_loss_mean: 0, zero_one_loss_mean: 1}
