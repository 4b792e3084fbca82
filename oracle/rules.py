"""Music rule programs and their losses, restated from music_rule_guidance/music_rules.py and rule_maps.py.

piano_like :23-26, total_pitch_class_histogram :29-43, note_density :46-83, note_density_class :86-94;
losses rule_maps.py:17-26; registries rule_maps.py:5-38.  Like the reference, the functions WRITE THROUGH their
input (pitch mask and the -0.95 threshold land in channel 0 of the roll that was passed in), so a later rule sees
what an earlier rule left behind -- the order dependence is part of the contract.
The chord rules (music21) are not restated: parity unpinned.
"""
from functools import partial

import torch
import torch.nn.functional as F

VERTICAL_ND_BOUNDS = [1.29, 2.7578125, 3.61, 4.4921875, 5.28125, 6.1171875, 7.22]
HORIZONTAL_ND_BOUNDS = [1.8, 2.6, 3.2, 3.6, 4.4, 4.8, 5.8]
MIN_PIANO, MAX_PIANO, OFF = 21, 108, -1


def piano_like(x):
    x[:, :, :MIN_PIANO, :] = OFF
    x[:, :, MAX_PIANO + 1:, :] = OFF
    return x


def total_pitch_class_histogram(roll):
    pr = piano_like(roll[:, :1, :, :])
    pr = ((pr + 1) / 2.0).squeeze(dim=1)
    per_pitch = pr.sum(dim=-1)                                            # [B,128]
    padded = torch.cat((per_pitch, torch.zeros(pr.shape[0], 4, device=pr.device)), dim=-1)  # [B,132]
    hist = padded.reshape(-1, 11, 12).permute(0, 2, 1).sum(dim=-1)        # fold pitch mod 12
    hist = hist / (hist.sum(dim=-1, keepdim=True) + 1e-12)
    return hist.squeeze(dim=0) if hist.shape[0] == 1 else hist


def note_density(roll, interval=128, quantize_factor=1, horizontal_scale=5):
    pr = roll[:, :1, :, :]
    B, L = pr.shape[0], pr.shape[-1]
    if quantize_factor != 1:
        pr = F.interpolate(pr, size=(128, L // quantize_factor), mode="nearest")
        interval = interval // quantize_factor
    pr = piano_like(pr)
    pr[pr < -0.95] = -1.0
    pr = (pr + 1) / 2.0
    pr[pr >= 1e-2] = 1.0
    pr[pr < 1e-2] = 0.0
    vertical = pr.sum(dim=2)                                  # notes per column [B,1,L]
    d = torch.diff(F.pad(pr, (1, 1), "constant"))
    d[d < 0] = 0
    horizontal = d.sum(dim=2)[:, :, :-1]                      # onsets per column
    horizontal[horizontal != 0.0] = 1
    v = vertical.reshape(B, 1, -1, interval).mean(dim=-1)
    h = horizontal.reshape(B, 1, -1, interval).sum(dim=-1) / horizontal_scale
    nd = torch.cat((v, h), dim=-1)
    return nd.squeeze() if B == 1 else nd.squeeze(dim=1)


def note_density_class(roll, interval=128, quantize_factor=1, horizontal_scale=1):
    vt = torch.tensor(VERTICAL_ND_BOUNDS).to(roll.device)
    hr = torch.tensor(HORIZONTAL_ND_BOUNDS).to(roll.device) / horizontal_scale
    nd = note_density(roll, interval=interval, quantize_factor=quantize_factor, horizontal_scale=horizontal_scale)
    n = nd.shape[-1]
    return torch.cat((torch.bucketize(nd[:, :n // 2], vt), torch.bucketize(nd[:, n // 2:], hr)), dim=-1)


def mse_loss_mean(gen, target):
    return F.mse_loss(gen.float(), target.float(), reduction="none").mean(dim=-1)


def zero_one_loss_mean(gen, target):
    return (target != gen).float().mean(dim=-1)


FUNC_DICT = {
    "pitch_hist": total_pitch_class_histogram,
    "note_density": note_density,
    "note_density_hr_1": partial(note_density, horizontal_scale=1.0),
    "note_density_hr_2": partial(note_density, horizontal_scale=2.0),
    "note_density_class": note_density_class,
    "note_density_pixel": partial(note_density, interval=16),
}
LOSS_DICT = {
    "pitch_hist": mse_loss_mean,
    "note_density": mse_loss_mean,
    "note_density_hr_1": mse_loss_mean,
    "note_density_hr_2": mse_loss_mean,
    "note_density_class": zero_one_loss_mean,
    "note_density_pixel": mse_loss_mean,
}
