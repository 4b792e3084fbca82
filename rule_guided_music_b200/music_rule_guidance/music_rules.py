"""Rule programs on the decoded piano roll -- mirror of music_rule_guidance/music_rules.py:23-94 of the reference,
evaluated by the warp-shuffle reduction kernels of csrc/rules.cu.

Signatures and results follow the reference: ``f(piano_roll [B,3,128,L] in [-1,1]) -> [B,K]`` (squeezed to ``[K]``
when B == 1), and -- like the reference -- channel 0 of the input is modified IN PLACE (piano mask, -0.95
threshold), so the order in which rules run matters.  CUDA tensors only; there is no CPU path here.
The chord rule (`get_chords`) is a host rule in chords.py: the reference's own parts are restated and pinned, the music21
analysis is a pluggable analyzer (docs/CHORD_SPEC.md).
"""
import torch

from .. import _lib

VERTICAL_ND_BOUNDS = [1.29, 2.7578125, 3.61, 4.4921875, 5.28125, 6.1171875, 7.22]
HORIZONTAL_ND_BOUNDS = [1.8, 2.6, 3.2, 3.6, 4.4, 4.8, 5.8]


def _check(roll):
    if not roll.is_cuda:
        raise _lib.RgmError("rule_guided_music_b200 rules run on the GPU only (no CPU path)")
    if roll.dim() != 4 or roll.shape[2] != 128 or roll.dtype != torch.float32:
        raise _lib.RgmError(f"rules expect an fp32 roll [B, C, 128, L], got {tuple(roll.shape)} {roll.dtype}")
    if not roll.is_contiguous():
        raise _lib.RgmError("rules write through their input like the reference; pass a contiguous roll")


def total_pitch_class_histogram(piano_roll):
    """music_rules.py:29-43."""
    _check(piano_roll)
    B, C, _, L = piano_roll.shape
    hist = torch.empty(B, 12, device=piano_roll.device, dtype=torch.float32)
    with torch.cuda.device(piano_roll.device):
        _lib.call("rgm_rule_pitch_hist", _lib.ptr(piano_roll), _lib.ptr(hist), B, C, L, _lib.stream_ptr())
    return hist.squeeze(0) if B == 1 else hist


def note_density(piano_roll, interval=128, quantize_factor=1, horizontal_scale=5):
    """music_rules.py:46-83.  quantize_factor != 1 (:59-61) first resamples the time axis with
    F.interpolate(mode="nearest") to L // quantize_factor columns -- for L divisible by the factor that is every
    quantize_factor-th column, and a NEW tensor, so the in-place threshold does not reach the caller's roll."""
    _check(piano_roll)
    if quantize_factor != 1:
        q = int(quantize_factor)
        if q < 1 or piano_roll.shape[-1] % q != 0 or interval % q != 0:
            raise _lib.RgmError("note_density: quantize_factor must divide the roll length and the interval")
        piano_roll = piano_roll[:, :1, :, ::q].contiguous()
        interval = interval // q
    B, C, _, L = piano_roll.shape
    out = torch.empty(B, 2 * (L // interval), device=piano_roll.device, dtype=torch.float32)
    with torch.cuda.device(piano_roll.device):
        _lib.call("rgm_rule_note_density", _lib.ptr(piano_roll), _lib.ptr(out), B, C, L, int(interval),
                  float(horizontal_scale), _lib.stream_ptr())
    return out.squeeze(0) if B == 1 else out


_BOUNDS = {}  # (device, horizontal_scale) -> (vertical, horizontal) class bounds on that device


def _class_bounds(device, horizontal_scale):
    """The bucket bounds as device tensors, built once per device: creating them per call is a pageable host-to-device
    copy, which synchronises and is illegal while a CUDA graph is being captured."""
    key = (str(device), float(horizontal_scale))
    b = _BOUNDS.get(key)
    if b is None:
        b = (torch.tensor(VERTICAL_ND_BOUNDS, device=device),
             torch.tensor(HORIZONTAL_ND_BOUNDS, device=device) / horizontal_scale)
        _BOUNDS[key] = b
    return b


def note_density_class(piano_roll, interval=128, quantize_factor=1, horizontal_scale=1):
    """music_rules.py:86-94 (bucketize on the fixed class bounds)."""
    nd = note_density(piano_roll, interval=interval, quantize_factor=quantize_factor,
                      horizontal_scale=horizontal_scale)
    vt, hr = _class_bounds(nd.device, horizontal_scale)
    n = nd.shape[-1]
    return torch.cat((torch.bucketize(nd[:, :n // 2].contiguous(), vt), torch.bucketize(nd[:, n // 2:].contiguous(), hr)),
                     dim=-1)


from .chords import get_chords  # noqa: E402,F401  (host rule: music_rules.py:97-130, see chords.py)
