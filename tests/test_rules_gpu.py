"""Rule kernels and SCG selection vs the reference's outputs (tests/golden/rules.npz) and the oracle.
Counts / thresholds / argmax / gather: exact.  Histogram and MSE: fp32 with a different summation order, 1e-6."""
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
from oracle import rules as orules
from rule_guided_music_b200 import _lib
from rule_guided_music_b200.music_rule_guidance.rule_maps import FUNC_DICT, LOSS_DICT

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rules.npz"))


def test_rules_match_reference(cuda):
    for case, roll in gi.rule_rolls().items():
        for name in gi.RULE_NAMES:
            key = f"{case}__{name}"
            if key + "__raises" in GOLD.files:
                continue
            got = FUNC_DICT[name](roll.clone().to(cuda)).cpu().numpy()
            ref = GOLD[key]
            assert got.shape == ref.shape, key
            if name.startswith("note_density"):
                np.testing.assert_array_equal(got, ref, err_msg=key)  # integer counts divided identically
            else:
                np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-7, err_msg=key)


def test_rules_write_through_like_the_reference(cuda):
    r = gi.rule_rolls()["order"].to(cuda)
    FUNC_DICT["note_density"](r)
    np.testing.assert_allclose(FUNC_DICT["pitch_hist"](r).cpu().numpy(), GOLD["order__pitch_hist_after_nd"], rtol=2e-6,
                               atol=1e-7)
    np.testing.assert_array_equal(r[:, 0, 55:70, :16].cpu().numpy(), GOLD["order__roll_after"])
    # whole-tensor side effects equal the oracle's
    a = gi.rule_rolls()["random"]
    b = a.clone().to(cuda)
    orules.FUNC_DICT["pitch_hist"](a)
    FUNC_DICT["pitch_hist"](b)
    assert torch.equal(a, b.cpu())
    orules.FUNC_DICT["note_density"](a)
    FUNC_DICT["note_density"](b)
    assert torch.equal(a, b.cpu())


def test_single_channel_roll_and_long_rolls(cuda):
    """The fused sampler decodes channel 0 only; rolls longer than 1024 columns (diff-collage) work."""
    g = torch.Generator(device="cpu").manual_seed(3)
    r = torch.where(torch.rand(3, 1, 128, 2048, generator=g) < 0.05, torch.rand(3, 1, 128, 2048, generator=g) * 2 - 1,
                    -torch.ones(3, 1, 128, 2048))
    for name in ("pitch_hist", "note_density", "note_density_pixel", "note_density_class"):
        ref = orules.FUNC_DICT[name](r.clone())
        got = FUNC_DICT[name](r.clone().to(cuda)).cpu()
        if ref.dtype == torch.int64 or name.startswith("note_density"):
            assert torch.equal(got, ref), name
        else:
            assert torch.allclose(got, ref, rtol=2e-6, atol=1e-7), name


def test_loss_accum_and_select(cuda):
    N, B, K = 5, 7, 16
    g = torch.Generator(device="cpu").manual_seed(11)
    gen = torch.rand(N * B, K, generator=g)
    tgt = torch.rand(B, K, generator=g)
    total = torch.zeros(N * B, device=cuda)
    gen_d, tgt_d = gen.to(cuda), tgt.to(cuda)  # keep the device tensors alive across the asynchronous call
    _lib.call("rgm_rule_loss_accum", _lib.ptr(gen_d), _lib.ptr(tgt_d), _lib.ptr(total), N * B, B, K, 0,
              0.7, _lib.stream_ptr())
    ref = -orules.LOSS_DICT["pitch_hist"](gen, tgt.repeat(N, 1)) * 0.7
    assert torch.allclose(total.cpu(), ref, rtol=1e-6, atol=1e-8)
    gi_ = torch.randint(0, 3, (N * B, K), generator=g).float()
    ti_ = torch.randint(0, 3, (B, K), generator=g).float()
    tot2 = torch.zeros(N * B, device=cuda)
    gi_d, ti_d = gi_.to(cuda), ti_.to(cuda)
    _lib.call("rgm_rule_loss_accum", _lib.ptr(gi_d), _lib.ptr(ti_d), _lib.ptr(tot2), N * B, B, K, 1,
              1.0, _lib.stream_ptr())
    assert torch.equal(tot2.cpu(), -orules.LOSS_DICT["note_density_class"](gi_, ti_.repeat(N, 1)))
    assert LOSS_DICT["pitch_hist"] is not None


@pytest.mark.parametrize("N,B,elems", [(3, 2, 8), (16, 64, 8192), (64, 5, 8192), (1, 3, 100), (33, 2, 6)])
def test_select_first_max_with_ties(cuda, N, B, elems):
    g = torch.Generator(device="cpu").manual_seed(N + B)
    # scores that are multiples of 1/8 (zero-one losses): ties are the norm; argmax must take the FIRST maximum
    total = -(torch.randint(0, 4, (N, B), generator=g).float() / 8)
    cand = torch.randn(N, B, elems, generator=g)
    out = torch.empty(B, elems, device=cuda)
    idx = torch.empty(B, dtype=torch.int64, device=cuda)
    total_d, cand_d = total.to(cuda), cand.to(cuda)
    _lib.call("rgm_scg_select", _lib.ptr(total_d), _lib.ptr(cand_d), _lib.ptr(out), _lib.ptr(idx), N, B,
              elems, _lib.stream_ptr())
    ref_idx = total.argmax(dim=0)
    assert torch.equal(idx.cpu(), ref_idx)
    assert torch.equal(out.cpu(), cand[ref_idx, torch.arange(B)])
    if N >= 3:
        assert torch.tensor([[0., 1], [0, 1], [-1, 1]]).argmax(0).tolist() == [0, 0]


def test_eval_rule_loss_runs_native_rules_from_a_cpu_roll(cuda):
    """scripts/sample_rule.py:241-243 hands eval_rule_loss a CPU roll rebuilt from the uint8 MIDI array; the native
    rule kernels run where the targets live.  Values equal the oracle's rules + losses on the same roll."""
    from rule_guided_music_b200.guided_diffusion import midi_util

    roll = gi.rule_rolls()["random"]
    B = roll.shape[0]
    tp = torch.tensor([[0.5, 0, 0, 0, 0.25, 0, 0, 0.25, 0, 0, 0, 0]]).repeat(B, 1)
    df = midi_util.eval_rule_loss(roll.clone(), {"pitch_hist": tp.to(cuda)})
    assert len(df) == B and list(df.columns) == ["pitch_hist.target_rule", "pitch_hist.gen_rule", "pitch_hist.loss"]
    ref_gen = orules.FUNC_DICT["pitch_hist"](roll.clone())
    ref_loss = orules.LOSS_DICT["pitch_hist"](ref_gen, tp)
    np.testing.assert_allclose(np.array(df["pitch_hist.gen_rule"].tolist()), ref_gen.numpy(), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(np.array(df["pitch_hist.loss"].tolist()), ref_loss.numpy(), rtol=1e-5, atol=1e-9)
