"""The chord-progression rule's reference-owned parts (docs/CHORD_SPEC.md parts 1 and 3) against the UNMODIFIED reference
(tests/golden/chords.npz): the integer roll, the extracted note list, the per-window vote and the degree tags -- all
exact.  The music21 analysis in between is third-party and not installable offline: it is exercised here through a stub
analyzer, which also shows how a user plugs in their own (rule_maps.FUNC_DICT stays overridable as in the reference)."""
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
from rule_guided_music_b200.music_rule_guidance import chords
from rule_guided_music_b200.music_rule_guidance.rule_maps import FUNC_DICT, LOSS_DICT

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chords.npz"))


def test_velocities_and_side_effects_match_reference():
    roll = gi.chord_rolls()
    keep = roll.clone()
    vel = chords.roll_to_velocities(roll)
    np.testing.assert_array_equal(vel, GOLD["velocities"])
    np.testing.assert_array_equal(roll[:, 0, ::8, ::16].numpy(), GOLD["roll_after_ch0_sub"])
    assert (not torch.equal(roll, keep)) == bool(GOLD["roll_changed"])
    assert torch.equal(roll[:, 1:], keep[:, 1:])  # only channel 0 is touched


def test_note_extraction_matches_reference():
    for i, v in enumerate(GOLD["velocities"]):
        got = np.array(chords.velocities_to_notes(v, fs=100), dtype=np.float64).reshape(-1, 4)
        np.testing.assert_array_equal(got, GOLD[f"notes_{i}"])
    got = np.array(chords.velocities_to_notes(GOLD["velocities"][0][:, :160], fs=12.5), dtype=np.float64).reshape(-1, 4)
    np.testing.assert_array_equal(got, GOLD["notes_fs12"])


def test_window_vote_and_tags_match_reference():
    for tag, (ch, end_time, win, total) in gi.chord_vote_cases().items():
        assert chords.get_longest_chords(ch, end_time, window_size=win, total_time=total) == list(GOLD["vote_" + tag]), tag
    assert [chords.chord_tag_num(f) for f in gi.CHORD_FIGURES] == list(GOLD["tags"])


def _stub_analyzer(notes, fs, given_key, total_time, need_key):
    """One 'chord' per note-off group: figure cycles through I, IV, V by the lowest pitch class."""
    figs = []
    for pitch, start, end, vel in notes:
        figs.append([end - start, start, ("I", "IV", "V")[pitch % 3]])
    return ("C major" if need_key else given_key), 0.9, (figs, max((n[2] for n in notes), default=0.0))


def test_get_chords_with_a_plugged_analyzer():
    roll = gi.chord_rolls()
    out = chords.get_chords(roll.clone(), analyzer=_stub_analyzer)
    assert tuple(out.shape) == tuple(GOLD["get_chords_shape"]) and out.dtype == torch.int64
    assert set(out.unique().tolist()) <= {0, 1, 4, 5}
    one = chords.get_chords(roll[:1].clone(), analyzer=_stub_analyzer)
    assert one.shape == (8,)                                   # B == 1 is squeezed like the reference
    c, k, r = chords.get_chords(roll.clone(), analyzer=_stub_analyzer, return_key=True)
    assert k == [chords.KEY_DICT["C major"]] * 3 and r == [0.9] * 3
    # no key found -> zeros + "no key", like piano_roll_to_chord.py:336-341
    none = chords.piano_roll_to_chords(GOLD["velocities"][0], analyzer=lambda *a: (None, 0., ([], 0.)))
    assert none["key"] == 24 and not none["chords"].any()
    # the module-level hook and the registry
    chords.ANALYZER = _stub_analyzer
    try:
        assert torch.equal(FUNC_DICT["chord_progression"](roll.clone()), out)
        loss = LOSS_DICT["chord_progression"](out, out.clone())
        assert loss.shape == (3,) and not loss.any()
    finally:
        chords.ANALYZER = None


def test_get_chords_without_music21_says_so():
    try:
        import music21  # noqa: F401
        pytest.skip("music21 is installed here")
    except ImportError:
        pass
    with pytest.raises(RuntimeError, match="music21"):
        chords.get_chords(gi.chord_rolls()[:1].clone())
