"""Chord-progression rule -- mirror of music_rule_guidance/music_rules.py:97-130 (`get_chords`) and
music_rule_guidance/piano_roll_to_chord.py of the reference.  docs/CHORD_SPEC.md is the written specification.

The rule has three parts:

  1. roll -> integer velocities -> note list           (music_rules.py:101-111, piano_roll_to_chord.py:166-256)
  2. note list -> key + one roman-numeral figure per chord   (music21 8.3.0: MIDI parse/quantise, chordify, key analysis,
                                                              romanNumeralFromChord; piano_roll_to_chord.py:25-67, 432)
  3. figures -> the figure with the longest overlap per 1.28 s window -> scale degree 0..7
                                                             (piano_roll_to_chord.py:70-120, 278-304)

Parts 1 and 3 are the reference's own code and are restated here, pinned against the reference on the CPU
(tests/golden/chords.npz).  Part 2 lives entirely in third-party packages the reference pins but does not vendor
(music21, and mido under its modified pretty_midi); they cannot be installed offline, so it is a pluggable ANALYZER:

    analyzer(notes, fs, given_key, total_time, need_key) -> (key_str or None, correlation, (figures, end_time))
        figures = [[duration_s, offset_s, roman-numeral figure], ...] of the chordified stream, end_time its length

`music21_analyzer` implements it with the reference's exact calls when music21 + pretty_midi are importable in the user's
environment; `get_chords` raises with that explanation when no analyzer is available.  The rule is a HOST rule (the
reference runs it in four worker processes on CPU copies of the roll, gaussian_diffusion.py:1363-1375): the sampler
treats it like any user callable -- materialised roll, eager step, Python loss.
"""
import numpy as np
import torch

MIN_PIANO, MAX_PIANO = 21, 108

KEY_DICT = {"D major": 0, "g minor": 1, "B- major": 2, "G major": 3, "d minor": 4, "c# minor": 5, "F major": 6,
            "E- major": 7, "e minor": 8, "f# minor": 9, "C major": 10, "F# major": 11, "g# minor": 12, "A major": 13,
            "a minor": 14, "B major": 15, "A- major": 16, "b- minor": 17, "E major": 18, "c minor": 19, "b minor": 20,
            "e- minor": 21, "f minor": 22, "C# major": 23, "no key": 24}


# ---- part 1: roll -> integer velocities -> notes ------------------------------------------------------------------------
def roll_to_velocities(piano_roll_batch):
    """music_rules.py:101-111: channel 0, piano range mask, `< -0.95 -> -1`, (x + 1) / 2 * 127, clamp [0, 127], truncate to
    int.  [B, C, 128, L] float -> int32 numpy [B, 128, L].  Like the reference, the mask and the threshold are written
    through to channel 0 of the caller's tensor."""
    pr = piano_roll_batch[:, :1, :, :]
    pr[:, :, :MIN_PIANO, :] = -1.
    pr[:, :, MAX_PIANO + 1:, :] = -1.
    pr[pr < -0.95] = -1.
    v = torch.clamp((pr + 1) / 2 * 127, min=0, max=127)
    return v[:, 0].detach().cpu().numpy().astype(np.intc)


def velocities_to_notes(piano_roll, fs=100):
    """piano_roll_to_chord.py:199-256 (`piano_roll_to_pretty_midi`, single-channel branch): int roll [128, frames] ->
    list of (pitch, start_s, end_s, velocity) in the reference's emission order (by note-off time, then pitch)."""
    piano_roll = np.array(piano_roll, dtype=np.intc, copy=True)
    notes_n, _ = piano_roll.shape
    background = piano_roll[:MIN_PIANO, :].max()
    piano_roll[piano_roll <= background] = 0
    piano_roll = np.pad(piano_roll, [(0, 0), (1, 1)], "constant")
    binary = piano_roll.copy()
    binary[binary != 0] = 1
    changes = np.nonzero(np.diff(binary).T)
    prev_vel = np.zeros(notes_n, dtype=int)
    on_time = np.zeros(notes_n)
    out = []
    for time, note in zip(*changes):
        velocity = piano_roll[note, time + 1]  # + 1: the padding column
        t = time / fs
        if velocity > 0:
            if prev_vel[note] == 0:
                on_time[note] = t
                prev_vel[note] = velocity
        else:
            out.append((int(note), float(on_time[note]), float(t), int(prev_vel[note])))
            prev_vel[note] = 0
    return out


# ---- part 3: figures -> window vote -> degree -------------------------------------------------------------------------
def get_longest_chords(chords, end_time, window_size=1.6, total_time=10.24):
    """piano_roll_to_chord.py:70-120: per window of `window_size` seconds the figure of the chord with the longest
    overlap ('null' for an empty window), padded with 'null' to int(total_time / window_size) entries.
    `chords`: [[duration_s, offset_s, figure], ...]."""
    result = []
    arr = np.array(chords)
    starts = arr[:, 1].astype(float)
    ends = starts + arr[:, 0].astype(float)
    current = 0
    while current < end_time:
        w0, w1 = current, current + window_size
        idx = np.where((starts < w1) & (ends > w0))[0]
        over = [[float(c[0]), float(c[1]), str(c[2])] for c in arr[idx]]
        if len(over) > 0:
            dur = [max(0, min(c[1] + c[0], w1) - max(c[1], w0)) for c in over]
            result.append(over[int(np.argmax(dur))][2])
        else:
            result.append("null")
        current += window_size
    while len(result) < int(total_time / window_size):
        result.append("null")
    return result


def chord_tag_num(figure):
    """piano_roll_to_chord.py:278-299: roman-numeral figure -> scale degree 1..7 (0 = none), by substring, in the
    reference's test order (VII, VI, IV, V, III, II, I)."""
    for tag, pats in ((7, ("VII", "vii")), (6, ("VI", "vi")), (4, ("IV", "iv")), (5, ("V", "v")), (3, ("III", "iii")),
                      (2, ("II", "ii")), (1, ("I", "i"))):
        if any(p in figure for p in pats):
            return tag
    return 0


# ---- part 2: the third-party analyzer ---------------------------------------------------------------------------------
def music21_analyzer(notes, fs, given_key=None, total_time=None, need_key=True):
    """The reference's music21 calls on the note list (piano_roll_to_chord.py:25-67, 142-163, 423-440): notes ->
    pretty_midi -> MIDI bytes -> music21 stream (its MIDI quantisation included) -> key analysis -> chordify ->
    romanNumeralFromChord.  Needs music21 (8.3.0 pinned by the reference) and a pretty_midi with `get_midi_data`
    (the reference's modified copy) in the environment."""
    try:
        import music21
        import pretty_midi
    except ImportError as e:  # pragma: no cover - not installable offline
        raise RuntimeError("chord_progression needs music21 and the reference's pretty_midi in the environment: " + str(e))
    pm = pretty_midi.PrettyMIDI()
    inst = pretty_midi.Instrument(program=0)
    for pitch, start, end, vel in notes:
        inst.notes.append(pretty_midi.Note(velocity=vel, pitch=pitch, start=start, end=end))
    pm.instruments.append(inst)
    stream = music21.midi.translate.midiStringToStream(pm.get_midi_data())
    key_str, corr = given_key, None
    if need_key:  # classify_keys_from_stream (:423-440); skipped when the key is given and not asked for (:329-333)
        try:
            fis = stream.analyze("key")
            key_str, corr = str(fis), fis.correlationCoefficient
        except Exception:
            return None, 0., ([], 0.)
    k = (given_key if given_key is not None else key_str).split(" ")[0]
    s_chords = stream.chordify().flatten()
    figures = []
    for c in s_chords.recurse().getElementsByClass(music21.chord.Chord):
        rn = music21.roman.romanNumeralFromChord(c, music21.key.Key(k))
        figures.append([float(c.seconds), float(c.offset / 120 * 60), str(rn.figure)])
    end_time = float(s_chords.highestTime) / 120 * 60
    return key_str, corr, (figures, end_time)


ANALYZER = None  # set to a callable to override music21 (tests, or a user's own harmonic analysis)


def piano_roll_to_chords(piano_roll, given_key=None, return_key=False, fs=100., window_size=1.28, analyzer=None):
    """piano_roll_to_chord.py:307-359 for one int roll [128, frames] -> {"chords": LongTensor[frames / fs / window]}
    (+ "key", "correlationCoefficient" like the reference when the key was analysed)."""
    analyzer = analyzer or ANALYZER or music21_analyzer
    total_time = piano_roll.shape[-1] / fs
    notes = velocities_to_notes(piano_roll, fs=fs)
    need_key = not (given_key is not None and not return_key)
    key_str, corr, analysed = analyzer(notes, fs, given_key, total_time, need_key)
    n_win = int(total_time / window_size)
    if key_str is None:
        return {"chords": torch.zeros(n_win, dtype=torch.long), "key": KEY_DICT["no key"], "correlationCoefficient": 0.}
    figures, end_time = analysed
    end_time = min(end_time, total_time)
    seq = get_longest_chords(figures, end_time, window_size=window_size, total_time=total_time) if figures else \
        ["null"] * n_win
    chords = torch.LongTensor([chord_tag_num(w) for w in seq])
    if given_key is not None and not return_key:
        return {"chords": chords}
    return {"chords": chords, "key": KEY_DICT.get(key_str, KEY_DICT["no key"]), "correlationCoefficient": corr}


def get_chords(piano_roll_batch, given_key=None, fs=100, window_size=1.28, return_key=False, analyzer=None):
    """music_rules.py:97-130: [B, C, 128, L] roll in [-1, 1] -> chord degrees [B, L / fs / window_size] int64 (squeezed
    to one row when B == 1, like the reference); with return_key also the keys and their correlation coefficients."""
    vel = roll_to_velocities(piano_roll_batch)
    outs = [piano_roll_to_chords(v, given_key=given_key, fs=fs, window_size=window_size, return_key=return_key,
                                 analyzer=analyzer) for v in vel]
    chords = torch.stack([o["chords"] for o in outs], dim=0)
    if chords.shape[0] == 1:
        chords = chords.squeeze(dim=0)
    if return_key:
        return chords, [o["key"] for o in outs], [o["correlationCoefficient"] for o in outs]
    return chords
