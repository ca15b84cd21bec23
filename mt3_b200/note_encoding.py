"""Notes -> event tokens: the inverse of mt3_b200.note_decoding (the label side of the token contract).

The reference produces its target token streams with
    note_sequences.note_sequence_to_onsets[_and_offsets[_and_programs]]   (note_sequences.py:141-200)
      -> run_length_encoding.encode_and_index_events                      (run_length_encoding.py:62-168)
      -> extract_target_sequence_with_indices (+ tie section)             (run_length_encoding.py:171-193)
      -> run_length_encode_shifts / remove_redundant_state_changes        (run_length_encoding.py:196-289)
and evaluates with the same tokens.  This module restates that direction over the plain-Python NoteSequence of
note_decoding, so that the stitch can be property-tested against its own inverse (notes -> tokens -> notes) and a
reference transcription can be turned into the token streams the model is scored on.  Host-side list / numpy code.

Differences in form, not behaviour: the per-frame index arrays are computed with two `searchsorted` calls instead of the
reference's incremental fill loop (same float comparisons: frame_time < step / steps_per_second), and the tf.data
preprocessors are plain functions over integer sequences.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import event_codec, vocabularies
from .note_decoding import DEFAULT_NOTE_DURATION, DEFAULT_VELOCITY, Note, NoteSequence, assign_instruments

Event = event_codec.Event


# ---------------------------------------------------------------------------------------------
# NoteSequence utilities (note_sequences.py:43-127)
# ---------------------------------------------------------------------------------------------
def extract_track(ns: NoteSequence, program: int, is_drum: bool) -> NoteSequence:
    track = NoteSequence(ticks_per_quarter=220)
    track.notes = [dataclasses.replace(n) for n in ns.notes if n.program == program and n.is_drum == is_drum]
    track.total_time = max((n.end_time for n in track.notes), default=0.0)
    return track


def trim_overlapping_notes(ns: NoteSequence) -> NoteSequence:
    """A copy in which a note is cut where the next note of the same (pitch, program, is_drum) starts; notes left with
    no duration are dropped (note_sequences.py:53-70)."""
    out = NoteSequence(ticks_per_quarter=ns.ticks_per_quarter, total_time=ns.total_time)
    out.notes = [dataclasses.replace(n) for n in ns.notes]
    by_channel: Dict[Tuple[int, int, bool], List[Note]] = {}
    for n in out.notes:
        by_channel.setdefault((n.pitch, n.program, n.is_drum), []).append(n)
    for notes in by_channel.values():
        notes.sort(key=lambda n: n.start_time)                  # stable, like sorted()
        for prev, nxt in zip(notes, notes[1:]):
            if prev.end_time > nxt.start_time:
                prev.end_time = nxt.start_time
    out.notes = [n for n in out.notes if n.start_time < n.end_time]
    return out


def validate_note_sequence(ns: NoteSequence) -> None:
    for n in ns.notes:
        if n.start_time >= n.end_time:
            raise ValueError('note has start time >= end time: %f >= %f' % (n.start_time, n.end_time))
        if n.velocity == 0:
            raise ValueError('note has zero velocity')


def note_arrays_to_note_sequence(onset_times: Sequence[float], pitches: Sequence[int],
                                 offset_times: Optional[Sequence[float]] = None, velocities: Optional[Sequence[int]] = None,
                                 programs: Optional[Sequence[int]] = None, is_drums: Optional[Sequence[bool]] = None) -> NoteSequence:
    """Arrays -> NoteSequence with the reference's defaults: 10 ms notes, velocity 100, program 0 (note_sequences.py:99-127)."""
    ns = NoteSequence(ticks_per_quarter=220)
    for i, (onset, pitch) in enumerate(zip(onset_times, pitches)):
        ns.add(onset, onset + DEFAULT_NOTE_DURATION if offset_times is None else offset_times[i], pitch,
               DEFAULT_VELOCITY if velocities is None else velocities[i], program=0 if programs is None else programs[i],
               is_drum=False if is_drums is None else is_drums[i])
    assign_instruments(ns)
    return ns


# ---------------------------------------------------------------------------------------------
# Notes -> timed event data (note_sequences.py:130-200)
# ---------------------------------------------------------------------------------------------
@dataclasses.dataclass
class NoteEventData:
    pitch: int
    velocity: Optional[int] = None
    program: Optional[int] = None
    is_drum: Optional[bool] = None
    instrument: Optional[int] = None


def note_sequence_to_onsets(ns: NoteSequence) -> Tuple[List[float], List[NoteEventData]]:
    """Onsets only.  Pitch order breaks ties in the stable time sort that follows (note_sequences.py:141-148)."""
    notes = sorted(ns.notes, key=lambda n: n.pitch)
    return [n.start_time for n in notes], [NoteEventData(pitch=n.pitch) for n in notes]


def note_sequence_to_onsets_and_offsets(ns: NoteSequence) -> Tuple[List[float], List[NoteEventData]]:
    """All offsets (velocity 0), then all onsets: at equal times an offset sorts before an onset (note_sequences.py:151-173)."""
    notes = sorted(ns.notes, key=lambda n: n.pitch)
    times = [n.end_time for n in notes] + [n.start_time for n in notes]
    values = ([NoteEventData(pitch=n.pitch, velocity=0) for n in notes] +
              [NoteEventData(pitch=n.pitch, velocity=n.velocity) for n in notes])
    return times, values


def note_sequence_to_onsets_and_offsets_and_programs(ns: NoteSequence) -> Tuple[List[float], List[NoteEventData]]:
    """As above with programs; drums have no offsets and sort after the melodic notes (note_sequences.py:176-200)."""
    notes = sorted(ns.notes, key=lambda n: (n.is_drum, n.program, n.pitch))
    melodic = [n for n in notes if not n.is_drum]
    times = [n.end_time for n in melodic] + [n.start_time for n in notes]
    values = ([NoteEventData(pitch=n.pitch, velocity=0, program=n.program, is_drum=False) for n in melodic] +
              [NoteEventData(pitch=n.pitch, velocity=n.velocity, program=n.program, is_drum=n.is_drum) for n in notes])
    return times, values


# ---------------------------------------------------------------------------------------------
# Event data -> events, with the state needed for tie sections (note_sequences.py:203-259)
# ---------------------------------------------------------------------------------------------
@dataclasses.dataclass
class NoteEncodingState:
    """(pitch, program) -> velocity bin of the last event seen for it; a non-zero bin means the note is sounding."""
    active_pitches: Dict[Tuple[int, int], int] = dataclasses.field(default_factory=dict)


def note_event_data_to_events(state: Optional[NoteEncodingState], value: NoteEventData, codec: event_codec.Codec) -> List[Event]:
    if value.velocity is None:                                   # onsets only
        return [Event('pitch', value.pitch)]
    vbin = vocabularies.velocity_to_bin(value.velocity, vocabularies.num_velocity_bins_from_codec(codec))
    if value.program is None:                                    # onsets + offsets + velocities, single track
        if state is not None:
            state.active_pitches[(value.pitch, 0)] = vbin
        return [Event('velocity', vbin), Event('pitch', value.pitch)]
    if value.is_drum:                                            # drums have their own pitch vocabulary and no state
        return [Event('velocity', vbin), Event('drum', value.pitch)]
    if state is not None:
        state.active_pitches[(value.pitch, int(value.program))] = vbin
    return [Event('program', value.program), Event('velocity', vbin), Event('pitch', value.pitch)]


def note_encoding_state_to_events(state: NoteEncodingState) -> List[Event]:
    """The tie section: (program, pitch) of every sounding note in (program, pitch) order, closed by a tie event."""
    events: List[Event] = []
    for pitch, program in sorted(state.active_pitches, key=lambda k: (k[1], k[0])):
        if state.active_pitches[(pitch, program)]:
            events += [Event('program', program), Event('pitch', pitch)]
    events.append(Event('tie', 0))
    return events


# ---------------------------------------------------------------------------------------------
# Timed events -> token stream indexed to audio frames (run_length_encoding.py:62-168)
# ---------------------------------------------------------------------------------------------
def encode_and_index_events(state, event_times: Sequence[float], event_values: Sequence, encode_event_fn: Callable,
                            codec: event_codec.Codec, frame_times: Sequence[float],
                            encoding_state_to_events_fn: Optional[Callable] = None):
    """Returns (events, event_start_indices, event_end_indices, state_events, state_event_indices).

    `events`: the event tokens in time order with every time step written as ONE single-step shift token (run-length
    encoded later), continued with shifts until the step after the last frame.  For audio frame f at step
    k = #{c >= 1 : c / steps_per_second <= frame_times[f]}, `event_start_indices[f]` is the position right after the
    k-th shift, `event_end_indices[f] = event_start_indices[f + 1]` (len(events) for the last frame), and
    `state_event_indices[f]` points at the state dump (tie section) valid at that step."""
    frame_times = np.asarray(frame_times, np.float64)
    if frame_times.size == 0:
        raise IndexError('encode_and_index_events needs at least one frame time')
    sps = codec.steps_per_second
    order = np.argsort(np.asarray(event_times, np.float64), kind='stable')
    steps = [round(float(event_times[i]) * sps) for i in order]          # Python round(): half to even, like the reference
    shift = codec.encode_event(Event('shift', 1))

    events: List[int] = []
    state_events: List[int] = []
    n_tok, n_state = [], []                                               # tokens / state tokens emitted per event
    cur = 0
    for step, i in zip(steps, order):
        if step > cur:
            events.extend([shift] * (step - cur))
            cur = step
        if encoding_state_to_events_fn is not None:                       # the state BEFORE the event
            dump = [codec.encode_event(e) for e in encoding_state_to_events_fn(state)]
            state_events.extend(dump)
            n_state.append(len(dump))
        toks = [codec.encode_event(e) for e in encode_event_fn(state, event_values[i], codec)]
        events.extend(toks)
        n_tok.append(len(toks))
    last_event_step = cur
    while cur / sps <= frame_times[-1]:        # not strict: a step that coincides with a frame start still needs its shift
        events.append(shift)
        cur += 1

    # frame f belongs to step k_f: the first c with frame_time < c / sps is k_f + 1
    thresholds = np.arange(1, cur + 1, dtype=np.int64) / sps
    k = np.searchsorted(thresholds, frame_times, side='right')
    ev_steps = np.asarray(steps, np.int64)
    before = np.searchsorted(ev_steps, k, side='left')                    # events strictly before step k_f
    tok_cum = np.concatenate([[0], np.cumsum(np.asarray(n_tok, np.int64))])
    start = np.where(k == 0, 0, k + tok_cum[before]).astype(np.int64)
    end = np.concatenate([start[1:], [len(events)]]).astype(np.int64)
    if encoding_state_to_events_fn is not None:
        st_cum = np.concatenate([[0], np.cumsum(np.asarray(n_state, np.int64))])
        before_s = np.searchsorted(ev_steps, np.minimum(k, last_event_step), side='left')   # not advanced by the trailing shifts
        state_idx = np.where(k == 0, 0, st_cum[before_s]).astype(np.int64)
    else:
        state_idx = np.zeros(frame_times.size, np.int64)
    return (np.asarray(events, np.int64), start, end, np.asarray(state_events, np.int64), state_idx)


# ---------------------------------------------------------------------------------------------
# Per-segment targets (run_length_encoding.py:171-289, :292-368)
# ---------------------------------------------------------------------------------------------
def extract_target_sequence_with_indices(targets: Sequence[int], event_start_indices: Sequence[int],
                                         event_end_indices: Sequence[int], state_events: Optional[Sequence[int]] = None,
                                         state_event_indices: Optional[Sequence[int]] = None,
                                         state_events_end_token: Optional[int] = None) -> np.ndarray:
    """Targets of one segment: the tokens between the first frame's start index and the last frame's end index, preceded
    (when `state_events_end_token` is given) by the state dump valid at the first frame, up to and including its end token."""
    out = np.asarray(targets)[int(event_start_indices[0]):int(event_end_indices[-1])]
    if state_events_end_token is not None:
        state_events = np.asarray(state_events)
        s = int(state_event_indices[0])
        e = s + 1
        while state_events[e - 1] != state_events_end_token:
            e += 1
        out = np.concatenate([state_events[s:e], out])
    return out


def remove_redundant_state_changes(tokens: Sequence[int], codec: event_codec.Codec,
                                   state_change_event_types: Sequence[str] = ()) -> List[int]:
    """Drop a state-change token (e.g. velocity, program) that repeats the current value of its type."""
    ranges = [codec.event_type_range(t) for t in state_change_event_types]
    current = [0] * len(ranges)
    out: List[int] = []
    for tok in tokens:
        tok = int(tok)
        redundant = False
        for i, (lo, hi) in enumerate(ranges):
            if lo <= tok <= hi:
                redundant = redundant or current[i] == tok
                current[i] = tok
        if not redundant:
            out.append(tok)
    return out


def run_length_encode_shifts(tokens: Sequence[int], codec: event_codec.Codec) -> List[int]:
    """Single-step shifts -> shift tokens that carry the time since the START of the segment (so every run restates the
    absolute step, in chunks of at most max_shift_steps), written only in front of a non-shift event; trailing shifts
    are dropped."""
    out: List[int] = []
    pending, total = 0, 0
    for tok in tokens:
        tok = int(tok)
        if codec.is_shift_event_index(tok):
            pending += 1
            total += 1
            continue
        if pending > 0:
            left = total
            while left > 0:
                n = min(codec.max_shift_steps, left)
                out.append(n)                      # the token of Event('shift', n) is n itself
                left -= n
            pending = 0
        out.append(tok)
    return out


def merge_run_length_encoded_targets(targets: np.ndarray, codec: event_codec.Codec) -> List[int]:
    """Merge several run-length-encoded tracks ([tracks, events], zero-padded) into one stream ordered by step; a shift
    that restates the current step is not repeated."""
    targets = np.asarray(targets)
    n_tracks, length = targets.shape
    offsets = [0] * n_tracks
    step, out = 0, []
    while True:
        best_step, best = codec.max_shift_steps + 1, -1
        for t in range(n_tracks):
            if offsets[t] == length or targets[t, offsets[t]] == 0:      # exhausted (0 is padding)
                continue
            tok = int(targets[t, offsets[t]])
            if not codec.is_shift_event_index(tok):                      # events before the track's first shift: step 0
                best_step, best = 0, t
            elif tok < best_step:
                best_step, best = tok, t
        if best < 0:
            return out
        a = offsets[best] + (1 if best_step == step and best_step > 0 else 0)
        b = a + 1
        while b < length and not codec.is_shift_event_index(int(targets[best, b])):
            b += 1
        out.extend(int(x) for x in targets[best, a:b])
        step, offsets[best] = best_step, b


# ---------------------------------------------------------------------------------------------
# NoteSequence -> per-segment target tokens, the way the transcription tasks chain the pieces (tasks.py:125-183)
# ---------------------------------------------------------------------------------------------
def note_sequence_to_segment_targets(ns: NoteSequence, codec: event_codec.Codec, frame_times: Sequence[float],
                                     segment_frames: int, onsets_only: bool = False, include_ties: bool = True,
                                     trim_overlaps: bool = False) -> List[List[int]]:
    """One list of event tokens (codec indices, before the vocabulary's +3 offset and EOS) per consecutive segment of
    `segment_frames` audio frames, chained as the transcription tasks do (preprocessors.py:92-180 tokenize, then
    tasks.py:160-183): notes validated, events encoded and indexed to the frames, each segment's slice prefixed with its
    tie section, shifts run-length encoded, redundant velocity / program changes removed.  `trim_overlaps` first cuts
    overlapping notes of one (pitch, program) -- what a decoded sequence can represent."""
    if onsets_only and include_ties:
        raise ValueError('Ties not supported when only modeling onsets.')
    validate_note_sequence(ns)
    if trim_overlaps:
        ns = trim_overlapping_notes(ns)
    if onsets_only:
        times, values = note_sequence_to_onsets(ns)
    else:
        times, values = note_sequence_to_onsets_and_offsets_and_programs(ns)
    events, start, end, state_events, state_idx = encode_and_index_events(
        NoteEncodingState() if include_ties else None, times, values, note_event_data_to_events, codec, frame_times,
        encoding_state_to_events_fn=note_encoding_state_to_events if include_ties else None)
    tie = codec.encode_event(Event('tie', 0)) if include_ties else None
    out = []
    for f0 in range(0, len(frame_times), segment_frames):
        f1 = min(len(frame_times), f0 + segment_frames)
        seg = extract_target_sequence_with_indices(events, start[f0:f1], end[f0:f1], state_events, state_idx[f0:f1], tie)
        seg = run_length_encode_shifts(seg, codec)
        out.append(remove_redundant_state_changes(seg, codec, ['velocity', 'program']))
    return out
