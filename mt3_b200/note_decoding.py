"""Event tokens -> notes: the "stitch" that completes InferenceModel.__call__.

Behaviour follows the reference's decode path
    metrics_utils.event_predictions_to_ns      (metrics_utils.py:119-146)
      -> decode_and_combine_predictions        (metrics_utils.py:59-116)
      -> run_length_encoding.decode_events     (run_length_encoding.py:371-423)
      -> note_sequences.decode_note_event / decode_note_onset_event / flush (note_sequences.py:262-408)
but is organised as one small state machine (`NoteDecoder`) over a plain-Python `NoteSequence`
(note_seq's protobuf is not installable here).  The three encoding specs of the reference
(note_sequences.py:415-446) are selected by name: 'NoteOnsetEncodingSpec', 'NoteEncodingSpec',
'NoteEncodingWithTiesSpec'.  Host-side integer/float bookkeeping; nothing here touches the GPU.
"""
from __future__ import annotations

import dataclasses
import struct
from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from . import event_codec, vocabularies

DEFAULT_VELOCITY = 100          # note_sequences.py:28
DEFAULT_NOTE_DURATION = 0.01    # note_sequences.py:29
MIN_NOTE_DURATION = 0.01        # note_sequences.py:32


@dataclasses.dataclass
class Note:
    pitch: int
    velocity: int
    start_time: float
    end_time: float
    program: int = 0
    is_drum: bool = False
    instrument: int = 0


@dataclasses.dataclass
class NoteSequence:
    """Stand-in for note_seq.NoteSequence: the fields the decode path fills."""
    ticks_per_quarter: int = 220
    notes: List[Note] = dataclasses.field(default_factory=list)
    total_time: float = 0.0

    def add(self, start_time, end_time, pitch, velocity, program=0, is_drum=False) -> Note:
        n = Note(pitch=int(pitch), velocity=int(velocity), start_time=float(start_time), end_time=float(end_time),
                 program=int(program), is_drum=bool(is_drum))
        self.notes.append(n)
        self.total_time = max(self.total_time, n.end_time)
        return n


def assign_instruments(ns: NoteSequence) -> None:
    """One instrument per program in order of first appearance, skipping 9 which is reserved for
    drums (note_sequences.py:72-84)."""
    by_program: Dict[int, int] = {}
    for note in ns.notes:
        if note.is_drum:
            note.instrument = 9
        else:
            if note.program not in by_program:
                k = len(by_program)
                by_program[note.program] = k if k < 9 else k + 1
            note.instrument = by_program[note.program]


class NoteDecoder:
    """Streaming decoder over consecutive segments of event tokens.

    mode 'onsets'  : every pitch event is a 10 ms note (NoteOnsetEncodingSpec)
    mode 'notes'   : velocity / program / pitch / drum events with note-offs (NoteEncodingSpec)
    mode 'ties'    : as 'notes', plus the tie section at the start of each segment that declares
                     which active notes continue (NoteEncodingWithTiesSpec)
    """

    MODES = {'NoteOnsetEncodingSpec': 'onsets', 'NoteEncodingSpec': 'notes', 'NoteEncodingWithTiesSpec': 'ties'}

    def __init__(self, codec: event_codec.Codec, encoding_spec: str):
        if encoding_spec not in self.MODES:
            raise ValueError('unknown encoding spec: %s' % (encoding_spec,))
        self.codec = codec
        self.mode = self.MODES[encoding_spec]
        self.ns = NoteSequence()
        self.current_time = 0.0
        self.velocity = DEFAULT_VELOCITY          # applies to following pitch events; 0 = note-off
        self.program = 0
        self.active: Dict[Tuple[int, int], Tuple[float, int]] = {}    # (pitch, program) -> (onset, velocity)
        self.tied: set = set()
        self.in_tie_section = False
        self.invalid_events = 0
        self.dropped_events = 0

    # -- per-segment -----------------------------------------------------------------------
    def begin_segment(self) -> None:
        if self.mode == 'ties':                       # begin_tied_pitches_section
            self.tied = set()
            self.in_tie_section = True

    def feed(self, tokens: Sequence[int], start_time: float, max_time: Optional[float]) -> Tuple[int, int]:
        """Decode one segment's tokens (run_length_encoding.py:371-423): shifts are absolute within
        the segment and reset by any non-shift event; events at a time beyond `max_time` (the next
        segment's start) are dropped together with everything after them."""
        invalid = dropped = 0
        steps = 0
        now = start_time
        n = len(tokens)
        for i, tok in enumerate(tokens):
            try:
                ev = self.codec.decode_event_index(int(tok))
            except ValueError:
                invalid += 1
                continue
            if ev.type == 'shift':
                steps += ev.value
                now = start_time + steps / self.codec.steps_per_second
                if max_time and now > max_time:
                    dropped = n - i
                    break
                continue
            steps = 0
            try:
                self._event(now, ev)
            except ValueError:
                invalid += 1
        self.invalid_events += invalid
        self.dropped_events += dropped
        return invalid, dropped

    # -- event handlers -------------------------------------------------------------------
    def _close(self, key: Tuple[int, int], end_time: float) -> None:
        onset, vel = self.active.pop(key)
        self.ns.add(onset, max(end_time, onset + MIN_NOTE_DURATION), key[0], vel, program=key[1])

    def _event(self, time: float, ev: event_codec.Event) -> None:
        if self.mode == 'onsets':
            if ev.type != 'pitch':
                raise ValueError('unexpected event type: %s' % ev.type)
            self.ns.add(time, time + DEFAULT_NOTE_DURATION, ev.value, DEFAULT_VELOCITY)
            return
        if time < self.current_time:
            raise ValueError('event time < current time, %f < %f' % (time, self.current_time))
        self.current_time = time
        if ev.type == 'pitch':
            key = (ev.value, self.program)
            if self.in_tie_section:
                if key not in self.active:
                    raise ValueError('inactive pitch/program in tie section: %d/%d' % key)
                if key in self.tied:
                    raise ValueError('pitch/program is already tied: %d/%d' % key)
                self.tied.add(key)
            elif self.velocity == 0:
                if key not in self.active:
                    raise ValueError('note-off for inactive pitch/program: %d/%d' % key)
                self._close(key, time)
            else:
                if key in self.active:                 # re-onset of a sounding note: end it first
                    self._close(key, time)
                self.active[key] = (time, self.velocity)
        elif ev.type == 'drum':
            if self.velocity == 0:
                raise ValueError('velocity cannot be zero for drum event')
            self.ns.add(time, max(time + DEFAULT_NOTE_DURATION, time + MIN_NOTE_DURATION), ev.value, self.velocity,
                        is_drum=True)
        elif ev.type == 'velocity':
            bins = vocabularies.num_velocity_bins_from_codec(self.codec)
            self.velocity = vocabularies.bin_to_velocity(ev.value, bins)
        elif ev.type == 'program':
            self.program = ev.value
        elif ev.type == 'tie':
            if not self.in_tie_section:
                raise ValueError('tie section end event when not in tie section')
            for key in [k for k in self.active if k not in self.tied]:
                self._close(key, self.current_time)
            self.in_tie_section = False
        else:
            raise ValueError('unexpected event type: %s' % ev.type)

    # -- end ------------------------------------------------------------------------------------
    def flush(self) -> NoteSequence:
        """End all still-sounding notes (note_sequences.py:396-408) and number the instruments."""
        if self.mode != 'onsets':
            for onset, _ in self.active.values():
                self.current_time = max(self.current_time, onset + MIN_NOTE_DURATION)
            for key in list(self.active.keys()):
                self._close(key, self.current_time)
            assign_instruments(self.ns)
        return self.ns


def event_predictions_to_ns(predictions: Sequence[Mapping[str, Any]], codec: event_codec.Codec,
                            encoding_spec: str) -> Mapping[str, Any]:
    """Convert a sequence of per-segment predictions ({'est_tokens', 'start_time', 'raw_inputs'})
    to one combined NoteSequence (metrics_utils.py:119-146).  A segment may not emit events past
    the start of the following segment (metrics_utils.py:101-111)."""
    order = sorted(range(len(predictions)), key=lambda i: predictions[i]['start_time'])
    dec = NoteDecoder(codec, encoding_spec)
    for rank, i in enumerate(order):
        pred = predictions[i]
        dec.begin_segment()
        max_time = predictions[order[rank + 1]]['start_time'] if rank + 1 < len(order) else None
        dec.feed(pred['est_tokens'], pred['start_time'], max_time)
    ns = dec.flush()
    raws = [np.asarray(predictions[i]['raw_inputs']) for i in order]
    raw_inputs = np.concatenate(raws, axis=0) if raws and all(r.ndim > 0 for r in raws) else np.zeros((0,))
    return {
        'raw_inputs': raw_inputs,
        'start_times': [predictions[i]['start_time'] for i in order],
        'est_ns': ns,
        'est_invalid_events': dec.invalid_events,
        'est_dropped_events': dec.dropped_events,
    }


# ---------------------------------------------------------------------------------------------
# Frame-level view of a NoteSequence (metrics_utils.py:149-196)
# ---------------------------------------------------------------------------------------------
def note_sequence_to_pianoroll(ns: NoteSequence, fps: float, is_drum: bool = False) -> np.ndarray:
    """[128, frames] piano roll of velocities (metrics_utils.get_prettymidi_pianoroll without pretty_midi, which is not
    installable here): every drum hit lasts 50 ms, every other note at least 50 ms; a note covers the frames
    int(start * fps) .. int(end * fps) - 1 and simultaneous notes of one pitch add their velocities (pretty_midi's
    get_piano_roll summed over instruments).  The input is not modified."""
    ends = [n.start_time + 0.05 if (is_drum or n.end_time - n.start_time < 0.05) else n.end_time for n in ns.notes]
    frames = int(max(ends, default=0.0) * fps)
    roll = np.zeros((128, frames), np.float64)
    for n, end in zip(ns.notes, ends):
        roll[n.pitch, int(n.start_time * fps):int(end * fps)] += n.velocity
    return roll


def frame_metrics(ref_pianoroll: np.ndarray, est_pianoroll: np.ndarray, velocity_threshold: int) -> Tuple[float, float, float]:
    """Frame precision, recall and F1 (metrics_utils.py:175-196): the shorter roll is zero-padded, reference frames count
    when louder than `velocity_threshold`, estimated frames when non-zero; an empty denominator gives 0 (sklearn's
    zero_division default, which the reference inherits)."""
    ref, est = np.asarray(ref_pianoroll), np.asarray(est_pianoroll)
    width = max(ref.shape[1], est.shape[1])
    ref = np.pad(ref, [(0, 0), (0, width - ref.shape[1])])
    est = np.pad(est, [(0, 0), (0, width - est.shape[1])])
    ref_on, est_on = ref > velocity_threshold, est > 0
    tp = int(np.count_nonzero(ref_on & est_on))
    n_est, n_ref = int(np.count_nonzero(est_on)), int(np.count_nonzero(ref_on))
    precision = tp / n_est if n_est else 0.0
    recall = tp / n_ref if n_ref else 0.0
    f1 = 2 * precision * recall / (precision + recall) if precision + recall else 0.0
    return precision, recall, f1


# ---------------------------------------------------------------------------------------------
# Standard MIDI file writer (the notebook's note_seq.sequence_proto_to_midi_file)
# ---------------------------------------------------------------------------------------------
def _vlq(n: int) -> bytes:
    out = [n & 0x7F]
    n >>= 7
    while n:
        out.append(0x80 | (n & 0x7F))
        n >>= 7
    return bytes(reversed(out))


def note_sequence_to_midi_bytes(ns: NoteSequence, qpm: float = 120.0) -> bytes:
    """Format-1 SMF: a tempo track plus one track per instrument; drums on channel 9."""
    tpq = ns.ticks_per_quarter
    ticks_per_second = tpq * qpm / 60.0
    tempo = int(round(60e6 / qpm))
    tracks = [b'\x00\xff\x51\x03' + struct.pack('>I', tempo)[1:] + b'\x00\xff\x2f\x00']
    by_inst: Dict[int, List[Note]] = {}
    for n in ns.notes:
        by_inst.setdefault(n.instrument, []).append(n)
    for inst in sorted(by_inst):
        notes = by_inst[inst]
        drum = notes[0].is_drum
        ch = 9 if drum else (inst % 16 if inst % 16 != 9 else 10)
        events = []
        if not drum:
            events.append((0, 0, bytes([0xC0 | ch, notes[0].program & 0x7F])))
        for n in notes:
            on = int(round(n.start_time * ticks_per_second))
            off = max(on + 1, int(round(n.end_time * ticks_per_second)))
            events.append((on, 2, bytes([0x90 | ch, n.pitch & 0x7F, max(1, min(127, n.velocity))])))
            events.append((off, 1, bytes([0x80 | ch, n.pitch & 0x7F, 0])))
        events.sort(key=lambda e: (e[0], e[1]))
        data, last = bytearray(), 0
        for t, _, msg in events:
            data += _vlq(t - last) + msg
            last = t
        data += b'\x00\xff\x2f\x00'
        tracks.append(bytes(data))
    out = bytearray(b'MThd' + struct.pack('>IHHH', 6, 1, len(tracks), tpq))
    for t in tracks:
        out += b'MTrk' + struct.pack('>I', len(t)) + t
    return bytes(out)


def note_sequence_to_midi_file(ns: NoteSequence, path: str, qpm: float = 120.0) -> None:
    with open(path, 'wb') as f:
        f.write(note_sequence_to_midi_bytes(ns, qpm))


# ---------------------------------------------------------------------------------------------
# Offline-eval surface: combine per-segment predictions by example id and write the
# T5X-`infer`-compatible JSON lines file (reference inference.py:34-138, metrics_utils.py:38-56).
# ---------------------------------------------------------------------------------------------
def combine_predictions_by_id(predictions: Sequence[Mapping[str, Any]], combine_predictions_fn) -> Mapping[str, Any]:
    """Concatenate predicted examples, grouping by 'unique_id' (metrics_utils.py:38-56)."""
    by_id: dict = {}
    for pred in predictions:
        by_id.setdefault(pred['unique_id'], []).append(pred)
    return {uid: combine_predictions_fn(preds) for uid, preds in by_id.items()}


def note_to_dict(note: Note) -> Mapping[str, Any]:
    return {'start_time': note.start_time, 'end_time': note.end_time, 'pitch': note.pitch, 'velocity': note.velocity,
            'program': note.program, 'is_drum': note.is_drum}


def write_inferences_to_file(path: str, inferences: Sequence[Any], task_ds: Sequence[Mapping[str, Any]], mode: str,
                             vocabulary=None, vocab_config=None, onsets_only: bool = False, use_ties: bool = True) -> None:
    """Writes model predictions as JSON lines {'id', 'est_notes': [...]}, one line per full example
    (reference inference.py:34-138).  `inferences`: RAW model ids per segment (the output of predict_batch,
    decoded here with `vocabulary.decode_tf` + trim at EOS); `task_ds`: per-segment dicts with 'unique_id' [1],
    'input_times', 'raw_inputs' and 'sequence' [1] (the example id stands in for the serialized reference
    NoteSequence, which needs note_seq: only the first segment of an example carries it, as in the reference)."""
    from . import vocabularies
    if mode == 'score':
        raise ValueError('`score` mode currently not supported in MT3')
    if not vocabulary:
        raise ValueError('`vocabulary` parameter required in `predict` mode')
    if onsets_only and use_ties:
        raise ValueError('ties not compatible with onset-only transcription')
    encoding_spec = 'NoteOnsetEncodingSpec' if onsets_only else ('NoteEncodingWithTiesSpec' if use_ties else 'NoteEncodingSpec')
    codec = vocabularies.build_codec(vocab_config)
    targets, predictions = [], []
    for inp, output in zip(task_ds, inferences):
        tokens = np.asarray(vocabulary.decode_tf(np.asarray(output, np.int32)))
        if vocabularies.DECODED_EOS_ID in tokens:
            tokens = tokens[:np.argmax(tokens == vocabularies.DECODED_EOS_ID)]
        start_time = inp['input_times'][0]
        start_time -= start_time % (1 / codec.steps_per_second)      # round down to the symbolic token step
        uid = inp['unique_id'][0]
        targets.append({'unique_id': uid, 'ref_id': inp['sequence'][0] if inp['sequence'][0] else None})
        predictions.append({'unique_id': uid, 'est_tokens': tokens, 'start_time': start_time, 'raw_inputs': inp['raw_inputs']})
    full_targets = {t['unique_id']: t['ref_id'] for t in targets if t['ref_id']}
    full_predictions = combine_predictions_by_id(
        predictions, lambda preds: event_predictions_to_ns(preds, codec=codec, encoding_spec=encoding_spec))
    assert sorted(full_targets.keys()) == sorted(full_predictions.keys())
    import json
    with open(path, 'w') as f:
        for uid in sorted(full_targets.keys()):
            f.write(json.dumps({'id': str(full_targets[uid]),
                                'est_notes': [note_to_dict(n) for n in full_predictions[uid]['est_ns'].notes]}) + '\n')
