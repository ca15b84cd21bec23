"""Notes -> event tokens (mt3_b200/note_encoding.py): the reference's known-answer tests for the encode direction,
restated with the same notes and expected tokens (note_sequences_test.py:42-289, run_length_encoding_test.py:45-104), and
the round trip through the stitch of mt3_b200.note_decoding (notes -> tokens per segment -> notes)."""
import numpy as np
import pytest

from mt3_b200 import event_codec as ec
from mt3_b200 import note_decoding as nd
from mt3_b200 import note_encoding as ne
from mt3_b200 import vocabularies

CODEC = ec.Codec(max_shift_steps=100, steps_per_second=100,
                 event_ranges=[ec.EventRange('pitch', 0, 127), ec.EventRange('velocity', 0, 127), ec.EventRange('drum', 0, 127),
                               ec.EventRange('program', 0, 127), ec.EventRange('tie', 0, 0)])     # note_sequences_test.py:25-37


def _ns(*notes):
    ns = nd.NoteSequence()
    for n in notes:
        ns.add(**n)
    return ns


def test_encode_and_index_onsets():
    ns = _ns(dict(start_time=1.0, end_time=1.1, pitch=61, velocity=100), dict(start_time=2.0, end_time=2.1, pitch=62, velocity=100),
             dict(start_time=3.0, end_time=3.1, pitch=63, velocity=100))
    frame_times = np.arange(0, 4, step=.001)
    times, values = ne.note_sequence_to_onsets(ns)
    events, start, end, _, _ = ne.encode_and_index_events(None, times, values, ne.note_event_data_to_events, CODEC, frame_times)
    assert len(start) == len(end) == len(frame_times) and len(events) == 403                 # :42-99
    np.testing.assert_array_equal(events, [1] * 100 + [162] + [1] * 100 + [163] + [1] * 100 + [164] + [1] * 100)
    assert (start[0], end[0]) == (0, 0)
    assert frame_times[1000] == 1.0 and (start[1000], end[1000]) == (100, 100)
    assert frame_times[2000] == 2.0 and (start[2000], end[2000]) == (201, 201)
    assert frame_times[3000] == 3.0 and (start[3000], end[3000]) == (302, 302)
    assert frame_times[-1] == 3.999 and (start[-1], end[-1]) == (402, 403)
    np.testing.assert_array_equal(end[:-1], start[1:])


def test_encode_and_index_velocity():
    ns = _ns(dict(start_time=1.0, end_time=3.0, pitch=61, velocity=1), dict(start_time=2.0, end_time=4.0, pitch=62, velocity=127))
    frame_times = np.arange(0, 4, step=.001)
    times, values = ne.note_sequence_to_onsets_and_offsets(ns)
    events, start, end, _, _ = ne.encode_and_index_events(None, times, values, ne.note_event_data_to_events, CODEC, frame_times)
    assert len(events) == 408                                                                 # :101-160
    np.testing.assert_array_equal(events, [1] * 100 + [230, 162] + [1] * 100 + [356, 163] + [1] * 100 + [229, 162] + [1] * 100 + [229, 163])
    assert (start[0], end[0]) == (0, 0) and (start[1000], end[1000]) == (100, 100) and (start[2000], end[2000]) == (202, 202)
    assert (start[3000], end[3000]) == (304, 304) and (start[-1], end[-1]) == (405, 408)


def test_encode_and_index_multitrack_with_state_events():
    ns = _ns(dict(start_time=0.0, end_time=1.0, pitch=37, velocity=127, is_drum=True),
             dict(start_time=1.0, end_time=3.0, pitch=61, velocity=127, program=0),
             dict(start_time=2.0, end_time=4.0, pitch=62, velocity=127, program=40))
    frame_times = np.arange(0, 4, step=.001)
    times, values = ne.note_sequence_to_onsets_and_offsets_and_programs(ns)
    tokens, start, end, state_tokens, state_idx = ne.encode_and_index_events(
        ne.NoteEncodingState(), times, values, ne.note_event_data_to_events, CODEC, frame_times,
        encoding_state_to_events_fn=ne.note_encoding_state_to_events)
    E = ec.Event
    want = ([E('velocity', 127), E('drum', 37)] + [E('shift', 1)] * 100 + [E('program', 0), E('velocity', 127), E('pitch', 61)] +
            [E('shift', 1)] * 100 + [E('program', 40), E('velocity', 127), E('pitch', 62)] + [E('shift', 1)] * 100 +
            [E('program', 0), E('velocity', 0), E('pitch', 61)] + [E('shift', 1)] * 100 + [E('program', 40), E('velocity', 0), E('pitch', 62)])
    assert len(tokens) == 414                                                                 # :162-257
    np.testing.assert_array_equal(tokens, [CODEC.encode_event(e) for e in want])
    want_state = [E('tie', 0), E('tie', 0), E('program', 0), E('pitch', 61), E('tie', 0), E('program', 0), E('pitch', 61),
                  E('program', 40), E('pitch', 62), E('tie', 0), E('program', 40), E('pitch', 62), E('tie', 0)]
    np.testing.assert_array_equal(state_tokens, [CODEC.encode_event(e) for e in want_state])
    assert len(start) == len(end) == len(state_idx) == len(frame_times)
    assert (start[0], end[0], state_idx[0]) == (0, 0, 0)
    assert (start[1000], end[1000], state_idx[1000]) == (102, 102, 1)
    assert (start[2000], end[2000], state_idx[2000]) == (205, 205, 2)
    assert (start[3000], end[3000], state_idx[3000]) == (308, 308, 5)
    assert (start[-1], end[-1], state_idx[-1]) == (410, len(want), 10)


def test_encode_and_index_last_token_alignment():
    ns = _ns(dict(start_time=0.0, end_time=0.1, pitch=60, velocity=100))
    frame_times = np.arange(0, 1.008, step=.008)
    times, values = ne.note_sequence_to_onsets(ns)
    events, start, end, _, _ = ne.encode_and_index_events(None, times, values, ne.note_event_data_to_events, CODEC, frame_times)
    assert len(start) == len(end) == len(frame_times) and len(events) == 102                 # :259-288
    np.testing.assert_array_equal(events, [161] + [1] * 101)
    assert (start[0], end[0]) == (0, 0) and (start[125], end[125]) == (101, 102)


def _reference_indexing_loop(steps_tokens, frame_times, sps):
    """The incremental fill of run_length_encoding.py:118-161 written out plainly (no state events), as an independent check
    of the searchsorted form on irregular inputs.  steps_tokens: [(step, n_tokens)] in order."""
    n_events, cur, cur_idx, start = 0, 0, 0, []

    def fill():
        while len(start) < len(frame_times) and frame_times[len(start)] < cur / sps:
            start.append(cur_idx)

    for step, n in steps_tokens:
        while step > cur:
            n_events += 1
            cur += 1
            fill()
            cur_idx = n_events
        n_events += n
    while cur / sps <= frame_times[-1]:
        n_events += 1
        cur += 1
        fill()
        cur_idx = n_events
    return start, n_events


@pytest.mark.parametrize("seed", range(6))
def test_frame_indices_equal_the_incremental_fill(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(0, 40))
    times = np.sort(rng.uniform(0, 5.0, n)).round(int(rng.integers(2, 5)))
    ns = nd.NoteSequence()
    for t in times:
        ns.add(float(t), float(t) + 0.01, int(rng.integers(20, 100)), 100)
    hop = [0.008, 0.001, 0.0173, 0.01][seed % 4]
    frame_times = np.arange(0, float(rng.uniform(0.5, 6.0)), hop)
    et, ev = ne.note_sequence_to_onsets(ns)
    events, start, end, _, _ = ne.encode_and_index_events(None, et, ev, ne.note_event_data_to_events, CODEC, frame_times)
    order = np.argsort(np.asarray(et), kind='stable')
    steps = [(round(float(et[i]) * 100), 1) for i in order]
    want_start, want_len = _reference_indexing_loop(steps, frame_times, 100)
    assert len(events) == want_len
    np.testing.assert_array_equal(start, want_start)
    np.testing.assert_array_equal(end, want_start[1:] + [want_len])


def test_remove_redundant_state_changes():
    got = ne.remove_redundant_state_changes([3, 525, 356, 161, 2, 525, 356, 161, 355, 394], CODEC, ['velocity', 'program'])
    assert got == [3, 525, 356, 161, 2, 161, 355, 394]                                        # run_length_encoding_test.py:45-56


def test_run_length_encode_shifts():
    assert ne.run_length_encode_shifts([1, 1, 1, 161, 1, 1, 1, 162, 1, 1, 1], CODEC) == [3, 161, 6, 162]      # :58-67
    assert ne.run_length_encode_shifts([1] * 202 + [161, 1, 1, 1], CODEC) == [100, 100, 2, 161]             # :69-78
    assert ne.run_length_encode_shifts([1, 1, 1, 161, 162, 1, 1, 1], CODEC) == [3, 161, 162]                # :80-89
    assert ne.run_length_encode_shifts([], CODEC) == [] and ne.run_length_encode_shifts([1, 1], CODEC) == []


def test_merge_run_length_encoded_targets():
    targets = np.array([[3, 161, 162, 5, 163], [160, 164, 3, 165, 0]])
    assert ne.merge_run_length_encoded_targets(targets, CODEC) == [160, 164, 3, 161, 162, 165, 5, 163]       # :91-104


def test_note_sequence_utilities():
    ns = _ns(dict(start_time=0.0, end_time=2.0, pitch=60, velocity=80), dict(start_time=1.0, end_time=3.0, pitch=60, velocity=90),
             dict(start_time=1.0, end_time=1.0, pitch=61, velocity=90), dict(start_time=0.5, end_time=0.7, pitch=60, velocity=70, program=5),
             dict(start_time=0.0, end_time=0.1, pitch=36, velocity=100, is_drum=True))
    trimmed = ne.trim_overlapping_notes(ns)
    assert [(n.pitch, n.program, n.start_time, n.end_time) for n in trimmed.notes] == [
        (60, 0, 0.0, 1.0), (60, 0, 1.0, 3.0), (60, 5, 0.5, 0.7), (36, 0, 0.0, 0.1)]            # overlap cut, zero-length note dropped
    assert ns.notes[0].end_time == 2.0                                                        # the input is not modified
    with pytest.raises(ValueError, match='start time >= end time'):
        ne.validate_note_sequence(ns)
    with pytest.raises(ValueError, match='zero velocity'):
        ne.validate_note_sequence(_ns(dict(start_time=0.0, end_time=1.0, pitch=60, velocity=0)))
    tr = ne.extract_track(ns, 0, True)
    assert [(n.pitch, n.is_drum) for n in tr.notes] == [(36, True)] and tr.total_time == 0.1
    assert ne.extract_track(ns, 99, False).total_time == 0.0
    a = ne.note_arrays_to_note_sequence([0.0, 1.0], [60, 62])
    assert [(n.pitch, n.velocity, n.start_time, n.end_time, n.program) for n in a.notes] == [(60, 100, 0.0, 0.01, 0), (62, 100, 1.0, 1.01, 0)]
    b = ne.note_arrays_to_note_sequence([0.0, 1.0], [60, 38], offset_times=[0.5, 1.2], velocities=[10, 20], programs=[3, 0], is_drums=[False, True])
    assert [(n.velocity, n.end_time, n.program, n.is_drum, n.instrument) for n in b.notes] == [(10, 0.5, 3, False, 0), (20, 1.2, 0, True, 9)]


def _random_ns(rng, seconds, n_notes, programs=(0, 24, 40), drums=True):
    ns = nd.NoteSequence()
    for _ in range(n_notes):
        onset = round(float(rng.uniform(0, seconds - 0.05)), 2)
        dur = round(float(rng.uniform(0.02, 1.5)), 2)
        if drums and rng.random() < 0.2:
            ns.add(onset, onset + 0.01, int(rng.integers(35, 82)), int(rng.integers(1, 128)), is_drum=True)
        else:
            ns.add(onset, min(seconds - 0.01, onset + dur), int(rng.integers(40, 90)), int(rng.integers(1, 128)),
                   program=int(rng.choice(programs)))
    return ne.trim_overlapping_notes(ns)


def _key(ns):
    return sorted((n.is_drum, n.program, n.pitch, round(n.start_time, 6), round(n.end_time, 6), n.velocity) for n in ns.notes)


@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("num_velocity_bins", [127, 1])
def test_round_trip_notes_tokens_notes_across_segments(seed, num_velocity_bins):
    """notes -> per-segment target tokens with tie sections (the mt3 task's chain) -> the stitch of note_decoding gives the
    same notes back: onsets / offsets on the 10 ms grid, programs, drums, velocities up to their quantisation -- the stitch
    and the encoder are written independently of each other, from the two halves of the reference."""
    rng = np.random.default_rng(100 + seed)
    vc = vocabularies.VocabularyConfig(num_velocity_bins=num_velocity_bins)
    codec = vocabularies.build_codec(vc)
    seconds = 9.0
    ns = _random_ns(rng, seconds, 60)
    frame_times = np.arange(int(seconds * 125)) / 125.0                # hop 128 at 16 kHz
    seg_frames = 256
    segs = ne.note_sequence_to_segment_targets(ns, codec, frame_times, seg_frames, include_ties=True)
    assert len(segs) == -(-len(frame_times) // seg_frames)
    preds = []
    for i, toks in enumerate(segs):
        start = float(frame_times[i * seg_frames])
        start -= start % (1 / codec.steps_per_second)
        preds.append({'est_tokens': toks, 'start_time': start, 'raw_inputs': []})
    res = nd.event_predictions_to_ns(preds, codec, 'NoteEncodingWithTiesSpec')
    assert res['est_invalid_events'] == 0 and res['est_dropped_events'] == 0
    want = nd.NoteSequence()
    for n in ns.notes:       # what the tokens can carry: velocity through its bin; drums as 10 ms hits
        vel = vocabularies.bin_to_velocity(vocabularies.velocity_to_bin(n.velocity, num_velocity_bins), num_velocity_bins)
        want.add(n.start_time, n.start_time + 0.01 if n.is_drum else n.end_time, n.pitch, vel, program=0 if n.is_drum else n.program,
                 is_drum=n.is_drum)
    assert _key(res['est_ns']) == _key(want)


def test_segment_targets_onsets_only_and_errors():
    ns = _ns(dict(start_time=0.5, end_time=0.6, pitch=60, velocity=100), dict(start_time=2.5, end_time=2.6, pitch=62, velocity=100))
    codec = vocabularies.build_codec(vocabularies.VocabularyConfig(num_velocity_bins=1))
    frame_times = np.arange(500) / 125.0
    segs = ne.note_sequence_to_segment_targets(ns, codec, frame_times, 256, onsets_only=True, include_ties=False)
    p60, p62 = (codec.encode_event(ec.Event('pitch', p)) for p in (60, 62))
    assert segs == [[50, p60], [46, p62]]          # shifts count from the segment start: frame 256 = 2.048 s lies in step 204, onset at step 250
    with pytest.raises(ValueError):
        ne.note_sequence_to_segment_targets(ns, codec, frame_times, 256, onsets_only=True, include_ties=True)
    with pytest.raises(IndexError):
        ne.encode_and_index_events(None, [], [], ne.note_event_data_to_events, codec, [])


def test_batch_infer_dataset_and_jsonl_from_encoded_targets(tmp_path):
    """The T5X-infer style flow on the CPU, with the label encoder standing in for the model: two recordings -> per-segment
    examples (InferenceModel.build_infer_dataset) -> 'inferences' = the reference transcription's own target tokens as raw
    model ids (+3, EOS, zero padding) -> write_inferences_to_file -> the JSON lines carry the notes of each recording."""
    import json
    from mt3_b200 import inference, spectrograms
    im = inference.InferenceModel.__new__(inference.InferenceModel)          # host-side pieces only: no GPU, no weights
    im.spectrogram_config = spectrograms.SpectrogramConfig()
    im.inputs_length, im.outputs_length = 256, 1024
    vc = vocabularies.VocabularyConfig(num_velocity_bins=1)
    im.codec = vocabularies.build_codec(vc)
    im.vocabulary = vocabularies.vocabulary_from_codec(im.codec)
    rng = np.random.default_rng(7)
    recs, truth = [], {}
    for name, seconds in (('song-b', 5.0), ('song-a', 3.0)):
        recs.append({'id': name, 'audio': np.zeros(int(seconds * 16000), np.float32)})
        truth[name] = _random_ns(rng, seconds, 25)
    task_ds, segs, n_valid = im.build_infer_dataset(recs)
    assert segs.shape == (3 + 2, 256 * 128) and [d['sequence'][0] for d in task_ds] == ['song-b', '', '', 'song-a', '']
    assert list(n_valid) == [256, 256, 114, 256, 120]                        # 80 000 + 128 and 48 000 + 128 padded samples
    inferences, i = [], 0
    for rec in recs:
        n_frames = sum(len(d['input_times']) for d in task_ds if d['unique_id'][0] == rec['id'])
        frame_times = np.arange(n_frames) / 125.0
        for toks in ne.note_sequence_to_segment_targets(truth[rec['id']], im.codec, frame_times, 256):
            row = np.zeros(1024, np.int32)
            ids = list(im.vocabulary.encode(toks)) + [1]
            row[:len(ids)] = ids
            inferences.append(row)
            i += 1
    assert i == len(task_ds)
    path = tmp_path / 'inferences.jsonl'
    nd.write_inferences_to_file(str(path), inferences, task_ds, 'predict', vocabulary=im.vocabulary, vocab_config=vc,
                                onsets_only=False, use_ties=True)
    lines = [json.loads(l) for l in path.read_text().splitlines()]
    assert [l['id'] for l in lines] == ['song-a', 'song-b']
    for l in lines:
        got = sorted((n['is_drum'], n['program'], n['pitch'], round(n['start_time'], 6), round(n['end_time'], 6)) for n in l['est_notes'])
        want = sorted((n.is_drum, 0 if n.is_drum else n.program, n.pitch, round(n.start_time, 6),
                       round(n.start_time + 0.01 if n.is_drum else n.end_time, 6)) for n in truth[l['id']].notes)
        assert got == want
