"""Known-answer tests of the token -> NoteSequence stitch, restated from the reference's
note_sequences_test.py:290-501 and metrics_utils_test.py:28-238 (same tokens, same expected notes)."""
import numpy as np
import pytest

from mt3_b200 import event_codec as ec
from mt3_b200 import note_decoding as nd


def _codec(*types):
    ranges = {'pitch': ec.EventRange('pitch', 0, 127), 'velocity': ec.EventRange('velocity', 0, 127),
              'drum': ec.EventRange('drum', 0, 127), 'program': ec.EventRange('program', 0, 127),
              'tie': ec.EventRange('tie', 0, 0)}
    return ec.Codec(max_shift_steps=100, steps_per_second=100, event_ranges=[ranges[t] for t in types])


FULL = _codec('pitch', 'velocity', 'drum', 'program', 'tie')     # note_sequences_test.py:23-35


def _decode(tokens, spec, start_time=0, max_time=None, codec=FULL):
    d = nd.NoteDecoder(codec, spec)
    inv, drop = d.feed(tokens, start_time, max_time)
    return d.flush(), inv, drop


def _notes(ns):
    return [(n.pitch, n.velocity, round(n.start_time, 6), round(n.end_time, 6), n.program, n.is_drum, n.instrument)
            for n in ns.notes]


def test_onsets_only():
    ns, inv, drop = _decode([25, 161, 50, 162], 'NoteOnsetEncodingSpec')          # :290-313
    assert (inv, drop) == (0, 0)
    assert _notes(ns) == [(60, 100, 0.25, 0.26, 0, False, 0), (61, 100, 0.50, 0.51, 0, False, 0)]
    assert ns.total_time == pytest.approx(0.51)
    ns, inv, drop = _decode([5, 161, 25, 162], 'NoteOnsetEncodingSpec')           # :315-338
    assert _notes(ns) == [(60, 100, 0.05, 0.06, 0, False, 0), (61, 100, 0.25, 0.26, 0, False, 0)]


def test_velocity_and_missing_offset():
    ns, inv, drop = _decode([5, 356, 161, 25, 229, 161], 'NoteEncodingSpec')       # :340-357
    assert (inv, drop) == (0, 0)
    assert _notes(ns) == [(60, 127, 0.05, 0.25, 0, False, 0)] and ns.total_time == pytest.approx(0.25)
    ns, inv, drop = _decode([5, 356, 161, 10, 161, 25, 229, 161], 'NoteEncodingSpec')   # :359-381
    assert _notes(ns) == [(60, 127, 0.05, 0.10, 0, False, 0), (60, 127, 0.10, 0.25, 0, False, 0)]


def test_multitrack():
    ns, inv, drop = _decode([5, 525, 356, 161, 15, 356, 394, 25, 525, 229, 161], 'NoteEncodingSpec')   # :383-409
    assert (inv, drop) == (0, 0)
    assert _notes(ns) == [(37, 127, 0.15, 0.16, 0, True, 9), (60, 127, 0.05, 0.25, 40, False, 0)]


def test_invalid_tokens_and_events():
    ns, inv, drop = _decode([5, -1, 161, -2, 25, 162, 9999], 'NoteOnsetEncodingSpec')   # :411-435
    assert (inv, drop) == (3, 0)
    assert _notes(ns) == [(60, 100, 0.05, 0.06, 0, False, 0), (61, 100, 0.25, 0.26, 0, False, 0)]
    ns, inv, drop = _decode([25, 230, 50, 161], 'NoteOnsetEncodingSpec')           # :483-501 (velocity event is invalid here)
    assert (inv, drop) == (1, 0)
    assert _notes(ns) == [(60, 100, 0.50, 0.51, 0, False, 0)]


def test_max_time():
    ns, inv, drop = _decode([161, 25, 162], 'NoteOnsetEncodingSpec', start_time=1.0, max_time=1.25)    # :437-461
    assert (inv, drop) == (0, 0)
    assert _notes(ns) == [(60, 100, 1.00, 1.01, 0, False, 0), (61, 100, 1.25, 1.26, 0, False, 0)]
    ns, inv, drop = _decode([5, 161, 30, 162], 'NoteOnsetEncodingSpec', start_time=1.0, max_time=1.25)  # :463-481
    assert (inv, drop) == (0, 2)
    assert _notes(ns) == [(60, 100, 1.05, 1.06, 0, False, 0)]


def _preds(tok_lists):
    return [{'raw_inputs': [i, i], 'start_time': 0.4 * i, 'est_tokens': t} for i, t in enumerate(tok_lists)]


def test_event_predictions_to_ns_onsets():
    res = nd.event_predictions_to_ns(_preds([[20, 160], [20, 161, 50, 162], [163, 20, 164]]), _codec('pitch'),
                                     'NoteOnsetEncodingSpec')                      # metrics_utils_test.py:28-80
    assert _notes(res['est_ns']) == [(59, 100, 0.20, 0.21, 0, False, 0), (60, 100, 0.60, 0.61, 0, False, 0),
                                     (62, 100, 0.80, 0.81, 0, False, 0), (63, 100, 1.00, 1.01, 0, False, 0)]
    assert res['est_ns'].total_time == pytest.approx(1.01)
    assert (res['est_invalid_events'], res['est_dropped_events']) == (0, 2)
    np.testing.assert_array_equal(res['raw_inputs'], [0, 0, 1, 1, 2, 2])
    assert res['start_times'] == [0.0, 0.4, 0.8]


def test_event_predictions_to_ns_with_offsets():
    res = nd.event_predictions_to_ns(_preds([[20, 356, 160], [20, 292, 161], [20, 229, 160, 161]]),
                                     _codec('pitch', 'velocity'), 'NoteEncodingSpec')   # :82-128
    assert _notes(res['est_ns']) == [(59, 127, 0.20, 1.00, 0, False, 0), (60, 63, 0.60, 1.00, 0, False, 0)]
    assert (res['est_invalid_events'], res['est_dropped_events']) == (0, 0)


def test_event_predictions_to_ns_multitrack_and_ties():
    c4 = _codec('pitch', 'velocity', 'drum', 'program')
    res = nd.event_predictions_to_ns(_preds([[20, 517, 356, 160], [20, 356, 399], [20, 517, 229, 160]]), c4,
                                     'NoteEncodingSpec')                           # :130-180
    assert _notes(res['est_ns']) == [(42, 127, 0.60, 0.61, 0, True, 9), (59, 127, 0.20, 1.00, 32, False, 0)]
    res = nd.event_predictions_to_ns(_preds([[613, 20, 517, 356, 160], [517, 160, 613, 20, 356, 399], [613]]), FULL,
                                     'NoteEncodingWithTiesSpec')                   # :182-238
    assert _notes(res['est_ns']) == [(42, 127, 0.60, 0.61, 0, True, 9), (59, 127, 0.20, 0.80, 32, False, 0)]
    assert res['est_ns'].total_time == pytest.approx(0.80)
    assert (res['est_invalid_events'], res['est_dropped_events']) == (0, 0)
    # predictions arriving out of order are sorted by start time (metrics_utils.py:88)
    p = _preds([[613, 20, 517, 356, 160], [517, 160, 613, 20, 356, 399], [613]])
    res2 = nd.event_predictions_to_ns(list(reversed(p)), FULL, 'NoteEncodingWithTiesSpec')
    assert _notes(res2['est_ns']) == _notes(res['est_ns'])


def test_assign_instruments_skips_drum_channel():
    ns = nd.NoteSequence()
    for prog in range(11):
        ns.add(0.0, 0.1, 60, 100, program=prog)
    ns.add(0.0, 0.1, 36, 100, is_drum=True)
    nd.assign_instruments(ns)
    assert [n.instrument for n in ns.notes] == [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 9]    # note_sequences.py:72-84


def test_midi_writer_roundtrip_header(tmp_path):
    ns, _, _ = _decode([5, 525, 356, 161, 15, 356, 394, 25, 525, 229, 161], 'NoteEncodingSpec')
    data = nd.note_sequence_to_midi_bytes(ns)
    assert data[:4] == b'MThd' and int.from_bytes(data[8:10], 'big') == 1
    assert int.from_bytes(data[10:12], 'big') == 3          # tempo track + melodic + drums
    assert int.from_bytes(data[12:14], 'big') == 220
    assert data.count(b'MTrk') == 3
    f = tmp_path / "t.mid"
    nd.note_sequence_to_midi_file(ns, str(f))
    assert f.read_bytes() == data


def _parse_smf(data: bytes):
    """A minimal, independent Standard MIDI File reader (SMF 1.0 spec): header, tracks, variable-length deltas, channel
    voice / meta events, no running status assumed away (it is handled).  Returns (format, division, tracks) with each track
    a list of (absolute tick, kind, channel, data1, data2)."""
    assert data[:4] == b'MThd' and int.from_bytes(data[4:8], 'big') == 6
    fmt, ntrk, div = (int.from_bytes(data[8 + 2 * i:10 + 2 * i], 'big') for i in range(3))
    pos, tracks = 14, []
    for _ in range(ntrk):
        assert data[pos:pos + 4] == b'MTrk'
        n = int.from_bytes(data[pos + 4:pos + 8], 'big')
        body, pos = data[pos + 8:pos + 8 + n], pos + 8 + n
        i, tick, status, ev = 0, 0, None, []
        while i < len(body):
            d = 0
            while True:                                   # variable-length quantity
                d = (d << 7) | (body[i] & 0x7F)
                i += 1
                if not body[i - 1] & 0x80:
                    break
            tick += d
            if body[i] & 0x80:
                status = body[i]
                i += 1
            if status == 0xFF:                            # meta: type, length, payload
                mtype, mlen = body[i], body[i + 1]
                ev.append((tick, 'meta', mtype, bytes(body[i + 2:i + 2 + mlen]), None))
                i += 2 + mlen
                continue
            kind, ch = status & 0xF0, status & 0x0F
            if kind in (0xC0, 0xD0):
                ev.append((tick, kind, ch, body[i], None))
                i += 1
            else:
                ev.append((tick, kind, ch, body[i], body[i + 1]))
                i += 2
        assert ev[-1][1] == 'meta' and ev[-1][2] == 0x2F  # end of track
        tracks.append(ev)
    assert pos == len(data)
    return fmt, div, tracks


def test_midi_writer_event_level_readback():
    """Every note of the NoteSequence comes back from the written file: pitch, velocity, onset / offset ticks
    (seconds x ticks_per_quarter x qpm / 60), program change per instrument track, drums on channel 9 (the notebook's
    note_seq.sequence_proto_to_midi_file target format)."""
    ns = nd.NoteSequence()
    ns.add(0.00, 0.50, 60, 100, program=0)
    ns.add(0.25, 0.75, 64, 90, program=0)
    ns.add(0.50, 0.50, 67, 80, program=0)              # zero-length note still gets one tick
    ns.add(0.10, 1.30, 40, 127, program=32)
    ns.add(1.00, 1.01, 36, 127, is_drum=True)
    ns.add(1.50, 1.51, 42, 127, is_drum=True)
    nd.assign_instruments(ns)
    qpm = 120.0
    fmt, div, tracks = _parse_smf(nd.note_sequence_to_midi_bytes(ns, qpm=qpm))
    assert fmt == 1 and div == ns.ticks_per_quarter == 220 and len(tracks) == 4     # tempo + 2 melodic + drums
    assert tracks[0][0][:4] == (0, 'meta', 0x51, (500000).to_bytes(3, 'big'))        # 120 qpm
    tps = div * qpm / 60.0
    got = []
    for tr in tracks[1:]:
        program, open_notes = None, {}
        for tick, kind, ch, d1, d2 in tr:
            if kind == 0xC0:
                program = d1
            elif kind == 0x90 and d2 > 0:
                open_notes.setdefault((ch, d1), []).append((tick, d2))
            elif kind == 0x80 or (kind == 0x90 and d2 == 0):
                on, vel = open_notes[(ch, d1)].pop(0)
                got.append((d1, vel, on, tick, program, ch))
        assert not any(open_notes.values())
    want = []
    for n in ns.notes:
        on = int(round(n.start_time * tps))
        off = max(on + 1, int(round(n.end_time * tps)))
        ch = 9 if n.is_drum else (n.instrument % 16 if n.instrument % 16 != 9 else 10)
        want.append((n.pitch, n.velocity, on, off, None if n.is_drum else n.program, ch))
    assert sorted(got) == sorted(want)
    assert {g[5] for g in got if g[4] is None} == {9}                                # drums: channel 9, no program change


def test_unknown_spec():
    with pytest.raises(ValueError):
        nd.NoteDecoder(FULL, 'NoSuchSpec')


def test_write_inferences_to_file(tmp_path):
    """Offline-eval surface (reference inference.py:34-138): raw model ids per segment -> JSON lines of notes per example."""
    import json
    from mt3_b200 import vocabularies
    vc = vocabularies.VocabularyConfig(num_velocity_bins=1)
    codec = vocabularies.build_codec(vc)
    vocab = vocabularies.vocabulary_from_codec(codec)

    def ids(*events):           # event -> raw model id (3 special ids first, vocabularies.py:241-271)
        return [codec.encode_event(ec.Event(t, v)) + 3 for t, v in events]

    seg0 = ids(('tie', 0), ('shift', 10), ('program', 0), ('velocity', 1), ('pitch', 60)) + [1, 0, 0]    # EOS = 1, then padding
    seg1 = ids(('program', 0), ('pitch', 60), ('tie', 0), ('shift', 50), ('program', 0), ('velocity', 0), ('pitch', 60)) + [1]
    other = ids(('tie', 0), ('shift', 20), ('program', 40), ('velocity', 1), ('pitch', 72),
                ('shift', 30), ('program', 40), ('velocity', 0), ('pitch', 72)) + [1]
    task_ds = [
        {'unique_id': ['b'], 'input_times': [0.0], 'raw_inputs': [], 'sequence': ['song-b']},
        {'unique_id': ['a'], 'input_times': [0.0], 'raw_inputs': [], 'sequence': ['song-a']},
        {'unique_id': ['a'], 'input_times': [2.048], 'raw_inputs': [], 'sequence': ['']},
    ]
    path = tmp_path / 'inferences.jsonl'
    nd.write_inferences_to_file(str(path), [other, seg0, seg1], task_ds, 'predict', vocabulary=vocab, vocab_config=vc,
                                onsets_only=False, use_ties=True)
    lines = [json.loads(l) for l in path.read_text().splitlines()]
    assert [l['id'] for l in lines] == ['song-a', 'song-b']
    a, b = lines[0]['est_notes'], lines[1]['est_notes']
    # example a: note-on at 0.10 s in segment 0, tied into segment 1 (starts at 2.04 s), released 0.50 s into it
    assert len(a) == 1 and a[0]['pitch'] == 60 and a[0]['start_time'] == pytest.approx(0.10) and a[0]['end_time'] == pytest.approx(2.54)
    assert len(b) == 1 and (b[0]['pitch'], b[0]['program']) == (72, 40)
    assert b[0]['start_time'] == pytest.approx(0.20) and b[0]['end_time'] == pytest.approx(0.30)
    with pytest.raises(ValueError):
        nd.write_inferences_to_file(str(path), [], [], 'score', vocabulary=vocab, vocab_config=vc)
    with pytest.raises(ValueError):
        nd.write_inferences_to_file(str(path), [], [], 'predict', vocabulary=vocab, vocab_config=vc, onsets_only=True, use_ties=True)


def test_frame_metrics_reference_kat_and_sklearn():
    ref, est = np.zeros((128, 5)), np.zeros((128, 5))                 # metrics_utils_test.py:240-256
    ref[10, 0] = ref[10, 1] = ref[10, 2] = 127                        # one overlapping frame, two false positives, two false negatives
    est[10, 2] = est[10, 3] = est[10, 4] = 127
    prec, rec, f1 = nd.frame_metrics(ref, est, velocity_threshold=1)
    assert prec == pytest.approx(1 / 3) and rec == pytest.approx(1 / 3) and f1 == pytest.approx(1 / 3)
    # rolls of different lengths, a quiet reference frame below the threshold; against the call the reference makes
    skm = pytest.importorskip("sklearn.metrics")
    rng = np.random.default_rng(0)
    ref = (rng.random((128, 40)) < 0.05) * rng.integers(1, 128, (128, 40))
    est = (rng.random((128, 33)) < 0.05) * rng.integers(1, 128, (128, 33))
    got = nd.frame_metrics(ref, est, velocity_threshold=30)
    est_p = np.pad(est, [(0, 0), (0, 7)])
    p, r, f, _ = skm.precision_recall_fscore_support((ref > 30).flatten(), (est_p > 0).flatten(), labels=[True, False])
    assert got == pytest.approx((p[0], r[0], f[0]))
    assert nd.frame_metrics(np.zeros((128, 3)), np.zeros((128, 0)), 0) == (0.0, 0.0, 0.0)


def test_note_sequence_to_pianoroll():
    ns = nd.NoteSequence()
    ns.add(0.0, 1.0, 60, 100)
    ns.add(0.5, 0.51, 60, 20)            # shorter than 50 ms: lengthened to 50 ms; overlaps the first note -> velocities add
    ns.add(2.0, 2.25, 36, 90, is_drum=True)
    roll = nd.note_sequence_to_pianoroll(ns, fps=100)
    assert roll.shape == (128, 225)
    assert roll[60, 0] == 100 and roll[60, 99] == 100 and roll[60, 100] == 0
    assert (roll[60, 50:55] == 120).all() and roll[60, 55] == 100
    assert (roll[36, 200:225] == 90).all()
    drums = nd.note_sequence_to_pianoroll(ns, fps=100, is_drum=True)          # fixed 50 ms for every note
    n = int((2.0 + 0.05) * 100)                                                # 204: int() of the float product, as pretty_midi does
    assert drums.shape == (128, n) and (drums[36, 200:n] == 90).all() and drums[60, 5] == 0 and drums[60, 4] == 100
    assert ns.notes[1].end_time == 0.51                                       # unlike the reference helper, the input is left alone
    assert nd.note_sequence_to_pianoroll(nd.NoteSequence(), 100).shape == (128, 0)
