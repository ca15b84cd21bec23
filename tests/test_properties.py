"""Property tests (hypothesis) of the host-side data formats on either side of the hot path: the vocabulary's decode forms
agree, run-length encoding is inverted by the stitch's time bookkeeping, resampling is linear and keeps the length
convention.  Size-independent properties in the sense of the parity contract: they hold for any input, not for fixtures."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from mt3_b200 import audio_io, event_codec as ec, note_encoding as ne, vocabularies as V

CODEC = V.build_codec(V.VocabularyConfig(num_velocity_bins=1))
VOCAB = V.vocabulary_from_codec(CODEC)


@settings(max_examples=200, deadline=None, derandomize=True)
@given(st.lists(st.integers(min_value=-5, max_value=1700), min_size=0, max_size=40))
def test_vocabulary_decode_forms_agree(ids):
    """decode_tf (array form: EOS and everything after it -> -1, invalid -> -2, length kept) cut at the first -1 equals the
    list form decode(), which stops at the first EOS (vocabularies.py:193-217 vs :241-271) -- for ANY ids, negative or huge."""
    arr = np.asarray(ids, np.int64).reshape(1, -1)
    tf_form = VOCAB.decode_tf(arr)[0]
    assert tf_form.shape == (len(ids),)
    cut = list(tf_form[:int(np.argmax(tf_form == -1)) + 1]) if (tf_form == -1).any() else list(tf_form)
    assert cut == list(VOCAB.decode(ids))
    assert all(t == -1 for t in tf_form[len(cut):])                      # EOS is sticky
    for i, t in zip(ids, tf_form):
        if t >= 0:
            assert t == i - 3 and 0 <= t < CODEC.num_classes
    good = [i for i in ids if 0 <= i < CODEC.num_classes]
    assert list(VOCAB.decode(VOCAB.encode(good))) == good                 # encode / decode are inverse on regular ids


@settings(max_examples=200, deadline=None, derandomize=True)
@given(st.lists(st.tuples(st.integers(0, 350), st.integers(0, 127)), min_size=0, max_size=30))
def test_run_length_encoded_shifts_restate_absolute_steps(events):
    """Any sequence of (gap in steps, pitch): written as single-step shifts and run-length encoded, the shift tokens in
    front of an event sum to its ABSOLUTE step since the segment start (in chunks of at most max_shift_steps), shift runs
    of length zero vanish and trailing shifts are dropped -- which is exactly what NoteDecoder.feed undoes."""
    shift1 = CODEC.encode_event(ec.Event('shift', 1))
    toks, steps, now = [], [], 0
    for gap, pitch in events:
        toks += [shift1] * gap
        now += gap
        toks.append(CODEC.encode_event(ec.Event('pitch', pitch)))
        steps.append(now)
    toks += [shift1] * 7                                                  # trailing silence
    rle = ne.run_length_encode_shifts(toks, CODEC)
    got_steps, run, last_emitted, seen_events = [], 0, 0, 0
    for t in rle:
        if CODEC.is_shift_event_index(t):
            assert 1 <= t <= CODEC.max_shift_steps
            run += t
        else:
            if run:
                last_emitted = run                                        # a shift run restates the absolute step
            got_steps.append(last_emitted)
            run = 0
            seen_events += 1
    assert run == 0 and seen_events == len(events)                        # no trailing shifts
    assert got_steps == steps
    assert not any(CODEC.is_shift_event_index(t) and t == 0 for t in rle)


@settings(max_examples=25, deadline=None, derandomize=True)
@given(st.sampled_from([8000, 11025, 22050, 32000, 44100, 48000]), st.integers(1, 4000), st.integers(0, 2 ** 31 - 1),
       st.floats(-2, 2), st.floats(-2, 2))
def test_resample_is_linear_and_keeps_the_length_convention(rate, n, seed, a, b):
    rng = np.random.default_rng(seed)
    x, y = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    rx, ry = audio_io.resample(x, rate, 16000), audio_io.resample(y, rate, 16000)
    g = math.gcd(rate, 16000)
    assert rx.shape == (-(-n * (16000 // g) // (rate // g)),)             # ceil(n * up / down), librosa's convention
    mix = audio_io.resample((a * x + b * y).astype(np.float32), rate, 16000)
    scale = 1e-5 * (abs(a) + abs(b) + 1) * max(1.0, float(np.abs(x).max() + np.abs(y).max()))
    assert np.abs(mix - (a * rx + b * ry)).max() <= scale
    assert np.abs(audio_io.resample(np.zeros(n, np.float32), rate, 16000)).max() == 0.0
