"""CPU tests of the audio input side (mt3_b200/audio_io.py): the notebook's
note_seq.audio_io.wav_data_to_samples_librosa(data, sample_rate=16000) contract -- decode, mono mix-down, resample --
checked against the standard library's `wave` writer, scipy's polyphase resampler run with the same filter, and
closed-form tones."""
import io
import math
import struct
import wave

import numpy as np
import pytest

from mt3_b200 import audio_io as A


def _riff(fmt_body: bytes, data: bytes, extra: bytes = b'') -> bytes:
    body = b'WAVE' + b'fmt ' + struct.pack('<I', len(fmt_body)) + fmt_body + extra + b'data' + struct.pack('<I', len(data)) + data
    if len(data) & 1:
        body += b'\x00'
    return b'RIFF' + struct.pack('<I', len(body)) + body


def _fmt(tag, channels, rate, bits, extensible_sub=None):
    base = struct.pack('<HHIIHH', tag if extensible_sub is None else A.WAVE_FORMAT_EXTENSIBLE, channels, rate,
                       rate * channels * bits // 8, channels * bits // 8, bits)
    if extensible_sub is None:
        return base
    guid_tail = bytes.fromhex('000000001000800000aa00389b71')
    return base + struct.pack('<HHI', 22, bits, 0) + struct.pack('<H', extensible_sub) + guid_tail


@pytest.mark.parametrize("bits", [8, 16, 24, 32])
def test_read_wav_pcm_widths(bits):
    rng = np.random.default_rng(bits)
    full = 1 << (bits - 1)
    v = rng.integers(-full, full, size=(50, 2))
    v[0] = [-full, full - 1]
    if bits == 8:
        raw = (v + 128).astype(np.uint8).tobytes()
    elif bits == 24:
        raw = b''.join(int(s).to_bytes(3, 'little', signed=True) for s in v.reshape(-1))
    else:
        raw = v.astype('<i%d' % (bits // 8)).tobytes()
    x, rate = A.read_wav(_riff(_fmt(A.WAVE_FORMAT_PCM, 2, 22050, bits), raw))
    assert rate == 22050 and x.shape == (50, 2) and x.dtype == np.float32
    np.testing.assert_allclose(x, v / float(full), rtol=0, atol=2.0 ** -24)
    assert x[0, 0] == -1.0 and (x.max() < 1.0 if bits <= 24 else x.max() <= 1.0)   # (2^31 - 1) / 2^31 rounds to 1.0 in float32


@pytest.mark.parametrize("bits", [32, 64])
def test_read_wav_float_and_extensible(bits):
    v = np.random.default_rng(1).uniform(-1, 1, size=(33, 3))
    raw = v.astype('<f%d' % (bits // 8)).tobytes()
    for fmt in (_fmt(A.WAVE_FORMAT_IEEE_FLOAT, 3, 48000, bits), _fmt(None, 3, 48000, bits, extensible_sub=A.WAVE_FORMAT_IEEE_FLOAT)):
        x, rate = A.read_wav(_riff(fmt, raw, extra=b'LIST' + struct.pack('<I', 3) + b'abc\x00'))   # odd-sized chunk before data
        assert rate == 48000 and x.shape == (33, 3)
        np.testing.assert_allclose(x, v.astype(np.float32), rtol=0, atol=1e-7)


def test_read_wav_matches_stdlib_wave_writer():
    v = (np.random.default_rng(2).uniform(-1, 1, size=(1000, 2)) * 32767).astype('<i2')
    buf = io.BytesIO()
    with wave.open(buf, 'wb') as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(44100)
        w.writeframes(v.tobytes())
    x, rate = A.read_wav(buf.getvalue())
    assert rate == 44100
    np.testing.assert_array_equal(x, v.astype(np.float32) / 32768.0)


def test_read_wav_errors():
    with pytest.raises(A.AudioIOError):
        A.read_wav(b'not a wav file at all')
    with pytest.raises(A.AudioIOError):
        A.read_wav(b'RIFF' + struct.pack('<I', 4) + b'WAVE')
    with pytest.raises(A.AudioIOError):
        A.read_wav(_riff(_fmt(0x0055, 1, 16000, 16), b'\x00' * 8))          # MP3-in-WAV tag
    with pytest.raises(A.AudioIOError):
        A.read_wav(_riff(_fmt(A.WAVE_FORMAT_PCM, 1, 16000, 12), b'\x00' * 8))
    x, _ = A.read_wav(_riff(_fmt(A.WAVE_FORMAT_PCM, 2, 16000, 16), b'\x01\x00\x02\x00\x03'))   # trailing partial frame dropped
    assert x.shape == (1, 2)


@pytest.mark.parametrize("orig", [44100, 48000, 22050, 8000, 32000, 11025])
def test_resample_equals_scipy_polyphase_with_the_same_filter(orig):
    ss = pytest.importorskip("scipy.signal")
    x = np.random.default_rng(orig).standard_normal(9000).astype(np.float32)
    y = A.resample(x, orig, 16000)
    g = math.gcd(orig, 16000)
    up, down = 16000 // g, orig // g
    ref = ss.resample_poly(x.astype(np.float64), up, down, window=A.resample_filter(up, down) / up)
    assert y.dtype == np.float32 and y.shape == ref.shape == (-(-9000 * up // down),)
    np.testing.assert_allclose(y, ref, rtol=0, atol=5e-7)


def test_resample_tones_and_aliasing():
    sr = 44100
    t = np.arange(2 * sr) / sr
    t16 = np.arange(-(-2 * sr * 160 // 441)) / 16000.0
    for f in (55.0, 440.0, 3520.0):
        y = A.resample((0.5 * np.sin(2 * np.pi * f * t)).astype(np.float32), sr, 16000)
        assert np.abs(y - 0.5 * np.sin(2 * np.pi * f * t16))[500:-500].max() < 2e-6
    # a 9 kHz partial lies above the new Nyquist frequency: it must vanish, not fold back to 7 kHz
    y = A.resample((0.5 * np.sin(2 * np.pi * 9000.0 * t)).astype(np.float32), sr, 16000)
    assert np.abs(y[500:-500]).max() < 2e-5
    # identity and empty input
    x = np.arange(10, dtype=np.float32)
    np.testing.assert_array_equal(A.resample(x, 16000, 16000), x)
    assert A.resample(np.zeros(0, np.float32), 44100, 16000).shape == (0,)
    with pytest.raises(ValueError):
        A.resample(np.zeros((2, 2)), 44100, 16000)


def test_wav_data_to_samples_mono_mix_and_roundtrip(tmp_path):
    sr = 48000
    t = np.arange(sr) / sr
    left, right = 0.6 * np.sin(2 * np.pi * 330 * t), 0.2 * np.sin(2 * np.pi * 660 * t)
    pcm = (np.stack([left, right], axis=1) * 32767).round().astype('<i2')
    data = _riff(_fmt(A.WAVE_FORMAT_PCM, 2, sr, 16), pcm.tobytes())
    y = A.wav_data_to_samples_librosa(data, sample_rate=16000)
    assert y.dtype == np.float32 and y.shape == (16000,)
    t16 = np.arange(16000) / 16000.0
    want = 0.5 * (0.6 * np.sin(2 * np.pi * 330 * t16) + 0.2 * np.sin(2 * np.pi * 660 * t16))
    assert np.abs(y - want)[300:-300].max() < 1e-4          # 16-bit quantisation of the source
    # samples_to_wav_data -> file -> load_audio round trip at the model rate: 16-bit quantisation only
    p = tmp_path / "a.wav"
    p.write_bytes(A.samples_to_wav_data(y, 16000))
    back = A.load_audio(str(p), 16000)
    assert back.shape == y.shape and np.abs(back - y).max() <= 1.0 / 32767 + 1e-7
    with wave.open(str(p), 'rb') as w:                       # the standard library reads what we wrote
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 16000, 16000)
