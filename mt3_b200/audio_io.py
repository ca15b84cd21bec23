"""Audio in: WAV bytes -> float32 mono samples at the model's sample rate.

The caller on the input side of boundary B1: the notebook feeds `InferenceModel.__call__` with
`note_seq.audio_io.wav_data_to_samples_librosa(data, sample_rate=16000)`
(colab/music_transcription_with_transformers.ipynb `upload_audio`), i.e. librosa.load(..., sr=sample_rate, mono=True):
decode the container, average the channels, resample.  note_seq and librosa are third-party and not installable here,
so this module restates that contract with the standard library and numpy only:

  * RIFF/WAVE decoding: PCM 8 / 16 / 24 / 32 bit, IEEE float 32 / 64, WAVE_FORMAT_EXTENSIBLE, any channel count,
    scaled to [-1, 1) like libsndfile does (int / 2^(bits-1); 8-bit is unsigned with a 128 offset);
  * mono = mean over channels (librosa.to_mono);
  * resampling by a rational factor with a Kaiser-windowed-sinc polyphase FIR (`resample`).  librosa's resampler
    (soxr / resampy, version-dependent) is a different filter of the same kind: outputs agree to the filters'
    stop-band level, not bit for bit -- the model's own frontend starts at the 16 kHz samples, where parity is defined.
"""
from __future__ import annotations

import io
import math
import struct
from typing import Tuple

import numpy as np

WAVE_FORMAT_PCM = 0x0001
WAVE_FORMAT_IEEE_FLOAT = 0x0003
WAVE_FORMAT_EXTENSIBLE = 0xFFFE


class AudioIOError(ValueError):
    """Malformed or unsupported audio data (note_seq.audio_io.AudioIOReadError)."""


def read_wav(wav_data: bytes) -> Tuple[np.ndarray, int]:
    """RIFF/WAVE bytes -> (float32 [frames, channels] in [-1, 1), native sample rate)."""
    if len(wav_data) < 12 or wav_data[:4] != b'RIFF' or wav_data[8:12] != b'WAVE':
        raise AudioIOError('not a RIFF/WAVE stream')
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(wav_data):
        cid, size = wav_data[pos:pos + 4], struct.unpack('<I', wav_data[pos + 4:pos + 8])[0]
        body = wav_data[pos + 8:pos + 8 + size]
        if cid == b'fmt ':
            fmt = body
        elif cid == b'data':
            data = body                      # a truncated data chunk is read as far as it goes (streamed files)
        pos += 8 + size + (size & 1)         # chunks are word-aligned
    if fmt is None or data is None or len(fmt) < 16:
        raise AudioIOError('WAVE stream without a fmt / data chunk')
    tag, channels, rate, _, block_align, bits = struct.unpack('<HHIIHH', fmt[:16])
    if tag == WAVE_FORMAT_EXTENSIBLE and len(fmt) >= 26:
        tag = struct.unpack('<H', fmt[24:26])[0]          # first two bytes of the sub-format GUID
    if channels < 1 or rate < 1:
        raise AudioIOError('bad channel count / sample rate')
    width = bits // 8
    n = len(data) // (width * channels) if width else 0
    raw = data[:n * width * channels]
    if tag == WAVE_FORMAT_PCM:
        if bits == 8:
            x = (np.frombuffer(raw, np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(raw, '<i2').astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(raw, np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            v = np.where(v & 0x800000, v - 0x1000000, v)
            x = v.astype(np.float32) / 8388608.0
        elif bits == 32:
            x = (np.frombuffer(raw, '<i4').astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            raise AudioIOError('unsupported PCM width: %d bits' % bits)
    elif tag == WAVE_FORMAT_IEEE_FLOAT:
        if bits == 32:
            x = np.frombuffer(raw, '<f4').astype(np.float32)
        elif bits == 64:
            x = np.frombuffer(raw, '<f8').astype(np.float32)
        else:
            raise AudioIOError('unsupported float width: %d bits' % bits)
    else:
        raise AudioIOError('unsupported WAVE format tag 0x%04x' % tag)
    return x.reshape(n, channels), int(rate)


def _kaiser_beta(atten_db: float) -> float:
    if atten_db > 50.0:
        return 0.1102 * (atten_db - 8.7)
    if atten_db >= 21.0:
        return 0.5842 * (atten_db - 21.0) ** 0.4 + 0.07886 * (atten_db - 21.0)
    return 0.0


def resample_filter(up: int, down: int, zero_crossings: int = 32, atten_db: float = 96.0) -> np.ndarray:
    """Odd-length low-pass prototype for a rate change of up/down: sinc with its cutoff at the lower of the two
    Nyquist frequencies (scaled by 0.96 so the transition band ends at Nyquist), Kaiser window, gain `up`."""
    q = max(up, down)
    half = zero_crossings * q
    n = np.arange(-half, half + 1, dtype=np.float64)
    cutoff = 0.96 / q                                   # in units of the up-sampled Nyquist frequency
    h = cutoff * np.sinc(cutoff * n) * np.kaiser(2 * half + 1, _kaiser_beta(atten_db))
    return h * (up / h.sum())


def resample(x: np.ndarray, orig_rate: int, target_rate: int, zero_crossings: int = 32) -> np.ndarray:
    """Polyphase FIR resampling of a 1-D signal: output sample m is sum_k h[m*down - k*up] x[k], i.e. zero-stuff by
    `up`, filter, keep every `down`-th sample, without ever forming the up-sampled signal.  Output length
    ceil(len * up / down) (librosa's convention).  float64 accumulation, float32 result."""
    x = np.asarray(x)
    if x.ndim != 1:
        raise ValueError('resample expects a 1-D signal')
    if orig_rate == target_rate or x.size == 0:
        return x.astype(np.float32, copy=True)
    g = math.gcd(int(orig_rate), int(target_rate))
    up, down = int(target_rate) // g, int(orig_rate) // g
    h = resample_filter(up, down, zero_crossings)
    half = (h.size - 1) // 2
    n_out = -(-x.size * up // down)
    xd = x.astype(np.float64)
    out = np.zeros(n_out, np.float64)
    # output m reads phase (m*down) mod up of the filter: group the outputs by phase, strided correlations in blocks
    for phase in range(up):
        m_all = np.arange((phase * pow(down, -1, up)) % up if up > 1 else 0, n_out, up)   # outputs with m*down = phase (mod up)
        if m_all.size == 0:
            continue
        # taps of this phase: h[half + phase + up*j] multiplies x[(m*down - phase)/up - j]
        j_lo, j_hi = -((half + phase) // up), (half - phase) // up
        taps = h[half + phase + up * np.arange(j_lo, j_hi + 1)][::-1].copy()
        block = max(1, (1 << 22) // taps.size)                           # bounds the gathered window matrix to 32 MB
        for b0 in range(0, m_all.size, block):
            m = m_all[b0:b0 + block]
            centre = (m * down - phase) // up                              # consecutive outputs of one phase: `down` apart
            lo, hi = int(centre[0] - j_hi), int(centre[-1] - j_lo)
            seg = np.pad(xd[max(lo, 0):min(hi, x.size - 1) + 1], (max(0, -lo), max(0, hi - (x.size - 1))))
            out[m] = np.lib.stride_tricks.sliding_window_view(seg, taps.size)[::down][:m.size] @ taps
    return out.astype(np.float32)


def wav_data_to_samples(wav_data: bytes, sample_rate: int) -> np.ndarray:
    """WAV bytes -> float32 mono samples at `sample_rate` (note_seq.audio_io.wav_data_to_samples_librosa)."""
    frames, native = read_wav(wav_data)
    mono = frames.mean(axis=1, dtype=np.float64).astype(np.float32) if frames.shape[1] > 1 else frames[:, 0]
    return resample(mono, native, int(sample_rate))


wav_data_to_samples_librosa = wav_data_to_samples      # the name the notebook calls


def load_audio(path: str, sample_rate: int) -> np.ndarray:
    with open(path, 'rb') as f:
        return wav_data_to_samples(f.read(), sample_rate)


def samples_to_wav_data(samples: np.ndarray, sample_rate: int) -> bytes:
    """float samples in [-1, 1] -> 16-bit PCM mono WAV bytes (note_seq.audio_io.samples_to_wav_data)."""
    pcm = np.clip(np.round(np.asarray(samples, np.float64) * 32767.0), -32768, 32767).astype('<i2').tobytes()
    out = io.BytesIO()
    out.write(b'RIFF' + struct.pack('<I', 36 + len(pcm)) + b'WAVE')
    out.write(b'fmt ' + struct.pack('<IHHIIHH', 16, WAVE_FORMAT_PCM, 1, int(sample_rate), int(sample_rate) * 2, 2, 16))
    out.write(b'data' + struct.pack('<I', len(pcm)) + pcm)
    return out.getvalue()
