"""Audio spectrogram functions: the reference's spectrograms.py surface on the B200 path.

Same names, defaults and argument meaning as /root/reference/mt3/spectrograms.py:23-82.
`compute_spectrogram` is drop-in boundary B2.
"""
from __future__ import annotations

import dataclasses

import numpy as np
import torch

from . import spectral_ops

# defaults for spectrogram config (spectrograms.py:23-25)
DEFAULT_SAMPLE_RATE = 16000
DEFAULT_HOP_WIDTH = 128
DEFAULT_NUM_MEL_BINS = 512

# fixed constants (spectrograms.py:27-29)
FFT_SIZE = 2048
MEL_LO_HZ = 20.0


@dataclasses.dataclass
class SpectrogramConfig:
    """Spectrogram configuration parameters (spectrograms.py:32-52)."""
    sample_rate: int = DEFAULT_SAMPLE_RATE
    hop_width: int = DEFAULT_HOP_WIDTH
    num_mel_bins: int = DEFAULT_NUM_MEL_BINS

    @property
    def abbrev_str(self):
        s = ''
        if self.sample_rate != DEFAULT_SAMPLE_RATE:
            s += 'sr%d' % self.sample_rate
        if self.hop_width != DEFAULT_HOP_WIDTH:
            s += 'hw%d' % self.hop_width
        if self.num_mel_bins != DEFAULT_NUM_MEL_BINS:
            s += 'mb%d' % self.num_mel_bins
        return s

    @property
    def frames_per_second(self):
        return self.sample_rate / self.hop_width


def split_audio(samples, spectrogram_config):
    """Split audio into hop-wide frames, zero-padding the end (spectrograms.py:55-61).
    Host-side view/copy; accepts numpy or torch."""
    hop = spectrogram_config.hop_width
    if isinstance(samples, torch.Tensor):
        n = samples.shape[-1]
        pad = (-n) % hop
        if pad:
            samples = torch.nn.functional.pad(samples, (0, pad))
        return samples.reshape(*samples.shape[:-1], -1, hop)
    samples = np.asarray(samples)
    n = samples.shape[-1]
    pad = (-n) % hop
    if pad:
        samples = np.concatenate([samples, np.zeros(samples.shape[:-1] + (pad,), samples.dtype)], axis=-1)
    return samples.reshape(*samples.shape[:-1], -1, hop)


def compute_spectrogram(samples, spectrogram_config, n_valid_frames=None, out=None):
    """Compute a log-mel spectrogram (spectrograms.py:64-73) with the fused sm_100a kernel.

    samples: CUDA float32 [n] or [S, n] -> [ceil(n/hop), bins] or [S, ceil(n/hop), bins]."""
    overlap = 1 - (spectrogram_config.hop_width / FFT_SIZE)
    return spectral_ops.compute_logmel(
        samples,
        bins=spectrogram_config.num_mel_bins,
        lo_hz=MEL_LO_HZ,
        overlap=overlap,
        fft_size=FFT_SIZE,
        sample_rate=spectrogram_config.sample_rate,
        n_valid_frames=n_valid_frames,
        out=out)


def flatten_frames(frames):
    """Convert frames back into a flat array of samples (spectrograms.py:76-78)."""
    return frames.reshape(-1)


def input_depth(spectrogram_config):
    return spectrogram_config.num_mel_bins
