"""Drop-in boundary B2: the public surface of the reference's `spectrograms` module
(/root/reference/mt3/spectrograms.py:23-82 -- same names, defaults and argument meaning) over the fused sm_100a
log-mel kernel.  Everything here is host-side glue; the arithmetic lives behind `mt3_logmel_f32`.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np
import torch

from . import spectral_ops

# Public constants of the reference module (its :23-29): the three configurable defaults and the two fixed values.
DEFAULT_SAMPLE_RATE, DEFAULT_HOP_WIDTH, DEFAULT_NUM_MEL_BINS = 16000, 128, 512
FFT_SIZE, MEL_LO_HZ = 2048, 20.0

# (field, default, tag) -- the tag names a non-default value in SpectrogramConfig.abbrev_str
_FIELDS = (("sample_rate", DEFAULT_SAMPLE_RATE, "sr"), ("hop_width", DEFAULT_HOP_WIDTH, "hw"),
           ("num_mel_bins", DEFAULT_NUM_MEL_BINS, "mb"))


@dataclasses.dataclass
class SpectrogramConfig:
    """What the frontend is parameterised by (reference :32-52)."""
    sample_rate: int = DEFAULT_SAMPLE_RATE
    hop_width: int = DEFAULT_HOP_WIDTH
    num_mel_bins: int = DEFAULT_NUM_MEL_BINS

    @property
    def abbrev_str(self) -> str:
        """'' for the defaults, otherwise e.g. 'sr22050hw256' (used by the reference in dataset / cache names)."""
        return "".join(f"{tag}{getattr(self, name):d}" for name, default, tag in _FIELDS if getattr(self, name) != default)

    @property
    def frames_per_second(self) -> float:
        return self.sample_rate / self.hop_width


def _pad_to_multiple(x, multiple: int):
    short = (-x.shape[-1]) % multiple
    if short == 0:
        return x
    if isinstance(x, torch.Tensor):
        return torch.nn.functional.pad(x, (0, short))
    tail = np.zeros(x.shape[:-1] + (short,), x.dtype)
    return np.concatenate([x, tail], axis=-1)


def split_audio(samples, spectrogram_config: SpectrogramConfig):
    """[..., n] samples -> [..., ceil(n / hop), hop] frames, the end zero-padded: tf.signal.frame(frame_length =
    frame_step = hop, pad_end=True) of the reference (:55-61).  numpy in, numpy out; torch in, torch out."""
    x = samples if isinstance(samples, torch.Tensor) else np.asarray(samples)
    x = _pad_to_multiple(x, spectrogram_config.hop_width)
    return x.reshape(*x.shape[:-1], -1, spectrogram_config.hop_width)


def compute_spectrogram(samples, spectrogram_config: SpectrogramConfig, n_valid_frames: Optional[torch.Tensor] = None,
                        out: Optional[torch.Tensor] = None):
    """Log-mel spectrogram of CUDA float32 samples [n] or [S, n] -> [T, bins] / [S, T, bins], T = ceil(n / hop)
    (reference :64-73: 2048-point frames, 512 HTK-mel bins from 20 Hz, log of the magnitude).  `n_valid_frames` and
    `out` are extensions for batched, pre-allocated use (see spectral_ops.compute_logmel)."""
    cfg = spectrogram_config
    return spectral_ops.compute_logmel(samples, lo_hz=MEL_LO_HZ, bins=cfg.num_mel_bins, fft_size=FFT_SIZE,
                                       overlap=1.0 - cfg.hop_width / FFT_SIZE, sample_rate=cfg.sample_rate,
                                       n_valid_frames=n_valid_frames, out=out)


def flatten_frames(frames):
    """[..., frames, hop] -> one flat run of samples (reference :76-78)."""
    return frames.reshape(-1)


def input_depth(spectrogram_config: SpectrogramConfig) -> int:
    """Feature depth the encoder sees (= number of mel bins; reference :81-82)."""
    return spectrogram_config.num_mel_bins
