"""Host side of the log-mel frontend: same names as the reference's spectral_ops.py.

`linear_to_mel_weight_matrix` builds (on the host, in float32, once) the matrix the
reference obtains from tf.signal.linear_to_mel_weight_matrix at spectral_ops.py:69-70;
`compute_logmel` runs the fused sm_100a kernel (csrc/logmel.cu) on a CUDA tensor.
"""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np
import torch

from . import _lib


def linear_to_mel_weight_matrix(num_mel_bins=20, num_spectrogram_bins=129, sample_rate=8000,
                                lower_edge_hertz=125.0, upper_edge_hertz=3800.0) -> np.ndarray:
    """HTK-mel triangular filterbank, float32 [num_spectrogram_bins, num_mel_bins]; the DC row
    is zero and triangles are linear in mel space (tf.signal.mel_ops semantics)."""
    f32 = np.float32

    def hz_to_mel(f):
        return (f32(1127.0) * np.log(f32(1.0) + np.asarray(f, f32) / f32(700.0))).astype(f32)

    nyquist = f32(sample_rate) / f32(2.0)
    linear = np.linspace(f32(0.0), nyquist, num_spectrogram_bins, dtype=f32)[1:]
    spec_mel = hz_to_mel(linear)[:, None]
    edges = np.linspace(hz_to_mel(f32(lower_edge_hertz)), hz_to_mel(f32(upper_edge_hertz)), num_mel_bins + 2, dtype=f32)
    lower, center, upper = edges[None, :-2], edges[None, 1:-1], edges[None, 2:]
    lower_slopes = (spec_mel - lower) / (center - lower)
    upper_slopes = (upper - spec_mel) / (upper - center)
    w = np.maximum(f32(0.0), np.minimum(lower_slopes, upper_slopes)).astype(f32)
    return np.ascontiguousarray(np.concatenate([np.zeros((1, num_mel_bins), f32), w], axis=0))


class _Frontend:
    """Owns one mt3_frontend handle (window, twiddles, banded mel matrix on the device)."""

    def __init__(self, sample_rate, hop, fft_size, bins, lo_hz, hi_hz, eps, device_index):
        lib = _lib.load()
        self.hop, self.bins = hop, bins
        mel = linear_to_mel_weight_matrix(bins, fft_size // 2 + 1, sample_rate, lo_hz, hi_hz)
        cfg = _lib.FrontendConfig(sample_rate, hop, fft_size, bins, eps)
        h = C.c_void_p()
        with torch.cuda.device(device_index):
            _lib.check(lib.mt3_frontend_create(C.byref(cfg), mel.ctypes.data_as(C.c_void_p), C.byref(h)))
        self.handle = h
        self._lib = lib

    def __del__(self):
        try:
            self._lib.mt3_frontend_destroy(self.handle)
        except Exception:
            pass


@functools.lru_cache(maxsize=16)
def _frontend(sample_rate, hop, fft_size, bins, lo_hz, hi_hz, eps, device_index):
    return _Frontend(sample_rate, hop, fft_size, bins, lo_hz, hi_hz, eps, device_index)


def compute_logmel(audio: torch.Tensor, lo_hz=80.0, hi_hz=7600.0, bins=64, fft_size=2048, overlap=0.75,
                   pad_end=True, sample_rate=16000, n_valid_frames: torch.Tensor | None = None,
                   out: torch.Tensor | None = None) -> torch.Tensor:
    """spectral_ops.compute_logmel (spectral_ops.py:76-88) on the GPU.

    audio: CUDA float32 [n] or [S, n] (each row an independent segment).
    Returns [T, bins] / [S, T, bins] with T = ceil(n / hop).  `n_valid_frames` (int32 [S],
    CUDA) zero-fills the rows past a short segment's end (models.py:96)."""
    if not pad_end:
        raise ValueError("compute_logmel: pad_end=False is not on the MT3 path (spectral_ops.py:35)")
    if not (isinstance(audio, torch.Tensor) and audio.is_cuda and audio.dtype == torch.float32):
        raise TypeError("compute_logmel: audio must be a CUDA float32 tensor (there is no CPU fallback)")
    hop = int(fft_size * (1.0 - overlap))
    squeeze = audio.dim() == 1
    a = audio.reshape(1, -1) if squeeze else audio
    if a.dim() != 2:
        raise ValueError(f"compute_logmel: expected [n] or [S, n], got {tuple(audio.shape)}")
    if a.stride(1) != 1:
        a = a.contiguous()
    S, n = a.shape
    T = -(-n // hop)
    fe = _frontend(int(sample_rate), hop, int(fft_size), int(bins), float(lo_hz), float(hi_hz), 1e-5, a.device.index)
    if out is None:
        out = torch.empty((S, T, bins), dtype=torch.float32, device=a.device)
    else:
        assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == S * T * bins
    if S == 0 or T == 0:          # empty input -> zero frames (tf.signal.frame), nothing to launch
        return out[0] if squeeze else out
    nv = None
    if n_valid_frames is not None:
        assert n_valid_frames.is_cuda and n_valid_frames.dtype == torch.int32 and n_valid_frames.numel() == S
        nv = n_valid_frames.contiguous()
    stream = torch.cuda.current_stream(a.device).cuda_stream
    with torch.cuda.device(a.device):
        _lib.check(fe._lib.mt3_logmel_f32(fe.handle, a.data_ptr(), a.stride(0) if S > 1 else max(n, 1), S, n,
                                          nv.data_ptr() if nv is not None else None, out.data_ptr(), stream))
    return out[0] if squeeze else out
