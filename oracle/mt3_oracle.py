"""CPU oracle for the MT3 audio -> event-token hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this file.  The product (`mt3_b200/`) never does: it
fails loudly when its CUDA library is missing.

This is a plain-numpy restatement of the reference's algorithm, written from the
reference's call sites (citations are `file:line` into /root/reference/mt3/):

  log-mel frontend   spectral_ops.py:29-88, spectrograms.py:55-82,
                     preprocessors.py:60-78,613-618
  layer ops          layers.py:51-82 (sinusoid), :85-157 (attention),
                     :164-355 (MHDPA + KV cache), :373-418 (DenseGeneral),
                     :435-486 (MlpBlock), :489-598 (Embed/FixedEmbed),
                     :604-621 (RMS LayerNorm)
  network            network.py:44-85 (EncoderLayer), :88-155 (DecoderLayer),
                     :158-193 (Encoder), :196-262 (Decoder), :275-361
  token id contract  vocabularies.py:148-282, notebook InferenceModel._trim_eos

PARITY STATUS
  * pinned by the reference's own tests (see tests/test_oracle_kats.py):
    dot_product_attention / MHDPA numerics (layers_test.py:285-330,375-387),
    KV-cache append semantics (layers_test.py:332-373), DenseGeneral
    (layers_test.py:452-484), relu MlpBlock golden (layers_test.py:502-541),
    vocabulary decode (vocabularies_test.py:47-83), codec (event_codec_test.py).
  * PARITY UNPINNED: the log-mel frontend and the full-model logits.  Their
    arithmetic lives in third-party packages that are neither vendored nor
    pinned by the reference (`tensorflow` tf.signal.*, `flax.linen.gelu`,
    `t5x` decode loop; setup.py:39-56 lists bare names / git HEADs) and none of
    them is installable here.  Those pieces are restated from their published
    algorithms (noted "[3p]" below) and anchored on the reference's call sites.
  * CROSS-CHECKED (tests/test_oracle_crosscheck.py) against independent
    third-party implementations of the same published algorithms that ARE in this
    image -- not the reference, but a shared misreading would have to be shared
    with them: the HTK mel matrix == transformers.audio_utils.mel_filter_bank
    (triangles in mel space) to 2e-13; the STFT magnitude (periodic Hann, hop 128,
    pad_end) == torch.stft to 1e-11 at FFT 1024 / 2048 / 4096; tanh-GELU and
    un-scaled biased attention == torch; one encoder layer and one decoder layer
    (pre-RMSNorm, un-scaled attention, gated-GELU MLP) == transformers' T5 v1.1
    T5Block with the same weights to 1e-6 (its RMS statistic is float32).

Every function takes `dtype` (np.float32 = what the reference computes in,
np.float64 = the truth the CUDA path and the fp32 oracle are both judged by).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# ----------------------------------------------------------------------------
# Frontend: samples -> log-mel frames
# ----------------------------------------------------------------------------

SAMPLE_RATE = 16000      # spectrograms.py:23
HOP_WIDTH = 128          # spectrograms.py:24
NUM_MEL_BINS = 512       # spectrograms.py:25
FFT_SIZE = 2048          # spectrograms.py:28
MEL_LO_HZ = 20.0         # spectrograms.py:29
MEL_HI_HZ = 7600.0       # spectral_ops.py:79 (compute_logmel default, never overridden)


def hann_periodic(n: int, dtype=np.float32) -> np.ndarray:
    """[3p] tf.signal.hann_window(periodic=True): 0.5 - 0.5 cos(2 pi k / n)."""
    k = np.arange(n, dtype=np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)
    return w.astype(dtype)


def frame_pad_end(x: np.ndarray, frame_length: int, frame_step: int) -> np.ndarray:
    """[3p] tf.signal.frame(..., pad_end=True): ceil(n/step) frames, zeros past the end.

    x: [..., n] -> [..., ceil(n/step), frame_length]
    """
    n = x.shape[-1]
    num_frames = -(-n // frame_step)
    padded_len = (num_frames - 1) * frame_step + frame_length if num_frames > 0 else 0
    pad = max(0, padded_len - n)
    xp = np.concatenate([x, np.zeros(x.shape[:-1] + (pad,), x.dtype)], axis=-1)
    idx = np.arange(num_frames)[:, None] * frame_step + np.arange(frame_length)[None, :]
    return xp[..., idx]


def stft_mag(audio: np.ndarray, frame_size: int = FFT_SIZE, hop: int = HOP_WIDTH,
             dtype=np.float32) -> np.ndarray:
    """spectral_ops.py:35-54: |tf.signal.stft(frame_length=N, frame_step=hop, pad_end=True)|.

    Not centred, periodic Hann, fft_length = frame_length (2048 is already a power
    of two).  For dtype=float32 the FFT itself is evaluated in float64 and rounded:
    numpy has no float32 pocketfft entry point in all versions, and a float32 FFT's
    own rounding noise is what the tolerance in the tests accounts for.
    """
    audio = np.asarray(audio, dtype)
    frames = frame_pad_end(audio, frame_size, hop)            # [..., T, N]
    win = hann_periodic(frame_size, dtype)
    xw = (frames * win).astype(dtype)
    spec = np.fft.rfft(xw.astype(np.float64), n=frame_size, axis=-1)
    return np.abs(spec).astype(dtype)                          # [..., T, N/2+1]


def _hertz_to_mel(f, dtype):
    # [3p] tf.signal.mel_ops: HTK mel, 1127 ln(1 + f/700)
    return (dtype(1127.0) * np.log(dtype(1.0) + np.asarray(f, dtype) / dtype(700.0))).astype(dtype)


def linear_to_mel_weight_matrix(num_mel_bins: int = NUM_MEL_BINS,
                                num_spectrogram_bins: int = FFT_SIZE // 2 + 1,
                                sample_rate: float = SAMPLE_RATE,
                                lower_edge_hertz: float = MEL_LO_HZ,
                                upper_edge_hertz: float = MEL_HI_HZ,
                                dtype=np.float32) -> np.ndarray:
    """[3p] tf.signal.linear_to_mel_weight_matrix (called at spectral_ops.py:69-70).

    Triangles are linear in *mel* space, the DC bin row is zero, everything is
    evaluated in `dtype`.  Returns W[num_spectrogram_bins, num_mel_bins].
    """
    dt = np.dtype(dtype).type
    bands_to_zero = 1
    nyquist = dt(sample_rate) / dt(2.0)
    linear_freqs = np.linspace(dt(0.0), nyquist, num_spectrogram_bins, dtype=dtype)[bands_to_zero:]
    spec_bins_mel = _hertz_to_mel(linear_freqs, dt)[:, None]
    edges = np.linspace(_hertz_to_mel(dt(lower_edge_hertz), dt), _hertz_to_mel(dt(upper_edge_hertz), dt),
                        num_mel_bins + 2, dtype=dtype)
    lower = edges[None, :-2]
    center = edges[None, 1:-1]
    upper = edges[None, 2:]
    lower_slopes = (spec_bins_mel - lower) / (center - lower)
    upper_slopes = (upper - spec_bins_mel) / (upper - center)
    w = np.maximum(dt(0.0), np.minimum(lower_slopes, upper_slopes)).astype(dtype)
    return np.concatenate([np.zeros((bands_to_zero, num_mel_bins), dtype), w], axis=0)


def safe_log(x: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """spectral_ops.py:29-32: log(where(x <= 0, eps, x)) -- a replace, not log(x+eps)."""
    dt = x.dtype.type
    return np.log(np.where(x <= dt(0.0), dt(eps), x))


def compute_logmel(audio: np.ndarray, bins: int = NUM_MEL_BINS, lo_hz: float = MEL_LO_HZ,
                   hi_hz: float = MEL_HI_HZ, fft_size: int = FFT_SIZE, hop: int = HOP_WIDTH,
                   sample_rate: int = SAMPLE_RATE, dtype=np.float32) -> np.ndarray:
    """spectral_ops.py:57-88: mag -> tensordot(mag, mel matrix) -> safe_log.

    audio [..., n] -> [..., ceil(n/hop), bins].  The mel matrix is always built in
    float32 (TF's default) and only then cast, so the fp64 oracle uses the very
    same filterbank coefficients as the fp32 one and the CUDA path.
    """
    mag = stft_mag(audio, fft_size, hop, dtype)
    w = linear_to_mel_weight_matrix(bins, fft_size // 2 + 1, sample_rate, lo_hz, hi_hz,
                                    np.float32).astype(dtype)
    mel = np.matmul(mag, w).astype(dtype)
    return safe_log(mel)


def audio_to_frames(audio: np.ndarray, hop: int = HOP_WIDTH, sample_rate: int = SAMPLE_RATE):
    """notebook InferenceModel._audio_to_frames / preprocessors.py:60-78.

    Pads with hop - n % hop zeros (so ALWAYS 1..hop samples), splits into hop-wide
    frames, times = i / frames_per_second.
    """
    audio = np.asarray(audio, np.float32)
    audio = np.pad(audio, [0, hop - len(audio) % hop], mode="constant")
    frames = frame_pad_end(audio, hop, hop)
    num_frames = len(audio) // hop
    times = np.arange(num_frames) / (sample_rate / hop)
    return frames, times


def split_to_segments(frames: np.ndarray, times: np.ndarray, inputs_length: int):
    """[3p] t5.data.preprocessors.split_tokens_to_inputs_length at the notebook call
    site: consecutive, non-overlapping chunks of `inputs_length` frames; the last
    one may be short (and is dropped only if empty)."""
    segs = []
    for s in range(0, frames.shape[0], inputs_length):
        segs.append((frames[s:s + inputs_length], times[s:s + inputs_length]))
    return segs


def compute_spectrogram(samples: np.ndarray, dtype=np.float32) -> np.ndarray:
    """spectrograms.py:64-73 (via preprocessors.py:613-618, per segment)."""
    return compute_logmel(samples, NUM_MEL_BINS, MEL_LO_HZ, MEL_HI_HZ, FFT_SIZE, HOP_WIDTH,
                          SAMPLE_RATE, dtype)


def pad_inputs(spec: np.ndarray, inputs_length: int) -> np.ndarray:
    """models.py:96 (seqio _pack_or_pad) [3p]: trim/pad rows to inputs_length with 0.0."""
    out = np.zeros((inputs_length, spec.shape[1]), spec.dtype)
    t = min(inputs_length, spec.shape[0])
    out[:t] = spec[:t]
    return out


# ----------------------------------------------------------------------------
# Model config / parameters
# ----------------------------------------------------------------------------

@dataclass
class T5Config:
    """network.py:25-41 with the values bound in gin/model.gin:47-59."""
    vocab_size: int = 1536
    emb_dim: int = 512
    num_heads: int = 6
    num_encoder_layers: int = 8
    num_decoder_layers: int = 8
    head_dim: int = 64
    mlp_dim: int = 1024
    mlp_activations: Tuple[str, ...] = ("gelu", "linear")
    input_depth: int = NUM_MEL_BINS

    @property
    def qkv_dim(self) -> int:
        return self.num_heads * self.head_dim


def param_shapes(cfg: T5Config) -> Dict[str, Tuple[int, ...]]:
    """Flax parameter tree; kernels stored 2-D (layers.py:406-415)."""
    d, q, f, v = cfg.emb_dim, cfg.qkv_dim, cfg.mlp_dim, cfg.vocab_size
    s: Dict[str, Tuple[int, ...]] = {}
    s["encoder/continuous_inputs_projection/kernel"] = (cfg.input_depth, d)
    for i in range(cfg.num_encoder_layers):
        p = f"encoder/layers_{i}/"
        s[p + "pre_attention_layer_norm/scale"] = (d,)
        for n in ("query", "key", "value"):
            s[p + f"attention/{n}/kernel"] = (d, q)
        s[p + "attention/out/kernel"] = (q, d)
        s[p + "pre_mlp_layer_norm/scale"] = (d,)
        s[p + "mlp/wi_0/kernel"] = (d, f)
        s[p + "mlp/wi_1/kernel"] = (d, f)
        s[p + "mlp/wo/kernel"] = (f, d)
    s["encoder/encoder_norm/scale"] = (d,)
    s["decoder/token_embedder/embedding"] = (v, d)
    for i in range(cfg.num_decoder_layers):
        p = f"decoder/layers_{i}/"
        s[p + "pre_self_attention_layer_norm/scale"] = (d,)
        for n in ("query", "key", "value"):
            s[p + f"self_attention/{n}/kernel"] = (d, q)
        s[p + "self_attention/out/kernel"] = (q, d)
        s[p + "pre_cross_attention_layer_norm/scale"] = (d,)
        for n in ("query", "key", "value"):
            s[p + f"encoder_decoder_attention/{n}/kernel"] = (d, q)
        s[p + "encoder_decoder_attention/out/kernel"] = (q, d)
        s[p + "pre_mlp_layer_norm/scale"] = (d,)
        s[p + "mlp/wi_0/kernel"] = (d, f)
        s[p + "mlp/wi_1/kernel"] = (d, f)
        s[p + "mlp/wo/kernel"] = (f, d)
    s["decoder/decoder_norm/scale"] = (d,)
    s["decoder/logits_dense/kernel"] = (d, v)
    return s


def _trunc_normal(rng, shape, std):
    # variance_scaling(..., 'truncated_normal'): N(0,1) truncated to [-2,2], rescaled so
    # that the *truncated* distribution has the requested std (flax divides by .8796...).
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2.0
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2.0
    return x * (std / 0.87962566103423978)


def init_params(cfg: T5Config, seed: int = 0, norm_scale_jitter: float = 0.0) -> Dict[str, np.ndarray]:
    """Synthetic fp32 weights with the reference's initialiser *distributions*:
    dense kernels truncated-normal fan-in (layers.py:385-386,449-450); attention
    kernels normal fan-in with the query kernel further / sqrt(head_dim)
    (layers.py:182-183,233-234); embedding N(0,1) (network.py:222); input
    projection lecun-normal (network.py:177); norm scales 1 (layers.py:608).
    `norm_scale_jitter` > 0 perturbs the norm scales so tests exercise them.
    """
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("/scale"):
            w = np.ones(shape)
            if norm_scale_jitter > 0.0:
                w = w + norm_scale_jitter * rng.standard_normal(shape)
        elif name.endswith("/embedding"):
            w = rng.standard_normal(shape)
        else:
            fan_in = shape[0]
            std = 1.0 / math.sqrt(fan_in)
            if "attention/" in name:
                w = rng.standard_normal(shape) * std
                if name.endswith("query/kernel"):
                    w = w / math.sqrt(cfg.head_dim)
            else:
                w = _trunc_normal(rng, shape, std)
        out[name] = np.ascontiguousarray(w, np.float32)
    return out


# ----------------------------------------------------------------------------
# Layer ops
# ----------------------------------------------------------------------------

def sinusoidal_table(max_len: int = 2048, features: int = 512, min_scale: float = 1.0,
                     max_scale: float = 10000.0) -> np.ndarray:
    """layers.py:51-82; float32 result exactly as the reference builds it."""
    pe = np.zeros((max_len, features), dtype=np.float32)
    position = np.arange(0, max_len)[:, np.newaxis]
    scale_factor = -np.log(max_scale / min_scale) / (features // 2 - 1)
    div_term = min_scale * np.exp(np.arange(0, features // 2) * scale_factor)
    pe[:, :features // 2] = np.sin(position * div_term)
    pe[:, features // 2:2 * (features // 2)] = np.cos(position * div_term)
    return pe


def rms_norm(x: np.ndarray, scale: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    """layers.py:604-621: x * rsqrt(mean(x^2) + eps) * scale (no mean subtraction)."""
    dt = x.dtype.type
    mean2 = np.mean(np.square(x), axis=-1, keepdims=True)
    return (x * (dt(1.0) / np.sqrt(mean2 + dt(eps)))) * scale.astype(x.dtype)


def dense(x: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """layers.py:373-418 with axis=-1: x @ kernel, no bias."""
    return np.matmul(x, kernel.astype(x.dtype))


def softmax(x: np.ndarray, axis: int = -1) -> np.ndarray:
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / np.sum(e, axis=axis, keepdims=True)


def dot_product_attention(query, key, value, bias=None):
    """layers.py:85-157 (deterministic): q,k,v [b, len, h, d]; bias broadcastable to
    [b, h, q, k].  Logits are NOT scaled by 1/sqrt(d) (layers.py:230-234)."""
    w = np.einsum("bqhd,bkhd->bhqk", query, key)
    if bias is not None:
        w = w + bias.astype(w.dtype)
    w = softmax(w, -1)
    return np.einsum("bhqk,bkhd->bqhd", w, value)


def gelu_tanh(x: np.ndarray) -> np.ndarray:
    """[3p] flax.linen.gelu default (approximate=True)."""
    dt = x.dtype.type
    c = dt(math.sqrt(2.0 / math.pi))
    return dt(0.5) * x * (dt(1.0) + np.tanh(c * (x + dt(0.044715) * x * x * x)))


_ACT = {
    "linear": lambda x: x,
    "relu": lambda x: np.maximum(x, x.dtype.type(0)),
    "gelu": gelu_tanh,
}


def mlp_block(x, wi: Sequence[np.ndarray], wo: np.ndarray, activations: Sequence[str]):
    """layers.py:435-486: prod_i act_i(x @ wi_i) @ wo."""
    hs = [_ACT[a](dense(x, w)) for a, w in zip(activations, wi)]
    h = hs[0]
    for o in hs[1:]:
        h = h * o
    return dense(h, wo)


def mask_to_bias(mask: Optional[np.ndarray], dtype) -> Optional[np.ndarray]:
    """layers.py:317-322: mask>0 -> 0, else -1e10."""
    if mask is None:
        return None
    return np.where(mask > 0, dtype(0.0), dtype(-1e10)).astype(dtype)


@dataclass
class KVCache:
    """layers.py:255-260: cached_key/value [b, h, d, L] (length LAST), cache_index."""
    cached_key: np.ndarray
    cached_value: np.ndarray
    cache_index: int = 0


def mha(params: Dict[str, np.ndarray], prefix: str, inputs_q, inputs_kv, num_heads: int, head_dim: int,
        mask=None, bias=None, cache: Optional[KVCache] = None):
    """layers.py:164-355.  With `cache` set this is the decode=True branch
    (:246-314): one query position, K/V written into slot cache_index by the
    one-hot multiply-add, index incremented, mask arange(L) <= index."""
    dt = inputs_q.dtype.type
    b = inputs_q.shape[0]
    q = dense(inputs_q, params[prefix + "query/kernel"]).reshape(b, -1, num_heads, head_dim)
    k = dense(inputs_kv, params[prefix + "key/kernel"]).reshape(b, -1, num_heads, head_dim)
    v = dense(inputs_kv, params[prefix + "value/kernel"]).reshape(b, -1, num_heads, head_dim)
    if cache is not None:
        bb, h, d, length = cache.cached_key.shape
        if q.shape != (bb, 1, h, d):
            raise ValueError("Autoregressive cache shape error, expected query shape %s instead got %s."
                             % ((bb, 1, h, d), q.shape))
        cur = cache.cache_index
        one_hot = (np.arange(length) == cur).astype(k.dtype)
        one_k = np.moveaxis(k, -3, -1)          # [b,h,d,1]
        one_v = np.moveaxis(v, -3, -1)
        key = cache.cached_key + one_k * one_hot
        val = cache.cached_value + one_v * one_hot
        cache.cached_key, cache.cached_value = key, val
        cache.cache_index = cur + 1
        k = np.moveaxis(key, -1, -3)            # [b,L,h,d]
        v = np.moveaxis(val, -1, -3)
        causal = np.broadcast_to(np.arange(length) <= cur, (bb, 1, 1, length)).astype(k.dtype)
        mask = causal if mask is None else np.logical_and(mask, causal).astype(k.dtype)
    attn_bias = mask_to_bias(mask, dt)
    if bias is not None:
        attn_bias = bias if attn_bias is None else attn_bias + bias
    x = dot_product_attention(q, k, v, attn_bias)           # [b,q,h,d]
    x = x.reshape(b, x.shape[1], num_heads * head_dim)
    return dense(x, params[prefix + "out/kernel"])


# ----------------------------------------------------------------------------
# Network
# ----------------------------------------------------------------------------

def _cast(params, dtype):
    return {k: v.astype(dtype) for k, v in params.items()}


def encode(params: Dict[str, np.ndarray], cfg: T5Config, x: np.ndarray, dtype=np.float32,
           return_layers: bool = False):
    """network.py:158-193,275-301 with dropout off.  x [B,T,input_depth] -> [B,T,emb]."""
    p = _cast(params, dtype)
    x = np.asarray(x, dtype)
    t = x.shape[1]
    pe = sinusoidal_table(2048, cfg.emb_dim).astype(dtype)
    h = dense(x, p["encoder/continuous_inputs_projection/kernel"]) + pe[None, :t]
    layers = [h]
    for i in range(cfg.num_encoder_layers):
        pre = f"encoder/layers_{i}/"
        a = rms_norm(h, p[pre + "pre_attention_layer_norm/scale"])
        # encoder mask is all ones (network.py:283-289) -> bias of zeros
        a = mha(p, pre + "attention/", a, a, cfg.num_heads, cfg.head_dim,
                mask=np.ones((x.shape[0], 1, t, t), dtype))
        h = h + a
        m = rms_norm(h, p[pre + "pre_mlp_layer_norm/scale"])
        m = mlp_block(m, [p[pre + "mlp/wi_0/kernel"], p[pre + "mlp/wi_1/kernel"]], p[pre + "mlp/wo/kernel"],
                      cfg.mlp_activations)
        h = h + m
        layers.append(h)
    out = rms_norm(h, p["encoder/encoder_norm/scale"])
    return (out, layers) if return_layers else out


def decode_teacher_forced(params, cfg: T5Config, encoded: np.ndarray, decoder_input_tokens: np.ndarray,
                          dtype=np.float32) -> np.ndarray:
    """network.py:196-262 over a whole token sequence with a causal mask.

    At inference the reference decodes one position at a time (decode=True); this
    full-sequence form computes the same logits and is what the step-wise oracle
    and the CUDA decode step are cross-checked against.  [B,L] int -> [B,L,V]."""
    p = _cast(params, dtype)
    enc = np.asarray(encoded, dtype)
    b, length = decoder_input_tokens.shape
    pe = sinusoidal_table(2048, cfg.emb_dim).astype(dtype)
    y = p["decoder/token_embedder/embedding"][decoder_input_tokens] + pe[None, :length]
    causal = np.tril(np.ones((length, length), dtype))[None, None]
    for i in range(cfg.num_decoder_layers):
        pre = f"decoder/layers_{i}/"
        a = rms_norm(y, p[pre + "pre_self_attention_layer_norm/scale"])
        a = mha(p, pre + "self_attention/", a, a, cfg.num_heads, cfg.head_dim, mask=causal)
        y = y + a
        c = rms_norm(y, p[pre + "pre_cross_attention_layer_norm/scale"])
        c = mha(p, pre + "encoder_decoder_attention/", c, enc, cfg.num_heads, cfg.head_dim)
        y = y + c
        m = rms_norm(y, p[pre + "pre_mlp_layer_norm/scale"])
        m = mlp_block(m, [p[pre + "mlp/wi_0/kernel"], p[pre + "mlp/wi_1/kernel"]], p[pre + "mlp/wo/kernel"],
                      cfg.mlp_activations)
        y = y + m
    y = rms_norm(y, p["decoder/decoder_norm/scale"])
    return dense(y, p["decoder/logits_dense/kernel"])


@dataclass
class DecodeState:
    """The reference's mutable 'cache' collection (network.py:303-361): per decoder
    layer a KVCache, plus FixedEmbed's position_embedder_index (layers.py:589-596),
    which the T5X cache-initialisation pass leaves at 0 [3p]."""
    self_cache: List[KVCache]
    position_index: int = 0
    cross_kv: Optional[List[Tuple[np.ndarray, np.ndarray]]] = None   # hoisted, optional


def init_decode_state(cfg: T5Config, batch: int, max_decode_length: int, dtype=np.float32) -> DecodeState:
    z = lambda: np.zeros((batch, cfg.num_heads, cfg.head_dim, max_decode_length), dtype)
    return DecodeState([KVCache(z(), z(), 0) for _ in range(cfg.num_decoder_layers)], 0)


def decode_step(params_cast, cfg: T5Config, encoded: np.ndarray, tokens: np.ndarray, state: DecodeState,
                hoist_cross_kv: bool = True) -> np.ndarray:
    """One decode=True step (network.py:303-361 -> :196-262 -> :88-155).

    tokens int[B] -> logits [B,V].  `params_cast` must already be in the compute
    dtype.  The reference re-projects the 256 encoder positions to K/V on every
    step (no decode= flag on the cross attention, network.py:129-135);
    hoist_cross_kv=True computes them once -- identical values.
    """
    p = params_cast
    dtype = encoded.dtype
    pe = sinusoidal_table(2048, cfg.emb_dim).astype(dtype)
    y = p["decoder/token_embedder/embedding"][tokens][:, None, :] + pe[state.position_index][None, None, :]
    state.position_index += 1
    b = tokens.shape[0]
    if hoist_cross_kv and state.cross_kv is None:
        state.cross_kv = []
        for i in range(cfg.num_decoder_layers):
            pre = f"decoder/layers_{i}/encoder_decoder_attention/"
            state.cross_kv.append((dense(encoded, p[pre + "key/kernel"]), dense(encoded, p[pre + "value/kernel"])))
    for i in range(cfg.num_decoder_layers):
        pre = f"decoder/layers_{i}/"
        a = rms_norm(y, p[pre + "pre_self_attention_layer_norm/scale"])
        a = mha(p, pre + "self_attention/", a, a, cfg.num_heads, cfg.head_dim, cache=state.self_cache[i])
        y = y + a
        c = rms_norm(y, p[pre + "pre_cross_attention_layer_norm/scale"])
        if hoist_cross_kv:
            q = dense(c, p[pre + "encoder_decoder_attention/query/kernel"]).reshape(b, 1, cfg.num_heads, cfg.head_dim)
            k, v = state.cross_kv[i]
            k = k.reshape(b, -1, cfg.num_heads, cfg.head_dim)
            v = v.reshape(b, -1, cfg.num_heads, cfg.head_dim)
            o = dot_product_attention(q, k, v, None).reshape(b, 1, cfg.qkv_dim)
            c = dense(o, p[pre + "encoder_decoder_attention/out/kernel"])
        else:
            c = mha(p, pre + "encoder_decoder_attention/", c, encoded, cfg.num_heads, cfg.head_dim)
        y = y + c
        m = rms_norm(y, p[pre + "pre_mlp_layer_norm/scale"])
        m = mlp_block(m, [p[pre + "mlp/wi_0/kernel"], p[pre + "mlp/wi_1/kernel"]], p[pre + "mlp/wo/kernel"],
                      cfg.mlp_activations)
        y = y + m
    y = rms_norm(y, p["decoder/decoder_norm/scale"])
    return dense(y, p["decoder/logits_dense/kernel"])[:, 0, :]


EOS_ID = 1          # vocabularies.py:157-159
PAD_ID = 0


def greedy_decode(params, cfg: T5Config, encoded: np.ndarray, max_decode_length: int = 1024,
                  dtype=np.float32, stop_at_eos: bool = True, hoist_cross_kv: bool = True,
                  forced_tokens: Optional[np.ndarray] = None, return_logits: bool = False):
    """Greedy loop standing in for t5x decoding with num_decodes=1 [3p] (models.py:127;
    SURVEY D7 records that T5X's beam-1 is not always identical to greedy).

    Start token 0; position p's logits pick token p; a sequence that emitted EOS
    keeps emitting PAD(0); the loop ends when every sequence has finished or at
    max_decode_length.  Returns int32 [B, max_decode_length] (zeros after EOS).
    `forced_tokens` [B, L] teacher-forces the *inputs* (for logit comparisons).
    """
    p = _cast(params, dtype)
    enc = np.asarray(encoded, dtype)
    b = enc.shape[0]
    state = init_decode_state(cfg, b, max_decode_length, dtype)
    out = np.zeros((b, max_decode_length), np.int32)
    logits_all = []
    cur = np.zeros((b,), np.int64)
    finished = np.zeros((b,), bool)
    for step in range(max_decode_length):
        logits = decode_step(p, cfg, enc, cur, state, hoist_cross_kv)
        if return_logits:
            logits_all.append(logits)
        nxt = np.argmax(logits, axis=-1)
        nxt = np.where(finished, PAD_ID, nxt)
        out[:, step] = nxt
        finished |= nxt == EOS_ID
        if forced_tokens is not None:
            if step + 1 < forced_tokens.shape[1]:
                cur = forced_tokens[:, step + 1].astype(np.int64)
        else:
            cur = nxt
        if stop_at_eos and finished.all():
            break
    if return_logits:
        return out, np.stack(logits_all, axis=1)
    return out


# ----------------------------------------------------------------------------
# Token id contract
# ----------------------------------------------------------------------------

DECODED_EOS_ID = -1       # vocabularies.py:30
DECODED_INVALID_ID = -2   # vocabularies.py:31
NUM_SPECIAL_TOKENS = 3    # vocabularies.py:153
DEFAULT_EXTRA_IDS = 100   # [3p] t5.data.DEFAULT_EXTRA_IDS at vocabularies.py:145


def codec_num_classes(num_velocity_bins: int, steps_per_second: int = 100, max_shift_seconds: int = 10) -> int:
    """vocabularies.py:119-140 + event_codec.py:63-64: shift | pitch | velocity | tie | program | drum."""
    return (steps_per_second * max_shift_seconds + 1) + 128 + (num_velocity_bins + 1) + 1 + 128 + 128


def num_embeddings(num_classes: int, extra_ids: int = DEFAULT_EXTRA_IDS) -> int:
    """vocabularies.py:280-282: vocab size padded up to a multiple of 128."""
    return 128 * math.ceil((NUM_SPECIAL_TOKENS + num_classes + extra_ids) / 128)


def vocab_decode(ids: np.ndarray, num_regular_tokens: int) -> np.ndarray:
    """vocabularies.py:241-271 (_decode_tf): id-3; first EOS and everything after -> -1;
    ids < 3 or >= 3 + num_regular -> -2."""
    ids = np.asarray(ids)
    eos_and_after = np.cumsum((ids == EOS_ID).astype(np.int32), axis=-1) > 0
    valid = (ids >= NUM_SPECIAL_TOKENS) & (ids < NUM_SPECIAL_TOKENS + num_regular_tokens)
    return np.where(eos_and_after, DECODED_EOS_ID,
                    np.where(valid, ids - NUM_SPECIAL_TOKENS, DECODED_INVALID_ID)).astype(np.int32)


def trim_eos(tokens: np.ndarray) -> np.ndarray:
    """notebook InferenceModel._trim_eos / tasks.py:58-63."""
    tokens = np.array(tokens, np.int32)
    if DECODED_EOS_ID in tokens:
        tokens = tokens[:np.argmax(tokens == DECODED_EOS_ID)]
    return tokens


# ----------------------------------------------------------------------------
# Synthetic audio
# ----------------------------------------------------------------------------

def sine_mix(num_samples: int, seed: int, sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """Seeded sum of 3..8 sinusoids at MIDI pitches 36..96, peak-normalised to 0.9."""
    rng = np.random.default_rng(seed)
    k = int(rng.integers(3, 9))
    t = np.arange(num_samples, dtype=np.float64) / sample_rate
    x = np.zeros(num_samples, np.float64)
    for _ in range(k):
        pitch = int(rng.integers(36, 97))
        f = 440.0 * 2.0 ** ((pitch - 69) / 12.0)
        amp = rng.uniform(0.05, 0.3)
        ph = rng.uniform(0.0, 2.0 * np.pi)
        x += amp * np.sin(2.0 * np.pi * f * t + ph)
    x *= 0.9 / max(1e-12, np.max(np.abs(x)))
    return x.astype(np.float32)
