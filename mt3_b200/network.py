"""T5-style encoder/decoder of MT3 on sm_100a: the reference's network.py surface.

`T5Config` mirrors network.py:25-41; `Transformer` exposes encode / decode (single-step,
decode=True) like the Flax module (network.py:265-361), plus `generate`, which stands in
for t5x's predict_batch_with_aux at num_decodes=1 (models.py:121-138).  All compute runs in
libmt3b200.so; tensors are only the memory container.  Drop-in boundary B3.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib, weights as W


@dataclasses.dataclass
class T5Config:
    """Global hyperparameters (network.py:25-41).  Defaults are the reference's class
    defaults; gin/model.gin:47-59 binds the MT3 values."""
    vocab_size: int
    dtype: str = 'float32'
    emb_dim: int = 512
    num_heads: int = 8
    num_encoder_layers: int = 6
    num_decoder_layers: int = 6
    head_dim: int = 64
    mlp_dim: int = 2048
    mlp_activations: Sequence[str] = ('relu',)
    dropout_rate: float = 0.1
    logits_via_embedding: bool = False
    input_depth: int = 512   # spectrograms.input_depth (models.py:135; gin/model.gin:43)


class Transformer:
    """An encoder-decoder Transformer model (network.py:265-409), inference only."""

    def __init__(self, config: T5Config, params: Dict[str, np.ndarray], *, device="cuda:0", max_batch: int = 8,
                 max_input_length: int = 256, max_decode_length: int = 1024, gemm_mode: int = _lib.GEMM_FP32_SIMT,
                 kv_format: int = _lib.KV_F32):
        if tuple(config.mlp_activations) != ('gelu', 'linear'):
            raise ValueError("only the gated-GELU MLP ('gelu','linear') of gin/model.gin:57 is built; got %r"
                             % (tuple(config.mlp_activations),))
        if config.logits_via_embedding:
            raise ValueError("logits_via_embedding=True is not on the MT3 path (gin/model.gin:59)")
        if config.dtype != 'float32':
            raise ValueError("the reference computes in float32 (gin/model.gin:50); got dtype=%r" % (config.dtype,))
        if not torch.cuda.is_available():
            raise RuntimeError("mt3_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.config = config
        self.device = torch.device(device)
        self.max_decode_length = int(max_decode_length)
        self._lib = _lib.load()
        self._cfg = _lib.ModelConfig(config.vocab_size, config.emb_dim, config.num_heads, config.head_dim,
                                     config.num_encoder_layers, config.num_decoder_layers, config.mlp_dim,
                                     config.input_depth, int(max_batch), int(max_input_length),
                                     int(max_decode_length), int(gemm_mode), int(kv_format))
        n = int(self._lib.mt3_model_num_params(C.byref(self._cfg)))
        if n != W.num_params(config):
            raise RuntimeError(f"parameter count mismatch: library {n}, host {W.num_params(config)}")
        blob = torch.from_numpy(W.flatten(params, config)).to(self.device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mt3_model_create(C.byref(self._cfg), blob.data_ptr(), C.byref(h), self._stream()))
            torch.cuda.synchronize(self.device)
        del blob
        self._h = h
        self._ws = None
        self._ws_key = None

    # -- plumbing ---------------------------------------------------------------------
    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.mt3_model_destroy(self._h)
        except Exception:
            pass

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _bind(self, batch: int, length: int):
        key = (batch, length)
        if self._ws_key == key:
            return
        nbytes = int(self._lib.mt3_workspace_bytes(self._h, batch, length))
        if nbytes <= 0:
            raise _lib.Mt3Error(-2, f"no workspace layout for batch={batch}, input_length={length}")
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        _lib.check(self._lib.mt3_model_set_workspace(self._h, self._ws.data_ptr(), self._ws.numel(), batch, length))
        self._ws_key = key

    def _check_inputs(self, x: torch.Tensor):
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32):
            raise TypeError("encoder_input_tokens must be a CUDA float32 tensor")
        assert x.dim() == 3, "encoder_input_tokens.ndim == 3  # (batch, length, depth) -- network.py:281"
        if x.shape[-1] != self.config.input_depth:
            raise ValueError(f"input depth {x.shape[-1]} != {self.config.input_depth}")
        return x.contiguous()

    # -- reference surface ------------------------------------------------------------
    def encode(self, encoder_input_tokens: torch.Tensor, encoder_segment_ids=None, enable_dropout: bool = False):
        """Transformer.encode (network.py:275-301): [B,T,depth] -> [B,T,emb]."""
        if enable_dropout:
            raise ValueError("inference only: enable_dropout must be False")
        if encoder_segment_ids is not None:
            raise ValueError("packing (encoder_segment_ids) is a training feature; not on this path")
        x = self._check_inputs(encoder_input_tokens)
        b, t, _ = x.shape
        self._bind(b, t)
        out = torch.empty((b, t, self.config.emb_dim), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mt3_encode(self._h, x.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def init_cache(self, encoded: torch.Tensor):
        """The cache-initialisation pass of t5x's predict_batch_with_aux: projects `encoded` to
        every decoder layer's cross K/V once, zeroes cache_index / position_embedder_index."""
        assert encoded.is_cuda and encoded.dtype == torch.float32 and encoded.dim() == 3
        b, t, _ = encoded.shape
        self._bind(b, t)
        enc = encoded.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mt3_cross_kv(self._h, enc.data_ptr(), self._stream()))

    def decode(self, encoded, encoder_input_tokens, decoder_input_tokens, decoder_target_tokens=None,
               encoder_segment_ids=None, decoder_segment_ids=None, decoder_positions=None, enable_dropout=False,
               decode=True, max_decode_length=None):
        """One Transformer.decode(decode=True) step (network.py:303-361): decoder_input_tokens
        int32 [B,1] -> logits float32 [B,1,V]; the KV cache lives in the model handle
        (call init_cache(encoded) first).  The teacher-forced decode=False form is training-only."""
        if not decode:
            raise NotImplementedError("decode=False (teacher-forced training pass) is out of scope; "
                                      "use teacher_forced_logits for evaluation")
        if encoder_segment_ids is not None:
            raise ValueError('During decoding, packing should not be used but `encoder_segment_ids` was passed '
                             'to `Transformer.decode`.')
        tok = decoder_input_tokens
        if not isinstance(tok, torch.Tensor) or tok.dtype not in (torch.int32, torch.int64):
            raise ValueError('Input type must be an integer or unsigned integer.')   # layers.py:528-529
        assert tok.dim() == 2  # [batch, len] -- network.py:211
        b = tok.shape[0]
        if tok.shape[1] != 1:
            raise ValueError('Autoregressive cache shape error, expected query shape %s instead got %s.'
                             % ((b, 1), tuple(tok.shape)))                           # layers.py:266-270
        tok = tok.to(torch.int32).contiguous()
        logits = torch.empty((b, 1, self.config.vocab_size), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mt3_decode_step(self._h, tok.data_ptr(), logits.data_ptr(), None, self._stream()))
        return logits

    # -- helpers built on the step ----------------------------------------------------
    def teacher_forced_logits(self, encoded: torch.Tensor, decoder_input_tokens: torch.Tensor) -> torch.Tensor:
        """Feeds decoder_input_tokens [B,L] one position at a time -> logits [B,L,V]."""
        self.init_cache(encoded)
        outs = [self.decode(encoded, None, decoder_input_tokens[:, i:i + 1]) for i in range(decoder_input_tokens.shape[1])]
        return torch.cat(outs, dim=1)

    def generate(self, encoder_input_tokens: torch.Tensor, num_steps: Optional[int] = None, stop_at_eos: bool = True,
                 use_graph: bool = True, out: Optional[torch.Tensor] = None, decode: str = 'greedy') -> torch.Tensor:
        """encode + decode from BOS=0 (models.py:121-138).  decode='greedy': argmax until EOS; decode='beam1': T5X's
        decoding.beam_search at num_decodes=1, the reference's decode_fn (models.py:127) -- they differ when EOS is one of the
        two best tokens without being decisive (tests/test_beam1.py).  Returns raw model ids int32 [B, max_decode_length],
        0 after EOS."""
        if decode not in ('greedy', 'beam1'):
            raise ValueError("decode must be 'greedy' or 'beam1'")
        x = self._check_inputs(encoder_input_tokens)
        b, t, _ = x.shape
        self._bind(b, t)
        steps = self.max_decode_length if num_steps is None else int(num_steps)
        if out is None:
            out = torch.empty((b, self.max_decode_length), dtype=torch.int32, device=self.device)
        flags = (_lib.GEN_STOP_AT_EOS if stop_at_eos else 0) | (_lib.GEN_USE_GRAPH if use_graph else 0) | \
                (_lib.GEN_BEAM1 if decode == 'beam1' else 0)
        ran = C.c_int32(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.mt3_generate(self._h, x.data_ptr(), steps, flags, out.data_ptr(), C.byref(ran), self._stream()))
        self.last_steps_run = int(ran.value)
        return out
