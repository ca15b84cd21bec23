"""The reference's layers.py surface that is useful on its own at inference time (layers.py:85-157, :624-830).

`dot_product_attention(query, key, value, bias=None)` runs the float32 CUDA op `mt3_dot_product_attention_f32`; the
encoder / decoder use specialised kernels of the same math inside libmt3b200.so.  Dropout arguments are accepted for
signature fidelity and must be off (inference only, deterministic=True paths).

The mask helpers (`make_attention_mask`, `make_causal_mask`, `combine_masks`, `combine_biases`, `make_decoder_mask`,
`mask_to_bias`) are host-side numpy: they build the `bias` argument of the op the way MultiHeadDotProductAttention does
(layers.py:316-328).  The inference path itself needs none of them -- the encoder and cross-attention masks are all ones
(network.py:283-289, :322-326) and the decode-step mask `arange(L) <= cache_index` (layers.py:297-314) is the length
argument of the decode attention kernel."""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from . import _lib


def dot_product_attention(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, bias: Optional[torch.Tensor] = None,
                          dropout_rng=None, dropout_rate: float = 0.0, deterministic: bool = True, dtype=torch.float32,
                          float32_logits: bool = False) -> torch.Tensor:
    """softmax(q k^T + bias) v (layers.py:85-157).  query [batch, q_len, heads, d]; key / value [batch, kv_len, heads, d];
    bias broadcastable to [batch, heads, q_len, kv_len].  CUDA float32 tensors only (there is no CPU path)."""
    del float32_logits                              # everything is float32 already
    if dropout_rate != 0.0 and not deterministic:
        raise ValueError("inference only: dropout must be off")
    if dtype != torch.float32:
        raise ValueError("the reference computes in float32 (gin/model.gin:50)")
    assert key.dim() == query.dim() == value.dim() == 4, 'q, k, v must have same rank.'                 # layers.py:118
    assert query.shape[:-3] == key.shape[:-3] == value.shape[:-3], 'q, k, v batch dims must match.'      # :119-120
    assert query.shape[-2] == key.shape[-2] == value.shape[-2], 'q, k, v num_heads must match.'          # :121-122
    assert key.shape[-3] == value.shape[-3], 'k, v lengths must match.'                                  # :123
    assert query.shape[-1] == key.shape[-1], 'q, k depths must match.'                                   # :124
    for t in (query, key, value):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise TypeError("dot_product_attention needs CUDA float32 tensors")
    b, tq, h, d = query.shape
    tk = key.shape[1]
    q, k, v = query.contiguous(), key.contiguous(), value.contiguous()
    bias_c = None
    if bias is not None:
        bias_c = bias.to(device=q.device, dtype=torch.float32).expand(b, h, tq, tk).contiguous()
    out = torch.empty_like(q)
    lib = _lib.load()
    with torch.cuda.device(q.device):
        _lib.check(lib.mt3_dot_product_attention_f32(q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                                     None if bias_c is None else bias_c.data_ptr(), b, tq, tk, h, d,
                                                     out.data_ptr(), torch.cuda.current_stream(q.device).cuda_stream))
    return out


# ---------------------------------------------------------------------------------------------
# Mask-making helpers (layers.py:624-830), numpy on the host.
# ---------------------------------------------------------------------------------------------
def make_attention_mask(query_input, key_input, pairwise_fn: Callable = np.multiply, extra_batch_dims: int = 0,
                        dtype=np.float32) -> np.ndarray:
    """[..., len_q] x [..., len_kv] -> [..., 1, len_q, len_kv]: pairwise_fn of every (query, key) pair, with a singleton
    heads axis and `extra_batch_dims` leading singleton axes (layers.py:627-659)."""
    q, k = np.asarray(query_input), np.asarray(key_input)
    mask = pairwise_fn(q[..., :, None], k[..., None, :])[..., None, :, :]
    return mask.reshape((1,) * extra_batch_dims + mask.shape).astype(dtype)


def make_causal_mask(x, extra_batch_dims: int = 0, dtype=np.float32) -> np.ndarray:
    """[..., len] -> [..., 1, len, len] lower-triangular mask; depends on the shape of x only (layers.py:662-690)."""
    x = np.asarray(x)
    idxs = np.broadcast_to(np.arange(x.shape[-1], dtype=np.int32), x.shape)
    return make_attention_mask(idxs, idxs, np.greater_equal, extra_batch_dims=extra_batch_dims, dtype=dtype)


def _same_rank(arrays):
    assert all(a.ndim == arrays[0].ndim for a in arrays), f'masks must have same rank: {tuple(a.ndim for a in arrays)}'


def combine_masks(*masks, dtype=np.float32) -> Optional[np.ndarray]:
    """Logical AND of the masks that are not None; None if there is none (layers.py:693-711)."""
    present = [np.asarray(m) for m in masks if m is not None]
    if not present:
        return None
    _same_rank(present)
    return np.logical_and.reduce([m != 0 for m in present]).astype(dtype)


def combine_biases(*masks) -> Optional[np.ndarray]:
    """Sum of the biases that are not None; None if there is none (layers.py:714-731)."""
    present = [np.asarray(m) for m in masks if m is not None]
    if not present:
        return None
    _same_rank(present)
    return sum(present[1:], present[0])


def make_decoder_mask(decoder_target_tokens, dtype=np.float32, decoder_causal_attention=None,
                      decoder_segment_ids=None) -> np.ndarray:
    """Decoder self-attention mask (layers.py:734-830): causal -- made bidirectional among positions whose
    `decoder_causal_attention` is 1 (prefix LM) -- AND not-padding (token > 0 on both sides) AND same packing segment."""
    tokens = np.asarray(decoder_target_tokens)
    allowed = make_causal_mask(tokens, dtype=bool)
    if decoder_causal_attention is not None:
        prefix = np.asarray(decoder_causal_attention) != 0
        allowed = allowed | make_attention_mask(prefix, prefix, np.logical_and, dtype=bool)
    parts = [allowed, make_attention_mask(tokens > 0, tokens > 0, dtype=bool)]
    if decoder_segment_ids is not None:
        seg = np.asarray(decoder_segment_ids)
        parts.append(make_attention_mask(seg, seg, np.equal, dtype=bool))
    return combine_masks(*parts, dtype=dtype)


def mask_to_bias(mask, dtype=np.float32) -> Optional[np.ndarray]:
    """Attention mask -> additive bias: 0 where mask > 0, -1e10 elsewhere (layers.py:316-322)."""
    if mask is None:
        return None
    return np.where(np.asarray(mask) > 0, 0.0, -1e10).astype(dtype)
