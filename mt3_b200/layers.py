"""The reference's layers.py surface that is useful on its own at inference time (layers.py:85-157).

`dot_product_attention(query, key, value, bias=None)` runs the float32 CUDA op `mt3_dot_product_attention_f32`; the
encoder / decoder use specialised kernels of the same math inside libmt3b200.so.  Dropout arguments are accepted for
signature fidelity and must be off (inference only, deterministic=True paths)."""
from __future__ import annotations

from typing import Optional

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
