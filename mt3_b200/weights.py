"""Parameter tree of the MT3 Transformer (Flax names) and its flat layout.

The C ABI takes one float32 blob; `param_shapes` lists the tree paths in blob order (the
order is also what mt3_model_param_offset reports -- tests check they agree).  A weight
file is a numpy .npz keyed by the same tree paths, so a T5X checkpoint converted to
{path: array} loads unchanged.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np


def param_shapes(cfg) -> Dict[str, Tuple[int, ...]]:
    d, q, f, v = cfg.emb_dim, cfg.num_heads * cfg.head_dim, cfg.mlp_dim, cfg.vocab_size
    depth = getattr(cfg, "input_depth", 512)
    s: Dict[str, Tuple[int, ...]] = {"encoder/continuous_inputs_projection/kernel": (depth, d)}
    for i in range(cfg.num_encoder_layers):
        p = f"encoder/layers_{i}/"
        s[p + "pre_attention_layer_norm/scale"] = (d,)
        s[p + "attention/query/kernel"] = (d, q)
        s[p + "attention/key/kernel"] = (d, q)
        s[p + "attention/value/kernel"] = (d, q)
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


def num_params(cfg) -> int:
    return sum(int(np.prod(s)) for s in param_shapes(cfg).values())


def flatten(params: Dict[str, np.ndarray], cfg) -> np.ndarray:
    """{tree path: array} -> float32 blob in ABI order; checks names and shapes."""
    shapes = param_shapes(cfg)
    missing = [k for k in shapes if k not in params]
    if missing:
        raise KeyError(f"weights are missing {len(missing)} parameters, e.g. {missing[:3]}")
    out = np.empty(num_params(cfg), np.float32)
    off = 0
    for name, shape in shapes.items():
        w = np.asarray(params[name], np.float32)
        if tuple(w.shape) != tuple(shape):
            raise ValueError(f"parameter {name}: shape {tuple(w.shape)}, expected {tuple(shape)}")
        n = w.size
        out[off:off + n] = w.reshape(-1)
        off += n
    return out


def unflatten(blob: np.ndarray, cfg) -> Dict[str, np.ndarray]:
    """Inverse of flatten(): views into the float32 blob keyed by tree path."""
    blob = np.asarray(blob, np.float32).reshape(-1)
    if blob.size != num_params(cfg):
        raise ValueError(f"weight blob has {blob.size} floats, expected {num_params(cfg)}")
    out, off = {}, 0
    for name, shape in param_shapes(cfg).items():
        n = int(np.prod(shape))
        out[name] = blob[off:off + n].reshape(shape)
        off += n
    return out


def synthetic_params(cfg, seed: int = 0) -> Dict[str, np.ndarray]:
    """Random fp32 weights drawn from the reference's initialiser distributions
    (layers.py:182-183,233-234,385-386,449-450; network.py:177,222; layers.py:608) --
    gs://mt3/checkpoints is unreachable offline, so benchmarks and tests use these."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("/scale"):
            w = np.ones(shape)
        elif name.endswith("/embedding"):
            w = rng.standard_normal(shape)
        else:
            std = 1.0 / math.sqrt(shape[0])
            if "attention/" in name:
                w = rng.standard_normal(shape) * std
                if name.endswith("query/kernel"):
                    w = w / math.sqrt(cfg.head_dim)
            else:
                x = rng.standard_normal(shape)
                bad = np.abs(x) > 2.0
                while bad.any():
                    x[bad] = rng.standard_normal(int(bad.sum()))
                    bad = np.abs(x) > 2.0
                w = x * (std / 0.87962566103423978)
        out[name] = np.ascontiguousarray(w, np.float32)
    return out


def save(path: str, params: Dict[str, np.ndarray]) -> None:
    np.savez(path, **{k.replace("/", "|"): v for k, v in params.items()})


def load(path: str) -> Dict[str, np.ndarray]:
    """A .npz written by save(), or a T5X checkpoint directory (the reference's `gs://mt3/checkpoints/<model>/`,
    notebook :247-262) read by mt3_b200.checkpoints without t5x / tensorstore."""
    import os
    if os.path.isdir(path) or os.path.basename(path) == "checkpoint":
        from . import checkpoints
        return checkpoints.load_t5x_checkpoint(path)
    with np.load(path) as z:
        return {k.replace("|", "/"): z[k] for k in z.files}
