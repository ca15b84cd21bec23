"""TEST INFRASTRUCTURE (not collected by pytest): CPU study of the K/V row storage formats.

    python tests/kv_format_study.py [decode_length] [small|full]

The float64 oracle's teacher-forced decoder is re-run with the K and V rows of every decoder layer rounded the way
`mt3_model_config.kv_cache_format` stores them (everything else float64), for diffuse attention (the oracle's
random-init weights) and for sharp attention (decoder query kernels scaled by 4 / 8 / 16).  It is the source of the
error table in DESIGN.md section 4 and of the numbers quoted by `test_kv_cache_formats_sharp_attention`; the GPU tests
measure the same quantities through the C-ABI.  Uses oracle/ as the checker only.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import mt3_oracle as O  # noqa: E402


def q_f32(x):
    return np.asarray(x, np.float32).astype(np.float64)


def q_f16(x):
    return np.asarray(x, np.float16).astype(np.float64)


def q_p24(x):
    """MT3_KV_P24 as gemm_simt.cuh::p24_encode / p24_decode define it: value bits = hi16 : lo8 : lo8 (the third byte is
    repeated as the fourth so that one byte permute rebuilds the float); the encoder takes the nearest such value."""
    bits = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.int64)
    t = bits >> 8
    best, err = None, None
    for cand in (t, t + 1, np.where((t & 0x7FFFFF) != 0, t - 1, t)):
        v = (cand << 8) | (cand & 255)
        e = np.abs(v - bits)
        if best is None:
            best, err = v, e
        else:
            best = np.where(e < err, v, best)
            err = np.minimum(e, err)
    return best.astype(np.uint32).view(np.float32).astype(np.float64)


def teacher_forced(params, cfg, enc, toks, qk, qv):
    """oracle decode_teacher_forced (network.py:88-155) in float64 with the stored K / V rows passed through qk / qv."""
    dtype = np.float64
    p = O._cast(params, dtype)
    enc = np.asarray(enc, dtype)
    b, length = toks.shape
    pe = O.sinusoidal_table(2048, cfg.emb_dim).astype(dtype)
    y = p["decoder/token_embedder/embedding"][toks] + pe[None, :length]
    causal = np.tril(np.ones((length, length), dtype))[None, None]
    H, d = cfg.num_heads, cfg.head_dim

    def att(pre, xq, xkv, mask):
        q = O.dense(xq, p[pre + "query/kernel"]).reshape(b, -1, H, d)
        k = qk(O.dense(xkv, p[pre + "key/kernel"])).reshape(b, -1, H, d)
        v = qv(O.dense(xkv, p[pre + "value/kernel"])).reshape(b, -1, H, d)
        x = O.dot_product_attention(q, k, v, O.mask_to_bias(mask, dtype)).reshape(b, -1, H * d)
        return O.dense(x, p[pre + "out/kernel"])

    for i in range(cfg.num_decoder_layers):
        pre = f"decoder/layers_{i}/"
        a = O.rms_norm(y, p[pre + "pre_self_attention_layer_norm/scale"])
        y = y + att(pre + "self_attention/", a, a, causal)
        c = O.rms_norm(y, p[pre + "pre_cross_attention_layer_norm/scale"])
        y = y + att(pre + "encoder_decoder_attention/", c, enc, None)
        m = O.rms_norm(y, p[pre + "pre_mlp_layer_norm/scale"])
        y = y + O.mlp_block(m, [p[pre + "mlp/wi_0/kernel"], p[pre + "mlp/wi_1/kernel"]], p[pre + "mlp/wo/kernel"], cfg.mlp_activations)
    y = O.rms_norm(y, p["decoder/decoder_norm/scale"])
    return O.dense(y, p["decoder/logits_dense/kernel"])


def main():
    length = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    full = len(sys.argv) > 2 and sys.argv[2] == "full"
    if full:
        cfg = O.T5Config()
        x = O.compute_spectrogram(np.stack([O.sine_mix(32768, 1234 + i) for i in range(2)]), np.float32)
        seed = 0
    else:   # the configuration of tests/test_gpu_parity.py::test_kv_cache_formats_sharp_attention
        cfg = O.T5Config(vocab_size=1536, num_encoder_layers=1, num_decoder_layers=3)
        x = np.random.default_rng(500).standard_normal((5, 64, 512)).astype(np.float32)
        seed = 33
    toks = np.random.default_rng(3).integers(3, 1500, size=(x.shape[0], length))
    toks[:, 0] = 0
    print(f"{'mt3' if full else '1+3-layer'} config, {x.shape[0]} sequences, {length} teacher-forced positions; max |logit error| / max |logit| vs float64")
    print(f"{'query scale':>12s} {'float32 arithmetic':>20s} {'F32 rows':>10s} {'P24 rows':>10s} {'F16 rows':>10s}")
    for qscale in (1.0, 4.0, 8.0, 16.0):
        params = O.init_params(cfg, seed=seed, norm_scale_jitter=0.05)
        for k in list(params):
            if k.startswith("decoder") and k.endswith("query/kernel"):
                params[k] = params[k] * np.float32(qscale)
        enc = O.encode(params, cfg, x, np.float64)
        ref = teacher_forced(params, cfg, enc, toks, lambda v: v, lambda v: v)
        sc = np.abs(ref).max()
        l32 = O.decode_teacher_forced(params, cfg, O.encode(params, cfg, x, np.float32), toks.astype(np.int32), np.float32)
        row = [np.abs(l32 - ref).max() / sc]
        for q in (q_f32, q_p24, q_f16):
            row.append(np.abs(teacher_forced(params, cfg, enc, toks, q, q) - ref).max() / sc)
        print(f"{qscale:12.0f} {row[0]:20.2e} {row[1]:10.2e} {row[2]:10.2e} {row[3]:10.2e}")


if __name__ == "__main__":
    main()
