"""Independent cross-checks of the oracle (CPU).  The reference's arithmetic lives in tf.signal / flax / t5x, none installable
here, so the oracle restates it (oracle/mt3_oracle.py: "parity unpinned" for the frontend and the full model).  These tests
pin the restatement against INDEPENDENT third-party implementations of the same published algorithms that ARE in this image:

  * tf.signal.linear_to_mel_weight_matrix (HTK mel scale, triangles built in mel space)  <->  transformers.audio_utils.mel_filter_bank
  * tf.signal.stft(pad_end=True) with the periodic Hann window                           <->  torch.stft(center=False)
  * flax.linen.gelu(approximate=True)                                                    <->  torch gelu(approximate='tanh')
  * layers.dot_product_attention (no 1/sqrt(d) scaling, additive bias)                   <->  torch scaled_dot_product_attention(scale=1)
  * the T5.1.1 encoder / decoder layer (pre-RMSNorm, un-scaled attention, gated-GELU MLP; network.py:44-155)
                                                                                         <->  transformers' T5Block (T5 v1.1)
They are not the reference, but a shared misreading would have to be shared with those libraries too."""
import numpy as np
import pytest
import torch

from oracle import mt3_oracle as O


def test_mel_matrix_vs_transformers_htk_filterbank():
    au = pytest.importorskip("transformers.audio_utils")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                   # "at least one mel filter has all zero values": true here too (2 of 512)
        ref = au.mel_filter_bank(num_frequency_bins=1025, num_mel_filters=512, min_frequency=20.0, max_frequency=7600.0,
                                 sampling_rate=16000, norm=None, mel_scale="htk", triangularize_in_mel_space=True)
    m64 = O.linear_to_mel_weight_matrix(dtype=np.float64)
    assert ref.shape == m64.shape == (1025, 512)
    assert np.abs(ref - m64).max() < 1e-12
    np.testing.assert_array_equal(ref != 0, m64 != 0)
    assert int((m64 != 0).sum()) == 1934
    # the float32 variant follows TF's float32 evaluation order: same support, values within float32 rounding of the exact ones
    m32 = O.linear_to_mel_weight_matrix(dtype=np.float32)
    np.testing.assert_array_equal(m32 != 0, m64 != 0)
    assert np.abs(m32 - m64).max() < 1e-4


@pytest.mark.parametrize("fft,n", [(2048, 32768), (2048, 5000), (1024, 9000), (4096, 9000)])
def test_stft_magnitude_vs_torch_stft(fft, n):
    x = O.sine_mix(n, 7).astype(np.float64) + 0.01 * np.random.default_rng(0).standard_normal(n)
    mag = O.stft_mag(x[None], frame_size=fft, hop=128, dtype=np.float64)[0]
    frames = -(-n // 128)                                  # pad_end=True: ceil(n / hop) frames
    assert mag.shape == (frames, fft // 2 + 1)
    xp = np.concatenate([x, np.zeros(fft)])                # zero padding past the end, like tf.signal.frame(pad_end=True)
    win = torch.hann_window(fft, periodic=True, dtype=torch.float64)
    ref = torch.stft(torch.from_numpy(xp), n_fft=fft, hop_length=128, win_length=fft, window=win, center=False,
                     return_complex=True).abs().T.numpy()[:frames]
    assert np.abs(ref - mag).max() <= 1e-11 * max(1.0, np.abs(mag).max())


def test_logmel_pipeline_vs_torch_and_transformers():
    au = pytest.importorskip("transformers.audio_utils")
    import warnings
    x = O.sine_mix(32768, 11).astype(np.float64)
    got = O.compute_logmel(x[None], dtype=np.float64)[0]
    xp = np.concatenate([x, np.zeros(2048)])
    win = torch.hann_window(2048, periodic=True, dtype=torch.float64)
    mag = torch.stft(torch.from_numpy(xp), 2048, 128, 2048, win, center=False, return_complex=True).abs().T.numpy()[:256]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mel = au.mel_filter_bank(1025, 512, 20.0, 7600.0, 16000, norm=None, mel_scale="htk", triangularize_in_mel_space=True)
    m = mag @ mel
    # the oracle (like TF) builds the filterbank in float32 and only then casts: its coefficients differ from the exact ones by
    # float32 rounding (< 1e-4, checked above), so compare in the linear mel domain at that tolerance ...
    got_lin = np.exp(got)
    bound = 1e-4 * (mag @ (mel != 0).astype(np.float64)) + 1e-9 * m.max()      # |sum_b mag_b dW_b| <= max|dW| * sum of the band's magnitudes
    assert (np.abs(got_lin - m) <= bound)[m > 1e-5].all()
    # ... and exactly once the same float32-built matrix is used: frame -> window -> rfft -> |.| -> matmul -> replace-not-add log
    m32 = mag @ O.linear_to_mel_weight_matrix(dtype=np.float32).astype(np.float64)
    want = np.where(m32 <= 0.0, 1e-5, m32)                 # safe_log's argument (spectral_ops.py:29-32)
    np.testing.assert_allclose(np.exp(got), want, rtol=1e-6, atol=1e-12)   # (two FFT implementations: 1e-14 apart, amplified in quiet bins)


def test_gelu_and_attention_vs_torch():
    x = np.linspace(-6, 6, 1001)
    np.testing.assert_allclose(O.gelu_tanh(x), torch.nn.functional.gelu(torch.from_numpy(x), approximate='tanh').numpy(), atol=1e-14)
    rng = np.random.default_rng(0)
    q, k, v = (rng.standard_normal((2, 7, 3, 16)) for _ in range(3))      # [b, len, heads, d]
    bias = rng.standard_normal((2, 3, 7, 7))
    got = O.dot_product_attention(q, k, v, bias)
    t = lambda a: torch.from_numpy(a).permute(0, 2, 1, 3)                  # -> [b, heads, len, d]
    ref = torch.nn.functional.scaled_dot_product_attention(t(q), t(k), t(v), attn_mask=torch.from_numpy(bias), scale=1.0)
    np.testing.assert_allclose(got, ref.permute(0, 2, 1, 3).numpy(), atol=1e-12)


def _hf_block(ocfg, params, prefix, is_decoder):
    t5 = pytest.importorskip("transformers.models.t5.modeling_t5")
    from transformers import T5Config
    cfg = T5Config(vocab_size=ocfg.vocab_size, d_model=ocfg.emb_dim, d_kv=ocfg.head_dim, num_heads=ocfg.num_heads, d_ff=ocfg.mlp_dim,
                   feed_forward_proj="gated-gelu", dropout_rate=0.0, layer_norm_epsilon=1e-6, is_decoder=is_decoder,
                   num_layers=1, num_decoder_layers=1)
    cfg._attn_implementation = "eager"
    try:
        blk = t5.T5Block(cfg, has_relative_attention_bias=False, layer_idx=0)
    except TypeError:
        blk = t5.T5Block(cfg, has_relative_attention_bias=False)
    blk = blk.double().eval()
    W = lambda name: torch.from_numpy(params[prefix + name].astype(np.float64).T.copy())      # Flax [in, out] -> Linear [out, in]
    S = lambda name: torch.from_numpy(params[prefix + name].astype(np.float64).reshape(-1))
    with torch.no_grad():
        att = blk.layer[0]
        sa = 'self_attention/' if is_decoder else 'attention/'
        att.SelfAttention.q.weight.copy_(W(sa + 'query/kernel')); att.SelfAttention.k.weight.copy_(W(sa + 'key/kernel'))
        att.SelfAttention.v.weight.copy_(W(sa + 'value/kernel')); att.SelfAttention.o.weight.copy_(W(sa + 'out/kernel'))
        att.layer_norm.weight.copy_(S('pre_self_attention_layer_norm/scale' if is_decoder else 'pre_attention_layer_norm/scale'))
        if is_decoder:
            ca = blk.layer[1]
            ca.EncDecAttention.q.weight.copy_(W('encoder_decoder_attention/query/kernel'))
            ca.EncDecAttention.k.weight.copy_(W('encoder_decoder_attention/key/kernel'))
            ca.EncDecAttention.v.weight.copy_(W('encoder_decoder_attention/value/kernel'))
            ca.EncDecAttention.o.weight.copy_(W('encoder_decoder_attention/out/kernel'))
            ca.layer_norm.weight.copy_(S('pre_cross_attention_layer_norm/scale'))
        ff = blk.layer[-1]
        ff.DenseReluDense.wi_0.weight.copy_(W('mlp/wi_0/kernel')); ff.DenseReluDense.wi_1.weight.copy_(W('mlp/wi_1/kernel'))
        ff.DenseReluDense.wo.weight.copy_(W('mlp/wo/kernel'))
        ff.layer_norm.weight.copy_(S('pre_mlp_layer_norm/scale'))
    return blk


def test_encoder_layer_vs_transformers_t5_block():
    """network.py:44-85 (EncoderLayer) restated in the oracle == transformers' T5 v1.1 block with the same weights."""
    ocfg = O.T5Config(vocab_size=1536, num_encoder_layers=1, num_decoder_layers=1)
    params = O.init_params(ocfg, seed=5, norm_scale_jitter=0.1)
    x = O.compute_spectrogram(O.sine_mix(24 * 128, 3)[None], np.float32)
    _, layers = O.encode(params, ocfg, x, np.float64, return_layers=True)
    blk = _hf_block(ocfg, params, 'encoder/layers_0/', is_decoder=False)
    with torch.no_grad():
        out = blk(torch.from_numpy(layers[0]))[0].numpy()
    # transformers evaluates the RMSNorm statistic in float32 even for float64 modules: agreement to float32 rounding
    np.testing.assert_allclose(out, layers[1], rtol=0, atol=1e-6 * np.abs(layers[1]).max())


def test_decoder_layer_vs_transformers_t5_block():
    """network.py:88-155 (DecoderLayer: causal self-attention, cross-attention over `encoded`, gated-GELU MLP) restated with
    the oracle's ops == transformers' T5 v1.1 decoder block with the same weights."""
    ocfg = O.T5Config(vocab_size=1536, num_encoder_layers=1, num_decoder_layers=1)
    params = O.init_params(ocfg, seed=6, norm_scale_jitter=0.1)
    p = O._cast(params, np.float64)
    rng = np.random.default_rng(1)
    b, L, T = 2, 9, 13
    y = rng.standard_normal((b, L, ocfg.emb_dim))
    enc = rng.standard_normal((b, T, ocfg.emb_dim))
    pre = 'decoder/layers_0/'
    causal = np.tril(np.ones((L, L)))[None, None]
    h = y + O.mha(p, pre + 'self_attention/', O.rms_norm(y, p[pre + 'pre_self_attention_layer_norm/scale']),
                  O.rms_norm(y, p[pre + 'pre_self_attention_layer_norm/scale']), ocfg.num_heads, ocfg.head_dim, mask=causal)
    h = h + O.mha(p, pre + 'encoder_decoder_attention/', O.rms_norm(h, p[pre + 'pre_cross_attention_layer_norm/scale']), enc,
                  ocfg.num_heads, ocfg.head_dim)
    m = O.rms_norm(h, p[pre + 'pre_mlp_layer_norm/scale'])
    want = h + O.mlp_block(m, [p[pre + 'mlp/wi_0/kernel'], p[pre + 'mlp/wi_1/kernel']], p[pre + 'mlp/wo/kernel'], ocfg.mlp_activations)
    blk = _hf_block(ocfg, params, pre, is_decoder=True)
    neg = torch.finfo(torch.float64).min
    mask = torch.from_numpy(np.where(causal > 0, 0.0, 1.0)) * neg        # additive causal mask [1, 1, L, L]
    with torch.no_grad():
        out = blk(torch.from_numpy(y), attention_mask=mask, encoder_hidden_states=torch.from_numpy(enc))[0].numpy()
    np.testing.assert_allclose(out, want, rtol=0, atol=1e-6 * np.abs(want).max())
