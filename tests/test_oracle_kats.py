"""Pins the oracle against the reference's own known-answer tests.

Each case restates a test from /root/reference/mt3/*_test.py (cited per test)
with numpy only: the expected side is derived exactly as the reference test
derives it (plain einsum / literal numbers), the actual side is the oracle.
"""
import numpy as np
import pytest

from oracle import mt3_oracle as O


@pytest.mark.parametrize("f", [20, 22])
def test_multihead_dot_product_attention(f):
    # layers_test.py:285-330
    b, q, h, d, k = 2, 3, 4, 5, 6
    np.random.seed(0)
    inputs_q = np.random.randn(b, q, f)
    inputs_kv = np.random.randn(b, k, f)
    query_kernel = np.random.randn(f, h, d)
    key_kernel = np.random.randn(f, h, d)
    value_kernel = np.random.randn(f, h, d)
    out_kernel = np.random.randn(h, d, f)
    params = {
        "a/query/kernel": query_kernel.reshape(f, -1),
        "a/key/kernel": key_kernel.reshape(f, -1),
        "a/value/kernel": value_kernel.reshape(f, -1),
        "a/out/kernel": out_kernel.reshape(-1, f),
    }
    for dtype, tol in ((np.float64, 1e-12), (np.float32, 1e-4)):
        p = {n: v.astype(dtype) for n, v in params.items()}
        y = O.mha(p, "a/", inputs_q.astype(dtype), inputs_kv.astype(dtype), h, d)
        query = np.einsum("bqf,fhd->bqhd", inputs_q, query_kernel)
        key = np.einsum("bkf,fhd->bkhd", inputs_kv, key_kernel)
        value = np.einsum("bkf,fhd->bkhd", inputs_kv, value_kernel)
        logits = np.einsum("bqhd,bkhd->bhqk", query, key)
        weights = O.softmax(logits, -1)
        combined = np.einsum("bhqk,bkhd->bqhd", weights, value)
        y_expected = np.einsum("bqhd,hdf->bqf", combined, out_kernel)
        np.testing.assert_allclose(y, y_expected, rtol=tol, atol=tol * 10)


def test_dot_product_attention_with_bias():
    # layers_test.py:375-387
    b, q, h, d, k = 2, 3, 4, 5, 6
    np.random.seed(0)
    query = np.random.randn(b, q, h, d)
    key = np.random.randn(b, k, h, d)
    value = np.random.randn(b, k, h, d)
    bias = np.random.randn(b, h, q, k)
    out = O.dot_product_attention(query, key, value, bias=bias)
    logits = np.einsum("bqhd,bkhd->bhqk", query, key)
    w = np.exp(logits + bias)
    w /= w.sum(-1, keepdims=True)
    expected = np.einsum("bhqk,bkhd->bqhd", w, value)
    np.testing.assert_allclose(out, expected, atol=1e-6)


def test_attention_caching_semantics():
    # layers_test.py:332-373: slot 0 written, index -> 1, layout [b, h, d, k].
    b, h, d, k = 2, 3, 4, 5
    f = h * d
    np.random.seed(1)
    inputs_q = np.random.randn(b, 1, f)
    inputs_kv = np.random.randn(b, 1, f)
    eye = np.eye(f)  # stands in for the reference's mock "projection = reshape"
    params = {"a/query/kernel": eye, "a/key/kernel": eye, "a/value/kernel": eye, "a/out/kernel": eye}
    cache = O.KVCache(np.zeros((b, h, d, k)), np.zeros((b, h, d, k)), 0)
    O.mha(params, "a/", inputs_q, inputs_kv, h, d, cache=cache)
    key = inputs_kv.reshape(b, -1, h, d)
    exp_k = np.zeros((b, h, d, k))
    exp_k[:, :, :, 0] = key[:, 0, :, :]
    np.testing.assert_allclose(cache.cached_key, exp_k)
    np.testing.assert_allclose(cache.cached_value, exp_k)
    assert cache.cache_index == 1
    # wrong query shape -> the reference's ValueError (layers.py:266-270)
    with pytest.raises(ValueError):
        O.mha(params, "a/", np.random.randn(b, 2, f), np.random.randn(b, 2, f), h, d, cache=cache)


def test_cached_decode_masks_future_slots():
    # layers.py:297-305: only slots <= cache_index are attended to.
    b, h, d, L = 1, 2, 4, 6
    f = h * d
    rng = np.random.default_rng(0)
    eye = np.eye(f)
    params = {"a/query/kernel": eye, "a/key/kernel": eye, "a/value/kernel": eye, "a/out/kernel": eye}
    cache = O.KVCache(np.zeros((b, h, d, L)), np.zeros((b, h, d, L)), 0)
    xs = rng.standard_normal((3, b, 1, f))
    outs = [O.mha(params, "a/", x, x, h, d, cache=cache) for x in xs]
    # full-sequence causal attention gives the same rows
    seq = np.concatenate(list(xs), axis=1)
    causal = np.tril(np.ones((3, 3)))[None, None]
    full = O.mha(params, "a/", seq, seq, h, d, mask=causal)
    np.testing.assert_allclose(np.concatenate(outs, axis=1), full, atol=1e-12)


def test_dense_general_ones():
    # layers_test.py:452-484 (3., 3., 4.)
    np.testing.assert_allclose(O.dense(np.ones((1, 3)), np.ones((3, 4))), np.full((1, 4), 3.0))
    np.testing.assert_allclose(O.dense(np.ones((1, 3)), np.ones((3, 4))).reshape(1, 2, 2), np.full((1, 2, 2), 3.0))
    np.testing.assert_allclose(O.dense(np.ones((1, 4)), np.ones((4, 3))), np.full((1, 3), 4.0))


def test_mlp_relu_golden():
    # layers_test.py:502-541 (the commented-out golden; numbers are the reference's).
    wi = np.array([[-0.8675811290740967, 0.08417510986328125, 0.022586345672607422, -0.9124102592468262],
                   [-0.19464373588562012, 0.49809837341308594, 0.7808468341827393, 0.9267289638519287]], np.float32)
    wo = np.array([[0.01154780387878418, 0.1397249698638916],
                   [0.974980354309082, 0.5903260707855225],
                   [-0.05997943878173828, 0.616570234298706],
                   [0.2934272289276123, 0.8181164264678955]], np.float32)
    inputs = np.array([[[1, 1], [1, 1], [1, 2]], [[2, 2], [3, 1], [2, 2]]], np.float32)
    expected = [[[0.5237172245979309, 0.8508185744285583],
                 [0.5237172245979309, 0.8508185744285583],
                 [1.2344461679458618, 2.3844780921936035]],
                [[1.0474344491958618, 1.7016371488571167],
                 [0.6809444427490234, 0.9663378596305847],
                 [1.0474344491958618, 1.7016371488571167]]]
    out = O.mlp_block(inputs, [wi], wo, ("relu",))
    np.testing.assert_allclose(out, expected, rtol=1e-6)


# The reference's own vectors, literally (vocabularies_test.py:47-83): GenericTokenVocabulary(32[, extra_ids=4]),
# i.e. 32 regular ids behind the 3 special ones.
VOCAB_TEST_VECTORS = [
    # (ids, decode_tf result) -- test_encode_decode :47-62
    ([4, 5, 6], [1, 2, 3]),
    # test_decode_invalid_ids :64-71
    ([0, 2, 3, 4, 34, 35], [-2, -2, 0, 1, 31, -2]),
    # test_decode_eos :73-83 (the TF form preserves the array length)
    ([0, 2, 3, 4, 1, 0, 1, 0], [-2, -2, 0, 1, -1, -1, -1, -1]),
]


def test_vocab_decode_reference_vectors():
    for ids, want in VOCAB_TEST_VECTORS:
        np.testing.assert_array_equal(O.vocab_decode(np.array(ids), 32), want)
    # the Python form of test_decode_eos truncates after the first EOS: [-2, -2, 0, 1, -1]
    dec = O.vocab_decode(np.array([0, 2, 3, 4, 1, 0, 1, 0]), 32)
    np.testing.assert_array_equal(dec[:int(np.argmax(dec == -1)) + 1], [-2, -2, 0, 1, -1])
    from mt3_b200 import vocabularies
    vocab = vocabularies.GenericTokenVocabulary(32, extra_ids=4)
    assert list(vocab.encode([1, 2, 3])) == [4, 5, 6]
    assert list(vocab.decode([0, 2, 3, 4, 34, 35])) == [-2, -2, 0, 1, 31, -2]
    assert list(vocabularies.GenericTokenVocabulary(32).decode([0, 2, 3, 4, 1, 0, 1, 0])) == [-2, -2, 0, 1, -1]


def test_vocab_decode_contract():
    # further cases in the spirit of vocabularies_test.py:47-83 (10 regular ids)
    n = 10
    np.testing.assert_array_equal(O.vocab_decode(np.array([3, 4, 5, 12]), n), [0, 1, 2, 9])
    # EOS is sticky: it and everything after -> -1
    np.testing.assert_array_equal(O.vocab_decode(np.array([3, 4, 1, 5, 6]), n), [0, 1, -1, -1, -1])
    # PAD/UNK and ids past the regular range (extra ids) -> -2
    np.testing.assert_array_equal(O.vocab_decode(np.array([0, 2, 13, 16, 3]), n), [-2, -2, -2, -2, 0])
    # batched
    got = O.vocab_decode(np.array([[3, 1, 3], [0, 3, 1]]), n)
    np.testing.assert_array_equal(got, [[0, -1, -1], [-2, 0, -1]])
    np.testing.assert_array_equal(O.trim_eos(np.array([5, 6, -1, -1])), [5, 6])
    np.testing.assert_array_equal(O.trim_eos(np.array([5, 6])), [5, 6])


def test_vocab_sizes():
    # mt3 1388 classes -> 1536 ; ismir2021 1514 -> 1664
    assert O.codec_num_classes(1) == 1388
    assert O.codec_num_classes(127) == 1514
    assert O.num_embeddings(1388) == 1536
    assert O.num_embeddings(1514) == 1664


def test_sinusoid_table_shape_and_values():
    pe = O.sinusoidal_table(2048, 512)
    assert pe.shape == (2048, 512) and pe.dtype == np.float32
    np.testing.assert_allclose(pe[0, :256], 0.0)
    np.testing.assert_allclose(pe[0, 256:], 1.0)
    np.testing.assert_allclose(pe[3, 0], np.sin(3.0), rtol=1e-6)
    np.testing.assert_allclose(pe[3, 255], np.sin(3.0 * 1e-4), rtol=1e-5)


def test_step_decode_equals_teacher_forced():
    cfg = O.T5Config(vocab_size=128, emb_dim=32, num_heads=2, num_encoder_layers=2, num_decoder_layers=2,
                     head_dim=8, mlp_dim=48, input_depth=16)
    params = O.init_params(cfg, seed=3, norm_scale_jitter=0.1)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 10, 16))
    enc = O.encode(params, cfg, x, np.float64)
    toks, logits = O.greedy_decode(params, cfg, enc, 12, np.float64, stop_at_eos=False, return_logits=True)
    dec_in = np.concatenate([np.zeros((2, 1), np.int64), toks[:, :-1]], axis=1)
    tf_logits = O.decode_teacher_forced(params, cfg, enc, dec_in, np.float64)
    np.testing.assert_allclose(logits, tf_logits, atol=1e-10)
    # the reference's per-step cross-K/V recomputation gives the same tokens
    toks2 = O.greedy_decode(params, cfg, enc, 12, np.float64, stop_at_eos=False, hoist_cross_kv=False)
    np.testing.assert_array_equal(toks, toks2)


def test_logmel_shapes_and_safe_log():
    x = O.sine_mix(32768, 7)
    lm = O.compute_spectrogram(x)
    assert lm.shape == (256, 512) and lm.dtype == np.float32
    z = O.compute_spectrogram(np.zeros(1000, np.float32))
    assert z.shape == (8, 512)
    np.testing.assert_allclose(z, np.log(np.float32(1e-5)))
    w = O.linear_to_mel_weight_matrix()
    assert w.shape == (1025, 512)
    assert (w[0] == 0).all()
    nnz_rows = np.nonzero(w.any(axis=1))[0]
    assert nnz_rows.max() <= 973
    assert (np.count_nonzero(w, axis=1) <= 2).all()


def test_audio_to_frames_always_pads():
    frames, times = O.audio_to_frames(np.ones(256, np.float32))
    assert frames.shape == (3, 128)        # 256 -> +128 pad -> 3 frames (notebook :321-322)
    assert (frames[2] == 0).all()
    frames, times = O.audio_to_frames(np.ones(32000, np.float32))
    assert frames.shape == (251, 128)
    np.testing.assert_allclose(times[:3], [0.0, 0.008, 0.016])


# ------------------------------------------------------------------------------------------------
# Mask helpers of the layer API (mt3_b200/layers.py; numpy on the host): the literal vectors of layers_test.py:117-283
# ------------------------------------------------------------------------------------------------
def _L():
    from mt3_b200 import layers
    return layers


def test_make_attention_mask_reference_vectors():
    L = _L()
    toks = np.array([[7, 0, 0], [8, 5, 0]])
    m = L.make_attention_mask(toks > 0, toks > 0, dtype=np.int32)                            # layers_test.py:117-125
    assert m.shape == (2, 1, 3, 3)
    np.testing.assert_array_equal(m[0, 0], [[1, 0, 0], [0, 0, 0], [0, 0, 0]])
    np.testing.assert_array_equal(m[1, 0], [[1, 1, 0], [1, 1, 0], [0, 0, 0]])
    seg = np.array([[1, 1, 2, 2, 2, 0], [1, 1, 1, 2, 0, 0]])
    m = L.make_attention_mask(seg, seg, pairwise_fn=np.equal, dtype=np.int32)                # :127-141 (padding is not special)
    assert m.shape == (2, 1, 6, 6)
    np.testing.assert_array_equal(m[0, 0], [[1, 1, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0], [0, 0, 1, 1, 1, 0], [0, 0, 1, 1, 1, 0],
                                            [0, 0, 1, 1, 1, 0], [0, 0, 0, 0, 0, 1]])
    np.testing.assert_array_equal(m[1, 0], [[1, 1, 1, 0, 0, 0], [1, 1, 1, 0, 0, 0], [1, 1, 1, 0, 0, 0], [0, 0, 0, 1, 0, 0],
                                            [0, 0, 0, 0, 1, 1], [0, 0, 0, 0, 1, 1]])


def test_make_causal_mask_reference_vectors():
    L = _L()
    tri = np.array([[1., 0., 0.], [1., 1., 0.], [1., 1., 1.]], np.float32)
    y = L.make_causal_mask(np.array([[7, 0, 0], [8, 5, 0]]))                                # :143-152
    assert y.shape == (2, 1, 3, 3) and y.dtype == np.float32
    np.testing.assert_array_equal(y[0, 0], tri)
    np.testing.assert_array_equal(y[1, 0], tri)
    assert L.make_causal_mask(np.ones((3, 3, 5)), extra_batch_dims=2).shape == (1, 1, 3, 3, 1, 5, 5)   # :154-157
    np.testing.assert_array_equal(L.make_causal_mask(np.ones((1, 3))), tri[None, None])     # :159-165


def test_combine_masks_and_biases_reference_vectors():
    L = _L()
    f = lambda *v: np.array(v, np.float32)
    np.testing.assert_array_equal(L.combine_masks(f(0, 1, 0, 1), None, f(1, 1, 1, 1), f(1, 1, 1, 0)), f(0, 1, 0, 0))   # :167-174
    np.testing.assert_array_equal(L.combine_biases(f(0, 1, 0, 1), None, f(0, 1, 1, 1), f(0, 1, 1, 0)), f(0, 3, 2, 2))  # :176-183
    assert L.combine_masks(None, None) is None and L.combine_biases() is None
    with pytest.raises(AssertionError):
        L.combine_masks(np.ones((2, 2)), np.ones((2,)))


def test_make_decoder_mask_reference_vectors():
    L = _L()
    m = L.make_decoder_mask(np.array([6, 7, 3, 0]), np.float32)                             # lm, unpacked :185-191
    np.testing.assert_array_equal(m, [[[1, 0, 0, 0], [1, 1, 0, 0], [1, 1, 1, 0], [0, 0, 0, 0]]])
    m = L.make_decoder_mask(np.array([[6, 7, 3, 4, 5, 0]]), np.float32, decoder_segment_ids=np.array([[1, 1, 1, 2, 2, 0]]))   # :193-203
    np.testing.assert_array_equal(m, [[[[1, 0, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0], [1, 1, 1, 0, 0, 0], [0, 0, 0, 1, 0, 0],
                                        [0, 0, 0, 1, 1, 0], [0, 0, 0, 0, 0, 0]]]])
    m = L.make_decoder_mask(np.array([[5, 6, 7, 3, 4, 0]]), np.float32,
                            decoder_causal_attention=np.array([[1, 1, 1, 0, 0, 0]]))        # prefix lm :205-216
    np.testing.assert_array_equal(m, [[[[1, 1, 1, 0, 0, 0], [1, 1, 1, 0, 0, 0], [1, 1, 1, 0, 0, 0], [1, 1, 1, 1, 0, 0],
                                        [1, 1, 1, 1, 1, 0], [0, 0, 0, 0, 0, 0]]]])
    m = L.make_decoder_mask(np.array([[5, 6, 7, 8, 3, 4, 0]]), np.float32,
                            decoder_causal_attention=np.array([[1, 1, 0, 1, 1, 0, 0]]),
                            decoder_segment_ids=np.array([[1, 1, 1, 2, 2, 2, 0]]))          # prefix lm, packed :218-231
    np.testing.assert_array_equal(m, [[[[1, 1, 0, 0, 0, 0, 0], [1, 1, 0, 0, 0, 0, 0], [1, 1, 1, 0, 0, 0, 0], [0, 0, 0, 1, 1, 0, 0],
                                        [0, 0, 0, 1, 1, 0, 0], [0, 0, 0, 1, 1, 1, 0], [0, 0, 0, 0, 0, 0, 0]]]])
    m = L.make_decoder_mask(np.array([[6, 7, 3, 0], [4, 5, 0, 0]]), np.float32,
                            decoder_causal_attention=np.array([[1, 1, 0, 0], [1, 0, 0, 0]]))   # :233-246
    assert m.shape == (2, 1, 4, 4)
    np.testing.assert_array_equal(m[0, 0], [[1, 1, 0, 0], [1, 1, 0, 0], [1, 1, 1, 0], [0, 0, 0, 0]])
    np.testing.assert_array_equal(m[1, 0], [[1, 0, 0, 0], [1, 1, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]])
    m = L.make_decoder_mask(np.array([[6, 7, 3, 4, 8, 9, 0]]), np.float32,
                            decoder_causal_attention=np.array([[1, 1, 0, 0, 1, 1, 0]]))     # composite :248-261
    assert m.shape == (1, 1, 7, 7)
    np.testing.assert_array_equal(m[0, 0], [[1, 1, 0, 0, 1, 1, 0], [1, 1, 0, 0, 1, 1, 0], [1, 1, 1, 0, 0, 0, 0], [1, 1, 1, 1, 0, 0, 0],
                                            [1, 1, 1, 1, 1, 1, 0], [1, 1, 1, 1, 1, 1, 0], [0, 0, 0, 0, 0, 0, 0]])
    m = L.make_decoder_mask(np.array([[6, 7, 3, 4, 8, 9, 2, 3, 4]]), np.float32,
                            decoder_causal_attention=np.array([[1, 1, 0, 0, 1, 1, 1, 1, 0]]),
                            decoder_segment_ids=np.array([[1, 1, 1, 1, 1, 1, 2, 2, 2]]))    # composite, packed :263-283
    assert m.shape == (1, 1, 9, 9)
    np.testing.assert_array_equal(m[0, 0], [[1, 1, 0, 0, 1, 1, 0, 0, 0], [1, 1, 0, 0, 1, 1, 0, 0, 0], [1, 1, 1, 0, 0, 0, 0, 0, 0],
                                            [1, 1, 1, 1, 0, 0, 0, 0, 0], [1, 1, 1, 1, 1, 1, 0, 0, 0], [1, 1, 1, 1, 1, 1, 0, 0, 0],
                                            [0, 0, 0, 0, 0, 0, 1, 1, 0], [0, 0, 0, 0, 0, 0, 1, 1, 0], [0, 0, 0, 0, 0, 0, 1, 1, 1]])


def test_mask_to_bias_and_masked_attention_in_the_oracle():
    """mask -> bias the way MultiHeadDotProductAttention does (layers.py:316-328), and the effect through the oracle's
    dot_product_attention: a key masked out for a query gets exactly zero weight (exp(-1e10 - max) underflows to 0)."""
    L = _L()
    mask = L.make_decoder_mask(np.array([[6, 7, 3, 0]]), np.float32)
    bias = L.mask_to_bias(mask)
    assert bias.dtype == np.float32 and set(np.unique(bias)) == {np.float32(0.0), np.float32(-1e10)}
    np.testing.assert_array_equal(bias, O.mask_to_bias(mask, np.float32))
    rng = np.random.default_rng(0)
    q, k, v = (rng.standard_normal((1, 4, 2, 8)) for _ in range(3))
    full = O.dot_product_attention(q, k, v, bias=bias.astype(np.float64))
    for i in range(3):                      # query i sees keys 0..i only: same as attention over the truncated key set
        part = O.dot_product_attention(q[:, i:i + 1], k[:, :i + 1], v[:, :i + 1])
        np.testing.assert_allclose(full[:, i:i + 1], part, rtol=0, atol=1e-12)


def test_baseline_config1_ismir2021_single_clip_cpu_plumbing():
    """BASELINE.json configs[0]: the ismir2021 (piano) configuration on ONE 2 s / 16 kHz synthetic clip through the CPU
    restatement only -- the plumbing case: 32 000 samples -> +128 pad -> 251 frames -> one segment of 512 ->
    log-mel [251, 512] -> 0.0 rows up to 512 -> encoder -> greedy decode -> vocabulary decode.  V = 1664 (1514 classes + 3 +
    100 extra ids, rounded up to a multiple of 128), velocity vocabulary of 127 bins, NoteEncodingSpec without
    ties.  The transformer runs with one layer on each side to stay in CPU-test time; shapes and the token contract do
    not depend on depth."""
    from mt3_b200 import note_decoding, vocabularies as V
    codec = V.build_codec(V.VocabularyConfig(num_velocity_bins=127))
    vocab = V.vocabulary_from_codec(codec)
    assert codec.num_classes == 1514 == O.codec_num_classes(127) and V.num_embeddings(vocab) == 1664 == O.num_embeddings(1514)
    audio = O.sine_mix(32000, seed=1234)
    frames, times = O.audio_to_frames(audio)
    assert frames.shape == (251, 128)
    segs = O.split_to_segments(frames, times, 512)
    assert len(segs) == 1 and segs[0][0].shape == (251, 128) and segs[0][1][0] == 0.0
    spec = O.compute_spectrogram(segs[0][0].reshape(-1))
    assert spec.shape == (251, 512) and np.isfinite(spec).all()
    x = O.pad_inputs(spec, 512)
    assert x.shape == (512, 512) and (x[251:] == 0.0).all() and (x[:251] != 0.0).any()
    cfg = O.T5Config(vocab_size=1664, num_encoder_layers=1, num_decoder_layers=1)
    params = O.init_params(cfg, seed=0)
    enc = O.encode(params, cfg, x[None], np.float32)
    assert enc.shape == (1, 512, 512)
    toks = O.greedy_decode(params, cfg, enc, 6, np.float32, stop_at_eos=True)
    assert toks.shape[0] == 1 and toks.dtype.kind == 'i' and (toks >= 0).all() and (toks < 1664).all()
    decoded = O.trim_eos(O.vocab_decode(toks, codec.num_classes)[0])
    assert ((decoded >= -2) & (decoded < codec.num_classes)).all() and (decoded != -1).all()
    np.testing.assert_array_equal(O.vocab_decode(toks, codec.num_classes), vocab.decode_tf(np.asarray(toks)))
    # the stitch accepts whatever came out (random weights: mostly invalid / out-of-context events, counted not raised)
    res = note_decoding.event_predictions_to_ns([{'est_tokens': decoded, 'start_time': 0.0, 'raw_inputs': []}], codec, 'NoteEncodingSpec')
    assert res['est_invalid_events'] + res['est_dropped_events'] <= len(decoded)
