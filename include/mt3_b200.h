/*
 * mt3_b200.h -- C ABI of the B200-native MT3 audio -> event-token hot path.
 *
 * The reference (magenta/mt3) is pure Python and has no FFI of its own; the seams
 * this library replaces are Python call signatures.  Each entry
 * point names the reference interface it stands in for (file:line under
 * /root/reference/mt3/).  INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer marked DEV is device memory
 *     owned by the caller (a torch tensor's data_ptr() is fine), HOST is host
 *     memory.  The library never allocates in the hot path: models and frontends
 *     allocate their tables once at *_create; all per-call scratch comes from the
 *     caller's workspace (mt3_workspace_bytes).
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered.
 *   - every function returns an int status (0 = MT3_OK); mt3_last_error() returns
 *     a thread-local description of the last failure.  These mirror the
 *     reference's fail-fast Python exceptions (rank asserts network.py:53,168,211,
 *     281; cache shape ValueError layers.py:266-270; bad token dtype layers.py:528).
 *   - handles are not thread-safe; one host thread per GPU/process.
 *   - inference only: dropout off, no RNG (deterministic=True paths).
 */
#ifndef MT3_B200_H_
#define MT3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MT3_ABI_VERSION 2

enum {
  MT3_OK = 0,
  MT3_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, bad flag          */
  MT3_ERR_SHAPE = -2,        /* shape the model/frontend was not created for        */
  MT3_ERR_UNSUPPORTED = -3,  /* valid request this build has no kernel for          */
  MT3_ERR_WORKSPACE = -4,    /* workspace too small / misaligned                    */
  MT3_ERR_CUDA = -5,         /* a CUDA runtime call failed (see mt3_last_error)     */
  MT3_ERR_STATE = -6         /* call order violated (e.g. decode before cross_kv)   */
};

int mt3_abi_version(void);
const char* mt3_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's
 * gpu_launches claim is read from here).  Launches replayed through a CUDA graph
 * are counted per replay. */
uint64_t mt3_kernel_launch_count(void);

/* ------------------------------------------------------------------------- */
/* Frontend: spectrograms.compute_spectrogram (spectrograms.py:64-73) ->      */
/* spectral_ops.compute_logmel (spectral_ops.py:76-88) -> compute_mel (:57-73) */
/* -> compute_mag / stft (:35-54) -> safe_log (:29-32).                        */
/* ------------------------------------------------------------------------- */

typedef struct mt3_frontend mt3_frontend;

typedef struct {
  int32_t sample_rate;   /* spectrograms.py:23  (16000)                       */
  int32_t hop_width;     /* spectrograms.py:24  (128)                         */
  int32_t fft_size;      /* spectrograms.py:28  (2048; the only parity point) */
  int32_t num_mel_bins;  /* spectrograms.py:25  (512)                         */
  float log_eps;         /* spectral_ops.py:29  (1e-5, replace-not-add)       */
} mt3_frontend_config;

/* mel_matrix HOST float32 [fft_size/2+1, num_mel_bins] row-major: the matrix the
 * reference obtains from tf.signal.linear_to_mel_weight_matrix at
 * spectral_ops.py:69-70.  It is converted to a banded per-mel-bin form here. */
int mt3_frontend_create(const mt3_frontend_config* cfg, const float* mel_matrix, mt3_frontend** out);
int mt3_frontend_destroy(mt3_frontend* fe);

/* Number of frames for n samples: ceil(n / hop) (tf.signal.frame pad_end=True,
 * spectral_ops.py:47). */
int mt3_frontend_num_frames(const mt3_frontend* fe, int64_t n_samples);

/* audio DEV f32 [S, n_samples] (row stride = audio_stride elements), each row an
 * independent segment (preprocessors.py:613-618 runs per example).
 * n_valid_frames DEV i32 [S] or NULL: rows >= n_valid_frames[s] are written as
 * 0.0 -- the feature converter's padding of a short last segment (models.py:96).
 * out DEV f32 [S, T, num_mel_bins], T = ceil(n_samples / hop). */
int mt3_logmel_f32(const mt3_frontend* fe, const float* audio, int64_t audio_stride, int32_t num_segments,
                   int32_t n_samples, const int32_t* n_valid_frames, float* out, void* stream);

/* ------------------------------------------------------------------------- */
/* Model: network.Transformer (network.py:265-409) with network.T5Config      */
/* (network.py:25-41) as bound in gin/model.gin:47-59.                         */
/* ------------------------------------------------------------------------- */

typedef struct mt3_model mt3_model;

typedef struct {
  int32_t vocab_size;          /* vocabularies.num_embeddings (vocabularies.py:280) */
  int32_t emb_dim;             /* 512  */
  int32_t num_heads;           /* 6    */
  int32_t head_dim;            /* 64   */
  int32_t num_encoder_layers;  /* 8    */
  int32_t num_decoder_layers;  /* 8    */
  int32_t mlp_dim;             /* 1024 */
  int32_t input_depth;         /* spectrograms.input_depth = num_mel_bins (512)     */
  int32_t max_batch;           /* largest B any call will use                        */
  int32_t max_input_length;    /* largest T (256 mt3, 512 ismir2021)                 */
  int32_t max_decode_length;   /* decoder length L (1024)                            */
  int32_t gemm_mode;           /* MT3_GEMM_* below                                   */
  int32_t kv_cache_format;     /* MT3_KV_* below (ABI version 2)                     */
} mt3_model_config;

enum {
  MT3_GEMM_FP32_SIMT = 0,   /* exact fp32 FMA on CUDA cores (debug / parity anchor)  */
  MT3_GEMM_TF32X3 = 1,      /* tcgen05 kind::tf32, 3-term split, fp32-faithful       */
  MT3_GEMM_TF32 = 2         /* tcgen05 kind::tf32 single pass (10-bit mantissa)      */
};

/* Storage format of the decoder's K/V rows -- the self-attention cache (layers.py:249-289) and the hoisted
 * cross-attention K/V.  All arithmetic (projections, scores, softmax, P.V) stays float32; only the stored rows are
 * rounded, and the rounding of a stored row is amplified by every later softmax, the more the sharper the attention
 * (error against the float64 oracle, diffuse / sharp attention):
 *   F32  the reference's values exactly                                   4e-6 / 2e-6    100 % of the bytes
 *   P24  float32 cut to 16 mantissa bits, a u16 + u8 plane inside a row   5e-6 / 3e-5     75 %   (default of InferenceModel)
 *   F16  IEEE half                                                        1.6e-4 / 1.2e-3  50 %   (bar: 5e-4)          */
enum {
  MT3_KV_F32 = 0,
  MT3_KV_F16 = 1,
  MT3_KV_P24 = 2
};

/* Number of float32 elements in the flat weight blob and the offset of a named
 * parameter inside it.  The blob is the Flax parameter tree in
 * the order documented in INTEGRATION.md; names are the tree paths, e.g.
 * "encoder/layers_0/attention/query/kernel".  Returns -1 for an unknown name. */
int64_t mt3_model_num_params(const mt3_model_config* cfg);
int64_t mt3_model_param_offset(const mt3_model_config* cfg, const char* name, int64_t* numel);

/* weights DEV f32 [mt3_model_num_params]; copied/re-laid-out into library-owned
 * device memory (fused QKV, folded norm scales, K-major tiles), so the caller's
 * blob may be freed afterwards. */
int mt3_model_create(const mt3_model_config* cfg, const float* weights, mt3_model** out, void* stream);
int mt3_model_destroy(mt3_model* m);

/* Bytes of caller-provided DEV workspace for calls with this (B, T). 256-byte
 * aligned base required. */
int64_t mt3_workspace_bytes(const mt3_model* m, int32_t batch, int32_t input_length);
int mt3_model_set_workspace(mt3_model* m, void* workspace, int64_t bytes, int32_t batch, int32_t input_length);

/* Transformer.encode(encoder_input_tokens, enable_dropout=False) (network.py:275-301).
 * x DEV f32 [B, T, input_depth] -> encoded DEV f32 [B, T, emb_dim]. */
int mt3_encode(mt3_model* m, const float* x, float* encoded, void* stream);

/* Projects `encoded` to the cross-attention K/V of every decoder layer once per
 * batch (the reference re-projects on every step: network.py:129-135 passes no
 * decode flag) and resets the decode state: cache_index = 0,
 * position_embedder_index = 0 (layers.py:255-260, :589-596). */
int mt3_cross_kv(mt3_model* m, const float* encoded, void* stream);

/* One Transformer.decode(decode=True) step (network.py:303-361): embeds tok_in
 * DEV i32 [B] at the current position, appends self-attention K/V to the cache,
 * and writes logits DEV f32 [B, vocab] (may be NULL) and the greedy next token
 * tok_out DEV i32 [B] (may be NULL).  Position advances by one. */
int mt3_decode_step(mt3_model* m, const int32_t* tok_in, float* logits, int32_t* tok_out, void* stream);

enum {
  MT3_GEN_STOP_AT_EOS = 1,   /* end the loop once every sequence emitted EOS (id 1)   */
  MT3_GEN_USE_GRAPH = 2,     /* replay one captured CUDA graph per step                */
  MT3_GEN_BEAM1 = 4          /* t5x decoding.beam_search bookkeeping at num_decodes = 1 (the reference's decode_fn,
                                models.py:127) instead of greedy-until-EOS: the live prefix continues with the best non-EOS
                                token, "prefix + EOS" hypotheses are ranked by log-prob / ((5 + length) / 6)^0.6, and the best
                                finished hypothesis is returned.  With MT3_GEN_STOP_AT_EOS the loop ends once every sequence's
                                finished score beats what its live prefix can still reach.                                */
};

/* predict_batch_with_aux stand-in (models.py:121-138 + t5x decode loop, greedy):
 * encode + cross_kv + up to num_steps greedy steps from BOS=0.  tokens_out DEV i32
 * [B, max_decode_length]: raw model ids, 0 after EOS and beyond num_steps.
 * x DEV f32 [B, T, input_depth].  steps_run HOST (may be NULL) receives the number
 * of steps executed.  Synchronises the stream before returning only when
 * MT3_GEN_STOP_AT_EOS is set (it has to poll the finished flag). */
int mt3_generate(mt3_model* m, const float* x, int32_t num_steps, int32_t flags, int32_t* tokens_out,
                 int32_t* steps_run, void* stream);

/* GenericTokenVocabulary._decode_tf (vocabularies.py:241-271): ids DEV i32 [B, L]
 * -> out DEV i32 [B, L] with id-3, EOS-and-after = -1, invalid = -2. */
int mt3_vocab_decode(const int32_t* ids, int32_t batch, int32_t length, int32_t num_regular_tokens, int32_t* out,
                     void* stream);

/* layers.dot_product_attention(query, key, value, bias=None) (layers.py:85-157) as a standalone op, float32, any head_dim:
 * softmax(q k^T + bias) v with no 1/sqrt(d) scaling.  q DEV f32 [batch, q_len, heads, head_dim]; k, v DEV f32
 * [batch, kv_len, heads, head_dim]; bias DEV f32 [batch, heads, q_len, kv_len] or NULL (the reference's combined mask /
 * relative-position bias, layers.py:143-146; broadcasting is the caller's job); out DEV f32 like q.  The model's own
 * attention kernels specialise this op for head_dim 64 and the two mask shapes MT3 uses (all-ones, causal-to-length). */
int mt3_dot_product_attention_f32(const float* q, const float* k, const float* v, const float* bias, int32_t batch,
                                  int32_t q_len, int32_t kv_len, int32_t num_heads, int32_t head_dim, float* out, void* stream);

/* Measurement hook for bench.py's roofline leg: launches one named hot kernel `iters` times at
 * the shapes of the bound workspace, cycling over the decoder/encoder layers so that successive
 * launches touch different memory.  kind: MT3_K_* below; `pos` = KV-cache length - 1 for the
 * self-attention kernel.  Results are scratch; call mt3_cross_kv again before decoding. */
enum {
  MT3_K_DEC_SELF_ATTN = 0,
  MT3_K_DEC_CROSS_ATTN = 1,
  MT3_K_DEC_QKV_GEMM = 2,
  MT3_K_ENC_QKV_GEMM = 3,
  MT3_K_ENC_ATTN = 4
};
int mt3_debug_launch(mt3_model* m, int32_t kind, int32_t pos, int32_t iters, void* stream);

/* Measurement hook: device timeline of ONE greedy decode step at cache position `pos` (>= 2), replayed as a
 * CUDA graph exactly like mt3_generate's step.  Every decode GEMM / attention node records into its slot of
 * `out` (host, [max_slots][16] u64): [0] earliest CTA start, [1] latest CTA end (ns, %globaltimer),
 * [2..6] SM-clock deltas of CTA (0,0) at its internal phase boundaries.  `names` receives the slot names,
 * newline separated.  Clobbers the decode state (position, tokens); call mt3_encode/mt3_cross_kv first. */
int mt3_debug_trace_step(mt3_model* m, int32_t pos, uint64_t* out, int32_t max_slots, char* names,
                         int32_t names_bytes, int32_t* n_slots, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MT3_B200_H_ */
