// Non-GEMM kernels of the MT3 encoder/decoder blocks (sm_100a): RMSNorm statistics,
// attention (encoder, and the Tq = 1 decode form over the KV cache), token embedding,
// greedy argmax with EOS bookkeeping, vocabulary decode, weight re-layout.
#pragma once

#include "common.cuh"

namespace mt3 {

// ---------------------------------------------------------------------------------
// RMSNorm (layers.py:604-621): rstd[m] = 1/sqrt(mean(x[m,:]^2) + eps).  One warp per row.
// ---------------------------------------------------------------------------------
// x_lo != null: the row is stored as a tf32 hi/lo pair (x = x_hi + x_lo exactly).
__global__ void row_rstd_kernel(const float* __restrict__ x, const float* __restrict__ x_lo, int ld, int M, int D,
                                float eps, float* __restrict__ rstd) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float4* p = reinterpret_cast<const float4*>(x + (long long)row * ld);
  const float4* pl = x_lo ? reinterpret_cast<const float4*>(x_lo + (long long)row * ld) : nullptr;
  float s = 0.f;
  for (int i = lane; i < D / 4; i += 32) {
    float4 v = p[i];
    if (pl) {
      const float4 w = pl[i];
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  s = warp_sum(s);
  if (lane == 0) rstd[row] = 1.0f / sqrtf(s / (float)D + eps);
}

// y = x * rstd * g  (materialised only where the reference's API returns the normed
// tensor itself: `encoded`, network.py:192).
__global__ void rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ x_lo, int ldx, int M, int D,
                               float eps, const float* __restrict__ g, float* __restrict__ y, int ldy) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float4* p = reinterpret_cast<const float4*>(x + (long long)row * ldx);
  const float4* pl = x_lo ? reinterpret_cast<const float4*>(x_lo + (long long)row * ldx) : nullptr;
  auto ld4 = [&](int i) {
    float4 v = p[i];
    if (pl) {
      const float4 w = pl[i];
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    return v;
  };
  float s = 0.f;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 v = ld4(i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  s = warp_sum(s);
  const float r = 1.0f / sqrtf(s / (float)D + eps);
  float4* q = reinterpret_cast<float4*>(y + (long long)row * ldy);
  const float4* gg = reinterpret_cast<const float4*>(g);
  for (int i = lane; i < D / 4; i += 32) {
    const float4 v = ld4(i);
    const float4 w = __ldg(gg + i);
    q[i] = make_float4(v.x * r * w.x, v.y * r * w.y, v.z * r * w.z, v.w * r * w.w);
  }
}

// ---------------------------------------------------------------------------------
// Encoder self-attention, exact fp32 (layers.py:85-157): softmax(q k^T) v, no 1/sqrt(d)
// scaling (layers.py:230-234), mask all ones (network.py:283-289).
// qkv [B*T, 3*H*64] rows = (b,t), columns = [q | k | v] each (h, d).  out [B*T, H*64].
// One CTA = 32 query rows of one (b, h); scores for all T keys live in shared memory.
// ---------------------------------------------------------------------------------
constexpr int kHD = 64;   // head_dim the kernels are specialised for (gin/model.gin:54)

__global__ void __launch_bounds__(256)
enc_attention_kernel(const float* __restrict__ qkv, int ld, int T, int H, float* __restrict__ out,
                     float* __restrict__ out_lo, int ldo) {
  extern __shared__ __align__(16) float sm[];
  constexpr int QT = 32, KT = 64;
  const int SP = T + 4;
  float* sQ = sm;                 // [QT][64]
  float* sS = sQ + QT * kHD;      // [QT][T+4]
  float* sT = sS + QT * SP;       // [64][KT+4]  K tile transposed ([d][k]) / V tile [k][d+4]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int tid = threadIdx.x;
  const int qoff = h * kHD, koff = H * kHD + h * kHD, voff = 2 * H * kHD + h * kHD;
  const float* base = qkv + (long long)b * T * ld;

  for (int i = tid; i < QT * kHD / 4; i += 256) {
    const int r = i / (kHD / 4), c4 = i % (kHD / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < T) v = *reinterpret_cast<const float4*>(base + (long long)(q0 + r) * ld + qoff + c4 * 4);
    *reinterpret_cast<float4*>(sQ + r * kHD + c4 * 4) = v;
  }
  // S = Q K^T : thread -> 2 query rows x 4 keys
  const int tq = tid / 16, tk = tid % 16;     // tq in [0,16): rows tq, tq+16 ; tk: keys tk*4..+3
  for (int k0 = 0; k0 < T; k0 += KT) {
    __syncthreads();
    for (int i = tid; i < KT * kHD / 4; i += 256) {
      const int r = i / (kHD / 4), c4 = i % (kHD / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + r < T) v = *reinterpret_cast<const float4*>(base + (long long)(k0 + r) * ld + koff + c4 * 4);
      sT[(c4 * 4 + 0) * (KT + 4) + r] = v.x;
      sT[(c4 * 4 + 1) * (KT + 4) + r] = v.y;
      sT[(c4 * 4 + 2) * (KT + 4) + r] = v.z;
      sT[(c4 * 4 + 3) * (KT + 4) + r] = v.w;
    }
    __syncthreads();
    float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int d = 0; d < kHD; ++d) {
      const float q_0 = sQ[tq * kHD + d], q_1 = sQ[(tq + 16) * kHD + d];
      const float4 kv = *reinterpret_cast<const float4*>(sT + d * (KT + 4) + tk * 4);
      a0[0] = fmaf(q_0, kv.x, a0[0]); a0[1] = fmaf(q_0, kv.y, a0[1]);
      a0[2] = fmaf(q_0, kv.z, a0[2]); a0[3] = fmaf(q_0, kv.w, a0[3]);
      a1[0] = fmaf(q_1, kv.x, a1[0]); a1[1] = fmaf(q_1, kv.y, a1[1]);
      a1[2] = fmaf(q_1, kv.z, a1[2]); a1[3] = fmaf(q_1, kv.w, a1[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tk * 4 + j;
      if (k < T) {
        sS[tq * SP + k] = a0[j];
        sS[(tq + 16) * SP + k] = a1[j];
      }
    }
  }
  __syncthreads();
  // softmax over keys: warp w handles rows w*4 .. w*4+3
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp * 4; r < warp * 4 + 4; ++r) {
      float* row = sS + r * SP;
      float mx = -INFINITY;
      for (int k = lane; k < T; k += 32) mx = fmaxf(mx, row[k]);
      mx = warp_max(mx);
      float sum = 0.f;
      for (int k = lane; k < T; k += 32) {
        const float e = expf(row[k] - mx);
        row[k] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      for (int k = lane; k < T; k += 32) row[k] *= inv;
    }
  }
  // O = P V : thread -> 2 query rows x 4 dims
  float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
  const int td = tid % 16;   // dims td*4..+3
  for (int k0 = 0; k0 < T; k0 += KT) {
    __syncthreads();
    for (int i = tid; i < KT * kHD / 4; i += 256) {
      const int r = i / (kHD / 4), c4 = i % (kHD / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + r < T) v = *reinterpret_cast<const float4*>(base + (long long)(k0 + r) * ld + voff + c4 * 4);
      *reinterpret_cast<float4*>(sT + r * (kHD + 4) + c4 * 4) = v;
    }
    __syncthreads();
    const int kmax = min(KT, T - k0);
    for (int k = 0; k < kmax; ++k) {
      const float p0 = sS[tq * SP + k0 + k], p1 = sS[(tq + 16) * SP + k0 + k];
      const float4 vv = *reinterpret_cast<const float4*>(sT + k * (kHD + 4) + td * 4);
      o0[0] = fmaf(p0, vv.x, o0[0]); o0[1] = fmaf(p0, vv.y, o0[1]);
      o0[2] = fmaf(p0, vv.z, o0[2]); o0[3] = fmaf(p0, vv.w, o0[3]);
      o1[0] = fmaf(p1, vv.x, o1[0]); o1[1] = fmaf(p1, vv.y, o1[1]);
      o1[2] = fmaf(p1, vv.z, o1[2]); o1[3] = fmaf(p1, vv.w, o1[3]);
    }
  }
  const long long obase = (long long)b * T * ldo + h * kHD + td * 4;
  auto store = [&](int row, const float (&o)[4]) {
    if (row >= T) return;
    const long long off = obase + (long long)row * ldo;
    if (out_lo) {
      float hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split_tf32(o[j], hi[j], lo[j]);
      *reinterpret_cast<float4*>(out + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(out_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    } else {
      *reinterpret_cast<float4*>(out + off) = make_float4(o[0], o[1], o[2], o[3]);
    }
  };
  store(q0 + tq, o0);
  store(q0 + tq + 16, o1);
}

// ---------------------------------------------------------------------------------
// Decoder input: y[b,:] = E[tok[b],:] + PE[pos,:]  (Embed one-hot == row gather,
// layers.py:516-537; FixedEmbed decode branch, layers.py:589-596).
// ---------------------------------------------------------------------------------
__global__ void embed_kernel(const int* __restrict__ tok, const float* __restrict__ emb, int D, int vocab,
                             const float* __restrict__ pe, const int* __restrict__ pos_ptr, float* __restrict__ y, int b0) {
  pdl_wait();
  pdl_trigger();
  const int b = b0 + blockIdx.x;       // b0: first sequence of this sub-batch (decode streams)
  int t = tok[b];
  t = min(max(t, 0), vocab - 1);
  const int pos = *pos_ptr;
  const float4* e = reinterpret_cast<const float4*>(emb + (long long)t * D);
  const float4* p = reinterpret_cast<const float4*>(pe + (long long)pos * D);
  float4* o = reinterpret_cast<float4*>(y + (long long)b * D);
  for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
    const float4 a = __ldg(e + i), c = __ldg(p + i);
    o[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
  }
}

// ---------------------------------------------------------------------------------
// Greedy pick + loop bookkeeping (stands in for t5x's decode loop at num_decodes=1,
// models.py:127): next = argmax(logits) (first maximum wins, like np/jnp.argmax);
// finished sequences emit PAD; EOS (id 1, vocabularies.py:157-159) marks finished.
// The last CTA to arrive advances the shared position and publishes all_finished.
// state: [0] pos, [1] arrival counter, [2] all_finished
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
argmax_step_kernel(const float* __restrict__ logits, int V, int B, int* __restrict__ tok_cur,
                   int* __restrict__ finished, int* __restrict__ tokens_out, int out_ld, int* __restrict__ tok_out_user,
                   int* __restrict__ state, int advance, int b0, const float* __restrict__ emb, const float* __restrict__ pe, int D,
                   float* __restrict__ y_next) {
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ int s_next[2];
  pdl_wait();
  pdl_trigger();
  // B = sequences in the whole batch (the arrival counter spans every sub-batch), b0 = first sequence of this launch
  const int b = b0 + blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* l = logits + (long long)b * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = tid; i < V; i += 256) {
    const float v = l[i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
    const int pos = state[0];
    int nxt = bi;
    if (finished) {
      if (finished[b]) nxt = 0;
      if (nxt == 1) finished[b] = 1;
    }
    s_next[0] = nxt;
    s_next[1] = pos + 1;
    if (tok_cur) tok_cur[b] = nxt;
    if (tokens_out) tokens_out[(long long)b * out_ld + pos] = nxt;
    if (tok_out_user) tok_out_user[b] = nxt;
    if (advance) {
      __threadfence();
      const int done = atomicAdd(&state[1], 1);
      if (done == B - 1) {
        __threadfence();
        int all = 1;
        if (finished) {
          for (int i = 0; i < B; ++i) all &= (*(volatile int*)&finished[i]) != 0;
        } else {
          all = 0;
        }
        state[2] = all;
        state[1] = 0;
        state[0] = pos + 1;
      }
    }
  }
  // generate loop: the next step's decoder input  y[b,:] = E[next,:] + PE[pos + 1,:]  (embed_kernel's work, layers.py:516-537,
  // :589-596) is written here, which takes one launch out of every step.  pos was read before this CTA arrived at the
  // counter, i.e. before the last CTA advances it.
  if (y_next) {
    __syncthreads();
    const int t = min(max(s_next[0], 0), V - 1), p1 = s_next[1];
    const float4* e = reinterpret_cast<const float4*>(emb + (long long)t * D);
    const float4* pp = reinterpret_cast<const float4*>(pe + (long long)p1 * D);
    float4* o = reinterpret_cast<float4*>(y_next + (long long)b * D);
    for (int i = tid; i < D / 4; i += 256) {
      const float4 a = __ldg(e + i), c = __ldg(pp + i);
      o[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
    }
  }
}

// ---------------------------------------------------------------------------------
// T5X decoding.beam_search at num_decodes = 1 (the reference's decode_fn, models.py:127), one CTA per sequence.
// With one live beam the search keeps the 2 best extensions of the live prefix per step:
//   live     continues with the best token that is NOT EOS (cumulative log-probability live_lp);
//   finished whenever EOS is among the 2 best, the hypothesis "prefix + EOS" competes with the best finished one so far on
//            (live_lp_before + log p(EOS)) / brevity_penalty(alpha, length),  brevity_penalty = ((5 + n) / 6)^alpha;
//   a sequence is settled once its finished score beats what the live prefix can still reach,
//            live_lp / brevity_penalty(alpha, max_len)   (`finished[b]`, feeds the loop's all-finished flag).
// tokens_out holds the LIVE sequence; beam1_finalize_kernel cuts it at the winning finish.  Ties follow the reference's
// stable top-k (lower vocabulary id first; an existing finished hypothesis beats an equal newcomer).
// beam_f [B][2] = {live_lp, finished score}, beam_i [B] = length of the best finished hypothesis (0 = none yet).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ bool cand_better(float v, int i, float ov, int oi) { return v > ov || (v == ov && i < oi); }

__global__ void __launch_bounds__(256)
beam1_step_kernel(const float* __restrict__ logits, int V, int B, int* __restrict__ tok_cur, int* __restrict__ finished,
                  int* __restrict__ tokens_out, int out_ld, int* __restrict__ state, float* __restrict__ beam_f,
                  int* __restrict__ beam_i, float alpha, int max_len, const float* __restrict__ emb, const float* __restrict__ pe,
                  int D, float* __restrict__ y_next) {
  __shared__ float s1v[8], s2v[8], smx[8], ssum[8];
  __shared__ int s1i[8], s2i[8];
  __shared__ int s_next[2];
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* l = logits + (long long)b * V;
  // best and second-best NON-EOS candidates (value, id), and the running maximum over everything for the log-sum-exp
  float v1 = -INFINITY, v2 = -INFINITY, mx = -INFINITY;
  int i1 = 0x7fffffff, i2 = 0x7fffffff;
  for (int i = tid; i < V; i += 256) {
    const float v = l[i];
    mx = fmaxf(mx, v);
    if (i == 1) continue;                        // EOS (vocabularies.py:157-159) is handled separately
    if (cand_better(v, i, v1, i1)) { v2 = v1; i2 = i1; v1 = v; i1 = i; }
    else if (cand_better(v, i, v2, i2)) { v2 = v; i2 = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov1 = __shfl_xor_sync(0xffffffffu, v1, o), ov2 = __shfl_xor_sync(0xffffffffu, v2, o);
    const int oi1 = __shfl_xor_sync(0xffffffffu, i1, o), oi2 = __shfl_xor_sync(0xffffffffu, i2, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (cand_better(ov1, oi1, v1, i1)) {         // theirs wins: second = better of (mine first, their second)
      if (cand_better(v1, i1, ov2, oi2)) { v2 = v1; i2 = i1; } else { v2 = ov2; i2 = oi2; }
      v1 = ov1; i1 = oi1;
    } else if (cand_better(ov1, oi1, v2, i2)) { v2 = ov1; i2 = oi1; }
  }
  if (lane == 0) { s1v[warp] = v1; s1i[warp] = i1; s2v[warp] = v2; s2i[warp] = i2; smx[warp] = mx; }
  __syncthreads();
  float gmx = smx[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) gmx = fmaxf(gmx, smx[w]);
  float se = 0.f;
  for (int i = tid; i < V; i += 256) se += expf(l[i] - gmx);
  se = warp_sum(se);
  if (lane == 0) ssum[warp] = se;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) {
      const float ov1 = s1v[w], ov2 = s2v[w];
      const int oi1 = s1i[w], oi2 = s2i[w];
      if (cand_better(ov1, oi1, v1, i1)) {
        if (cand_better(v1, i1, ov2, oi2)) { v2 = v1; i2 = i1; } else { v2 = ov2; i2 = oi2; }
        v1 = ov1; i1 = oi1;
      } else if (cand_better(ov1, oi1, v2, i2)) { v2 = ov1; i2 = oi1; }
    }
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += ssum[w];
    const float lse = gmx + logf(tot);
    const float veos = l[1];
    const int pos = state[0];
    const float live_prev = beam_f[2 * b];
    float fin_score = beam_f[2 * b + 1];
    // EOS is one of the 2 best extensions iff it beats the second-best non-EOS candidate
    if (cand_better(veos, 1, v2, i2)) {
      const float cand = (live_prev + (veos - lse)) / powf((5.0f + (float)(pos + 1)) / 6.0f, alpha);
      if (cand > fin_score) {
        fin_score = cand;
        beam_f[2 * b + 1] = cand;
        beam_i[b] = pos + 1;
      }
    }
    const float live = live_prev + (v1 - lse);
    beam_f[2 * b] = live;
    tok_cur[b] = i1;
    tokens_out[(long long)b * out_ld + pos] = i1;
    finished[b] = (beam_i[b] > 0 && fin_score > live / powf((5.0f + (float)max_len) / 6.0f, alpha)) ? 1 : 0;
    s_next[0] = i1;
    s_next[1] = pos + 1;
    __threadfence();
    const int done = atomicAdd(&state[1], 1);
    if (done == B - 1) {
      __threadfence();
      int all = 1;
      for (int i = 0; i < B; ++i) all &= (*(volatile int*)&finished[i]) != 0;
      state[2] = all;
      state[1] = 0;
      state[0] = pos + 1;
    }
  }
  if (y_next) {
    __syncthreads();
    const int t = min(max(s_next[0], 0), V - 1), p1 = s_next[1];
    const float4* e = reinterpret_cast<const float4*>(emb + (long long)t * D);
    const float4* pp = reinterpret_cast<const float4*>(pe + (long long)p1 * D);
    float4* o = reinterpret_cast<float4*>(y_next + (long long)b * D);
    for (int i = tid; i < D / 4; i += 256) {
      const float4 a = __ldg(e + i), c = __ldg(pp + i);
      o[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
    }
  }
}

// the best finished hypothesis = the live prefix up to its finish + EOS, zeros after; a sequence that never finished keeps
// its live tokens (the reference returns the live beam when nothing finished)
__global__ void beam1_finalize_kernel(int* __restrict__ tokens, int B, int L, int steps, const int* __restrict__ beam_i) {
  const int b = blockIdx.x;
  const int n = beam_i[b];
  if (n <= 0) {
    for (int i = steps + threadIdx.x; i < L; i += blockDim.x) tokens[(long long)b * L + i] = 0;
    return;
  }
  if (threadIdx.x == 0) tokens[(long long)b * L + n - 1] = 1;
  for (int i = n + threadIdx.x; i < L; i += blockDim.x) tokens[(long long)b * L + i] = 0;
}

// ---------------------------------------------------------------------------------
// layers.dot_product_attention (layers.py:85-157) as a standalone op, exact fp32, any head_dim, optional additive bias
// [b, h, q, kv] (layers.py:143-146: combined mask / relative-position bias).  No 1/sqrt(d) scaling (the reference folds it
// into the query initialiser, layers.py:230-234).  One warp per (batch, head, query); scores in shared memory.
// q [b, tq, h, d], k / v [b, tk, h, d], out [b, tq, h, d].
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
dot_product_attention_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                             const float* __restrict__ bias, int B, int TQ, int TK, int H, int D, float* __restrict__ out) {
  extern __shared__ float att_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * 4 + warp;          // (b, h, qi)
  if (item >= (long long)B * H * TQ) return;
  const int qi = (int)(item % TQ), h = (int)((item / TQ) % H), b = (int)(item / ((long long)TQ * H));
  float* sc = att_sm + (size_t)warp * TK;
  const float* qr = q + (((long long)b * TQ + qi) * H + h) * D;
  float mx = -INFINITY;
  for (int j = lane; j < TK; j += 32) {
    const float* kr = k + (((long long)b * TK + j) * H + h) * D;
    float s = 0.f;
    for (int d = 0; d < D; ++d) s = fmaf(qr[d], kr[d], s);
    if (bias) s += bias[(((long long)b * H + h) * TQ + qi) * TK + j];
    sc[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < TK; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.0f / sum;
  float* orow = out + (((long long)b * TQ + qi) * H + h) * D;
  for (int d = lane; d < D; d += 32) {
    float acc = 0.f;
    for (int j = 0; j < TK; ++j) acc = fmaf(sc[j], v[(((long long)b * TK + j) * H + h) * D + d], acc);
    orow[d] = acc * inv;
  }
}

// advance the position without an argmax (decode_step called with tok_out == NULL)
__global__ void advance_pos_kernel(int* state) {
  pdl_wait();
  state[0] += 1;
}

// GenericTokenVocabulary._decode_tf (vocabularies.py:241-271).
__global__ void vocab_decode_kernel(const int* __restrict__ ids, int B, int L, int num_regular, int* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  bool eos = false;
  for (int i = 0; i < L; ++i) {
    const int id = ids[(long long)b * L + i];
    eos = eos || (id == 1);
    int o;
    if (eos) o = -1;
    else if (id >= 3 && id < 3 + num_regular) o = id - 3;
    else o = -2;
    out[(long long)b * L + i] = o;
  }
}

// dst[k, col_off + n*col_stride] = src[k, n] * (g ? g[k] : 1)   -- weight re-layout at model creation
__global__ void scale_copy_cols_kernel(const float* __restrict__ src, int K, int N, const float* __restrict__ g,
                                       float* __restrict__ dst, int ldd, int col_off, int col_stride) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)K * N) return;
  const int k = (int)(i / N), n = (int)(i % N);
  dst[(long long)k * ldd + col_off + (long long)n * col_stride] = src[i] * (g ? g[k] : 1.f);
}

// hi/lo split of a dense fp32 array (tf32x3 operands)
__global__ void split_pair_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  float4 h, l;
  split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
  reinterpret_cast<float4*>(hi)[i] = h;
  reinterpret_cast<float4*>(lo)[i] = l;
}

// src [K, N] row-major -> dst_hi (/dst_lo) [N, K]: the K-major ("transposed") weight layout the UMMA B operand reads.
__global__ void transpose_split_kernel(const float* __restrict__ src, int K, int N, float* __restrict__ dst_hi,
                                       float* __restrict__ dst_lo) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? src[(long long)k * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) {
      const float v = tile[threadIdx.x][i];
      if (dst_lo) {
        float h, l;
        split_tf32(v, h, l);
        dst_hi[(long long)n * K + k] = h;
        dst_lo[(long long)n * K + k] = l;
      } else {
        dst_hi[(long long)n * K + k] = v;
      }
    }
  }
}

}  // namespace mt3
