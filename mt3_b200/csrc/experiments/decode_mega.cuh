// The whole decode step as ONE persistent kernel (sm_100a).  MEASURED ALTERNATIVE, not the default (MT3_DEC_MEGA=1):
// parity-green but 773 vs 563 ms per batch when it was written -- a software grid barrier costs ~5 us against the
// ~0.8 us gap between graph nodes, and one attention CTA per SM holds too few bytes in flight.
// It still uses the unfused 8-GEMM layer and the pull-style cluster reduction of that time.
//
// A greedy decode step is a chain of 67 small, strictly dependent operations (embed, 8 x [QKV GEMM,
// self-attention, out GEMM, q GEMM, cross-attention, out GEMM, MLP-in GEMM, MLP-out GEMM], logits GEMM,
// argmax).  Launched as 67 kernels each pays ~4-5 us of launch/ramp overhead for ~1-3 us of work
// (profiles/: sgemm_dec_cluster 8.6-12 us wall, ~3 us active).  Here the step is a "program" of phases
// executed by a grid of co-resident CTAs (clusters of 8); consecutive phases are separated by a
// device-side grid barrier (~1 us) instead of a kernel boundary.
//
//   GEMM phase     one cluster per 64 x 32 output tile, the 8 CTAs split K; partial tiles are reduced through
//                  distributed shared memory in rank order (same code path as sgemm_dec_cluster_kernel)
//   ATTENTION      one CTA per (sequence, head): producer warp streams K then V tiles with cp.async.bulk into
//                  an mbarrier ring, four consumer warps do scores / softmax / P.V (as dec_attention_bulk_kernel)
//   EMBED, ARGMAX  one CTA per sequence
//
// Cross-phase data lives in global memory; every load of data produced inside the kernel bypasses L1
// (cp.async.cg, TMA bulk copies, ld.global.cg), and the barrier's __threadfence orders the rest.
// All spins are bounded (trap instead of hanging the GPU).
#pragma once

#include "common.cuh"
#include "decode.cuh"

namespace mt3 {

enum { PH_EMBED = 0, PH_GEMM = 1, PH_ATTN = 2, PH_ARGMAX = 3 };

struct MegaPhase {
  int type;
  int kc;                       // GEMM: K chunk per CTA (48 / 64 / 128), K == 8 * kc
  DecGemmArgs g;                // GEMM
  // ATTENTION
  const float* q; const float* kv; int cap; const int* len_ptr; int len_add; float* out;
};

struct MegaArgs {
  const MegaPhase* prog; int n_phases;
  unsigned* bar;                // [0] arrival count, [1] generation
  int B, H, Q, D, V, L, max_len;
  // embed
  const int* tok_in; const float* emb; const float* pe; float* y;
  // argmax / bookkeeping
  const float* logits; int* tok_cur; int* finished; int* tokens_out; int* tok_user; int* state; int greedy;
};

constexpr int kMegaThreads = 160;

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned n_ctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned* gen = bar + 1;
    const unsigned my_gen = *gen;
    __threadfence();
    const unsigned prev = atomicAdd(bar, 1u);
    if (prev == n_ctas - 1) {
      bar[0] = 0;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      unsigned spins = 0;
      while (*gen == my_gen) {
        if (++spins > (1u << 28)) __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// ---- GEMM tile: body of sgemm_dec_cluster_kernel with (tile, rank) instead of blockIdx -----------------
template <int KC>
__device__ __forceinline__ void mega_gemm_tile(const DecGemmArgs& p, int tile, unsigned rank, float* dsm) {
  constexpr int BM = kDecBM, BN = kDecBN, NT = 128, LDA = KC + 4;
  constexpr int WQ = KC * 8 / NT, AQ = KC * 16 / NT;
  float* As = dsm;
  float* Bs = dsm + BM * LDA;
  float* Ps = dsm;
  float* Ss = dsm + BM * BN;
  const int tid = threadIdx.x;
  const bool worker = tid < NT;                 // warp 4 only joins the barriers
  const int tx = tid % 8, ty = (tid / 8) & 15;
  const int n0 = tile * BN;
  const int kbeg = (int)rank * KC;
  if (worker) {
#pragma unroll
    for (int i = 0; i < WQ; ++i) {
      const int idx = tid + i * NT;
      const int kr = idx >> 3, nq = idx & 7;
      const bool ok = n0 + nq * 4 < p.N;
      const float* src = p.W + (long long)(kbeg + kr) * p.ldw + (ok ? n0 + nq * 4 : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(&Bs[kr * BN + nq * 4])), "l"(src),
                   "r"(ok ? 16 : 0) : "memory");
    }
#pragma unroll
    for (int i = 0; i < AQ; ++i) {
      const int idx = tid + i * NT;
      const int row = idx / (KC / 4), kq = idx % (KC / 4);
      const bool ok = row < p.M;
      const float* src = p.A + (long long)(ok ? row : 0) * p.lda + kbeg + kq * 4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(&As[row * LDA + kq * 4])), "l"(src),
                   "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ss = 0.f;
  if (worker) {
#pragma unroll 4
    for (int k = 0; k < KC; k += 4) {
      float4 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(&As[(ty + 16 * i) * LDA + k]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) b[kk] = *reinterpret_cast<const float4*>(&Bs[(k + kk) * BN + tx * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          acc[i][0] = fmaf(av[kk], b[kk].x, acc[i][0]);
          acc[i][1] = fmaf(av[kk], b[kk].y, acc[i][1]);
          acc[i][2] = fmaf(av[kk], b[kk].z, acc[i][2]);
          acc[i][3] = fmaf(av[kk], b[kk].w, acc[i][3]);
        }
      }
    }
    if (p.norm) {        // same order as sgemm_dec_cluster_kernel: thread -> (row tid/2, half of the chunk)
      const float* ar = As + (tid >> 1) * LDA + (tid & 1) * (KC / 2);
#pragma unroll
      for (int k = 0; k < KC / 2; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(ar + k);
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
      }
      ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    }
  }
  __syncthreads();
  if (worker) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(&Ps[(ty + 16 * i) * BN + tx * 4]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if ((tid & 1) == 0) Ss[tid >> 1] = ss;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (worker) {
    const int c2 = (tid & 15) * 2;
    const int n = n0 + c2;
    const int m = (int)rank * 8 + (tid >> 4);
    float2 v = make_float2(0.f, 0.f);
    float sst = 0.f;
    const uint32_t my_p = tc::smem_u32(&Ps[m * BN + c2]);
    const uint32_t my_s = tc::smem_u32(&Ss[m]);
    float2 t[8];
    float tss[8];
#pragma unroll
    for (unsigned s = 0; s < 8; ++s) {
      uint32_t rp, rs_addr;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rp) : "r"(my_p), "r"(s));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rs_addr) : "r"(my_s), "r"(s));
      asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(t[s].x), "=f"(t[s].y) : "r"(rp));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(tss[s]) : "r"(rs_addr));
    }
#pragma unroll
    for (unsigned s = 0; s < 8; ++s) {
      v.x += t[s].x; v.y += t[s].y; sst += tss[s];
    }
    if (m < p.M && n < p.N) {
      const float rs = p.norm ? 1.0f / sqrtf(sst / (float)p.K + p.eps) : 1.f;
      v.x *= rs; v.y *= rs;
      if (p.epi == EPI_GATED_GELU) {
        p.C[(long long)m * p.ldc + (n >> 1)] = gelu_tanh(v.x) * v.y;
      } else {
        if (p.epi == EPI_RESIDUAL) {
          const float2 q = __ldcg(reinterpret_cast<const float2*>(p.R + (long long)m * p.ldr + n));
          v.x += q.x; v.y += q.y;
        }
        if (n < p.n_split) {
          *reinterpret_cast<float2*>(p.C + (long long)m * p.ldc + n) = v;
        } else {
          const int pos = p.hm_pos ? __ldcg(p.hm_pos) : 0;
          *reinterpret_cast<float2*>(p.C1 + kv_dest(m, n - p.n_split, p.hm_rows_per_b, p.hm_cap, p.hm_H, pos)) = v;
        }
      }
    }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- attention item: body of dec_attention_bulk_kernel for one (b, h) ---------------------------------
__device__ __forceinline__ void mega_attention_item(const float* __restrict__ q, int ldq, const float* __restrict__ kv, int H,
                                                    int cap, int len, int max_len, float* __restrict__ out, int ldo, int b,
                                                    int h, float* sm, float* s_stat) {
  float* ring = sm;
  float* sP = ring + kAttStages * kAttTileFloats;
  float* sRed = sP + ((max_len + 3) & ~3);
  uint64_t* full = reinterpret_cast<uint64_t*>(sRed + 8 * 64);
  uint64_t* empty = full + kAttStages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nt = (len + kAttKT - 1) / kAttKT;
  const float* kbase = kv + (((long long)b * 2 + 0) * H + h) * (long long)cap * 64;
  const float* vbase = kv + (((long long)b * 2 + 1) * H + h) * (long long)cap * 64;
  __syncthreads();                               // previous users of this shared memory are done
  if (tid == 0) {
    for (int s = 0; s < kAttStages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 4);
    }
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (warp == 4) {
    if (lane == 0) {
      const uint64_t policy = l2_evict_first_policy();
      for (int j = 0; j < 2 * nt; ++j) {
        const int s = j % kAttStages;
        const uint32_t ph = (j / kAttStages) & 1;
        tc::mbar_wait(&empty[s], ph ^ 1);
        const int t = j < nt ? j : j - nt;
        const int keys = min(kAttKT, len - t * kAttKT);
        const uint32_t bytes = (uint32_t)keys * 64 * 4;
        const float* src = (j < nt ? kbase : vbase) + (long long)t * kAttTileFloats;
        tc::mbar_arrive_expect_tx(&full[s], bytes);
        bulk_g2s(ring + s * kAttTileFloats, src, bytes, &full[s], policy);
      }
    }
  } else {
    const int c = lane & 15, half = lane >> 4;
    const float4 q4 = __ldcg(reinterpret_cast<const float4*>(q + (long long)b * ldq + h * 64 + c * 4));
    float lmax = -INFINITY;
    for (int j = 0; j < nt; ++j) {
      const int s = j % kAttStages;
      tc::mbar_wait(&full[s], (j / kAttStages) & 1);
      const float* tile = ring + s * kAttTileFloats;
      const int k0 = j * kAttKT;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = warp * 8 + i * 2 + half;
        float d = 0.f;
        if (k0 + kk < len) {
          const float4 kx = *reinterpret_cast<const float4*>(tile + kk * 64 + c * 4);
          d = q4.x * kx.x + q4.y * kx.y + q4.z * kx.z + q4.w * kx.w;
        }
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        d += __shfl_xor_sync(0xffffffffu, d, 8);
        if (k0 + kk < len) {
          if (c == 0) sP[k0 + kk] = d;
          lmax = fmaxf(lmax, d);
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&empty[s]);
    }
    lmax = warp_max(lmax);
    if (lane == 0) s_stat[warp] = lmax;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const float mx = fmaxf(fmaxf(s_stat[0], s_stat[1]), fmaxf(s_stat[2], s_stat[3]));
    float lsum = 0.f;
    for (int k = tid; k < len; k += 128) {
      const float e = expf(sP[k] - mx);
      sP[k] = e;
      lsum += e;
    }
    lsum = warp_sum(lsum);
    if (lane == 0) s_stat[4 + warp] = lsum;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const float inv = 1.0f / (s_stat[4] + s_stat[5] + s_stat[6] + s_stat[7]);
    const int kg = tid >> 4, d4 = tid & 15;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = nt; j < 2 * nt; ++j) {
      const int s = j % kAttStages;
      tc::mbar_wait(&full[s], (j / kAttStages) & 1);
      const float* tile = ring + s * kAttTileFloats;
      const int k0 = (j - nt) * kAttKT;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = kg + i * 8;
        if (k0 + kk < len) {
          const float pk = sP[k0 + kk];
          const float4 v = *reinterpret_cast<const float4*>(tile + kk * 64 + d4 * 4);
          acc.x = fmaf(pk, v.x, acc.x); acc.y = fmaf(pk, v.y, acc.y);
          acc.z = fmaf(pk, v.z, acc.z); acc.w = fmaf(pk, v.w, acc.w);
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&empty[s]);
    }
    *reinterpret_cast<float4*>(sRed + kg * 64 + d4 * 4) = acc;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (tid < 64) {
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) s += sRed[g * 64 + tid];
      out[(long long)b * ldo + h * 64 + tid] = s * inv;
    }
  }
  __syncthreads();
  if (tid == 0) {                                 // barriers are re-initialised by the next attention phase
    for (int s = 0; s < kAttStages; ++s) {
      asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(&full[s])) : "memory");
      asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(&empty[s])) : "memory");
    }
  }
}

__global__ void __launch_bounds__(kMegaThreads)
decode_mega_kernel(const MegaArgs a) {
  extern __shared__ __align__(128) float msm[];
  __shared__ float s_stat[8];
  __shared__ float sv[8];
  __shared__ int si[8];
  const unsigned G = gridDim.x;
  const unsigned n_clusters = G / 8;
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const unsigned cluster_id = blockIdx.x / 8;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int ph = 0; ph < a.n_phases; ++ph) {
    const MegaPhase& P = a.prog[ph];
    if (P.type == PH_EMBED) {
      const int pos = __ldcg(a.state);
      for (int b = blockIdx.x; b < a.B; b += G) {
        int t = __ldcg(a.tok_in + b);
        t = min(max(t, 0), a.V - 1);
        const float4* e = reinterpret_cast<const float4*>(a.emb + (long long)t * a.D);
        const float4* p = reinterpret_cast<const float4*>(a.pe + (long long)pos * a.D);
        float4* o = reinterpret_cast<float4*>(a.y + (long long)b * a.D);
        for (int i = tid; i < a.D / 4; i += kMegaThreads) {
          const float4 x = __ldg(e + i), c = __ldg(p + i);
          o[i] = make_float4(x.x + c.x, x.y + c.y, x.z + c.z, x.w + c.w);
        }
      }
    } else if (P.type == PH_GEMM) {
      const int n_tiles = (P.g.N + kDecBN - 1) / kDecBN;
      for (int tile = (int)cluster_id; tile < n_tiles; tile += (int)n_clusters) {   // uniform across the cluster
        if (P.kc == 48) mega_gemm_tile<48>(P.g, tile, rank, msm);
        else if (P.kc == 64) mega_gemm_tile<64>(P.g, tile, rank, msm);
        else mega_gemm_tile<128>(P.g, tile, rank, msm);
      }
    } else if (P.type == PH_ATTN) {
      const int len = (P.len_ptr ? __ldcg(P.len_ptr) : 0) + P.len_add;
      for (int item = blockIdx.x; item < a.B * a.H; item += G)
        mega_attention_item(P.q, a.Q, P.kv, a.H, P.cap, len, a.max_len, P.out, a.Q, item / a.H, item % a.H, msm, s_stat);
    } else {  // PH_ARGMAX: greedy pick + loop bookkeeping (argmax_step_kernel); the position is advanced by CTA 0
      if (a.greedy) {
        const int pos = __ldcg(a.state);
        for (int b = blockIdx.x; b < a.B; b += G) {
          const float* l = a.logits + (long long)b * a.V;
          float best = -INFINITY;
          int bi = 0x7fffffff;
          for (int i = tid; i < a.V; i += kMegaThreads) {
            const float v = __ldcg(l + i);
            if (v > best || (v == best && i < bi)) { best = v; bi = i; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
          }
          __syncthreads();
          if (lane == 0) { sv[warp] = best; si[warp] = bi; }
          __syncthreads();
          if (tid == 0) {
            for (int w = 1; w < kMegaThreads / 32; ++w)
              if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
            int nxt = bi;
            if (a.finished) {
              if (__ldcg(a.finished + b)) nxt = 0;
              if (nxt == 1) a.finished[b] = 1;
            }
            if (a.tok_cur) a.tok_cur[b] = nxt;
            if (a.tokens_out) a.tokens_out[(long long)b * a.L + pos] = nxt;
            if (a.tok_user) a.tok_user[b] = nxt;
          }
        }
      }
    }
    grid_barrier(a.bar, G);
  }
  // bookkeeping after the last barrier: one thread advances the position and publishes all_finished
  if (blockIdx.x == 0 && tid == 0) {
    int all = 0;
    if (a.greedy && a.finished) {
      all = 1;
      for (int i = 0; i < a.B; ++i) all &= __ldcg(a.finished + i) != 0;
    }
    a.state[2] = all;
    a.state[0] = __ldcg(a.state) + 1;
  }
}

inline size_t mega_smem_bytes(int max_len) {
  const size_t gemm = (size_t)(kDecBM * (128 + 4) + 128 * kDecBN) * sizeof(float);
  const size_t attn = dec_attention_smem(max_len);
  return gemm > attn ? gemm : attn;
}

}  // namespace mt3
