// Cluster-local GEMM chains for the decode step (sm_100a).  MEASURED ALTERNATIVE, not the default (MT3_DEC_CHAIN=1):
// parity-green but 1199 vs 566 ms per batch -- every cluster re-reads all weights (8x the L2 traffic) with ~54 KB
// in flight per SM, ~15 us per stage.
//
// Sequences are independent through the whole decoder, so the grid-wide dependency between two consecutive
// small GEMMs (every output column of GEMM i feeds every output of GEMM i+1) is only grid-wide because the
// GEMMs were tiled over columns for ALL rows.  Here a thread-block cluster owns a block of 8 sequences and
// runs a CHAIN of up to 4 dependent GEMMs for them:
//     chain A (after self-attention)   out-proj + residual  ->  RMSNorm + cross-attention query
//     chain B (after cross-attention)  out-proj + residual  ->  RMSNorm + gated-GELU MLP-in  ->  MLP-out +
//                                      residual  ->  RMSNorm + next layer's QKV + KV-cache append (or logits)
// CTA c of the cluster computes output columns [c*N/CL, (c+1)*N/CL) of every stage for the cluster's 8 rows
// with the full K (no split-K, fixed summation order); stage outputs go to global memory (L2) and the only
// synchronisation between stages is the hardware cluster barrier.  The step shrinks from 67 to 35 launches.
// Weights are streamed with cp.async through a 4-deep ring; the first chunks of the NEXT stage are already in
// flight while the current stage reduces, stores and waits at the barrier.
//
// Thread mapping (256 threads): thread = (column pair, k-part).  Per 4 k-steps a thread issues 8 broadcast
// 16-byte reads of the activations (8 rows) and 4 8-byte reads of its weight columns for 64 FMAs; partial sums
// of the k-parts are reduced through shared memory in k-part order (bit-reproducible).
#pragma once

#include "common.cuh"
#include "decode.cuh"

namespace mt3 {

constexpr int kChRows = 8;            // sequences per cluster
constexpr int kChThreads = 256;
constexpr int kChSlot = 4608;         // floats per ring slot: a chunk is KB = (kChSlot / nc) & ~3 weight rows
constexpr int kChStages = 4;          // ring depth
constexpr int kChMaxK = 1024;
constexpr int kChMaxNc = 256;         // columns per CTA and stage (N / cluster size)

struct ChainStage {
  const float* A; int lda; int K;       // [B, K] input rows (global; produced by an earlier kernel or stage)
  const float* W; int N;                // [K, N] row-major weights
  int norm;                             // scale by rsqrt(mean(A[row,:]^2) + eps): needs K == row length
  int epi;                              // EPI_STORE / EPI_RESIDUAL (in place on C) / EPI_GATED_GELU
  float* C; int ldc;                    // output (columns < n_split)
  int n_split; float* kv; int kv_cap; int kv_H; const int* kv_pos;   // columns >= n_split: head-major KV append
};

struct ChainArgs {
  ChainStage st[4];
  int n_stages;
  int B;
  float eps;
};

constexpr int kChRed = kChThreads * 2 * kChRows;   // one float2 x 8 rows per thread
inline size_t chain_smem_bytes() {
  return (size_t)(kChRows * kChMaxK + kChStages * kChSlot + kChRed + 64) * sizeof(float);
}
__host__ __device__ inline int chain_kb(int nc) { return (kChSlot / nc) & ~3; }

__device__ __forceinline__ void chain_issue_chunk(const ChainStage& S, int nc, int n0, int chunk, float* ring_slot) {
  // rows [chunk*KB, +KB) x columns [n0, n0+nc) -> ring_slot[KB][nc]; nc is a multiple of 4
  const int KB = chain_kb(nc);
  const int q_per_row = nc >> 2;
  const int total = KB * q_per_row;
  const int k0 = chunk * KB;
  for (int i = threadIdx.x; i < total; i += kChThreads) {
    const int r = i / q_per_row, q = i % q_per_row;
    const bool ok = k0 + r < S.K;
    const float* src = S.W + (long long)(ok ? k0 + r : 0) * S.N + n0 + q * 4;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(ring_slot + r * nc + q * 4)), "l"(src),
                 "r"(ok ? 16 : 0) : "memory");
  }
}

__global__ void __launch_bounds__(kChThreads, 1)
dec_chain_kernel(const ChainArgs a) {
  extern __shared__ __align__(16) float csm[];
  float* As = csm;                                       // [8][K] activations (row-major)
  float* ring = As + kChRows * kChMaxK;                  // [stages][kChSlot]: chunk = [KB][nc]
  float* red = ring + kChStages * kChSlot;               // [kparts][8][nc] partial sums
  float* s_rs = red + kChRed;                            // [8] RMSNorm factors
  unsigned rank, csize;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  const int cluster_id = blockIdx.x / csize;
  const int row0 = cluster_id * kChRows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  pdl_wait();
  pdl_trigger();
  // prefetch the first chunks of stage 0 (weights do not depend on anything)
  {
    const ChainStage& S = a.st[0];
    const int nc = S.N / (int)csize, n0 = (int)rank * nc, KB = chain_kb(nc), nch = (S.K + KB - 1) / KB;
    for (int c = 0; c < kChStages - 1; ++c) {
      if (c < nch) chain_issue_chunk(S, nc, n0, c, ring + c * kChSlot);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  for (int s = 0; s < a.n_stages; ++s) {
    const ChainStage& S = a.st[s];
    const int K = S.K;
    const int nc = S.N / (int)csize;           // columns of this CTA (multiple of 8)
    const int n0 = (int)rank * nc;
    const int ncp = nc >> 1;                   // column pairs
    const int KP = kChThreads / ncp;           // k-parts (threads beyond KP*ncp idle in the k-loop)
    const int cp = tid % ncp, kpart = tid / ncp;
    const bool active = kpart < KP;
    const int KB = chain_kb(nc);
    const int nch = (K + KB - 1) / KB;

    // ---- activations: 8 rows x K from global (L2) into As; rows past B are zero ----
    {
      const int q_per_row = K >> 2;
      for (int i = tid; i < kChRows * q_per_row; i += kChThreads) {
        const int r = i / q_per_row, q = i % q_per_row;
        const bool ok = row0 + r < a.B;
        const float* src = S.A + (long long)(ok ? row0 + r : 0) * S.lda + q * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(As + r * K + q * 4)), "l"(src),
                     "r"(ok ? 16 : 0) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    float acc[kChRows][2];
#pragma unroll
    for (int r = 0; r < kChRows; ++r) acc[r][0] = acc[r][1] = 0.f;

    for (int c = 0; c < nch; ++c) {
      // keep the ring full: chunk c + stages - 1 of this stage
      {
        const int cn = c + kChStages - 1;
        if (cn < nch) chain_issue_chunk(S, nc, n0, cn, ring + (cn % kChStages) * kChSlot);
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      // commit order is [.. w(c) w(c+1) w(c+2) | A (first iteration only) | w(c+3)]: chunk c has landed when at
      // most 3 groups are pending; in the first iteration the activations (committed after w1, w2) must have
      // landed too, so everything but the newest group is waited for.
      if (c == 0) asm volatile("cp.async.wait_group 1;" ::: "memory");
      else asm volatile("cp.async.wait_group %0;" ::"n"(kChStages - 1) : "memory");
      __syncthreads();
      if (c == 0 && S.norm) {                  // RMSNorm statistic: warp w <-> row w (layers.py:613-616)
        float ss = 0.f;
        for (int k = lane; k < K; k += 32) {
          const float v = As[warp * K + k];
          ss = fmaf(v, v, ss);
        }
        ss = warp_sum(ss);
        if (lane == 0) s_rs[warp] = 1.0f / sqrtf(ss / (float)K + a.eps);
      }
      if (active) {
        const float* wt = ring + (c % kChStages) * kChSlot;
        const int k0 = c * KB;
        for (int q = kpart; q < KB / 4; q += KP) {             // this thread's k-quads inside the chunk
          const int kk = q * 4;
          if (k0 + kk >= K) break;
          float4 av[kChRows];
#pragma unroll
          for (int r = 0; r < kChRows; ++r) av[r] = *reinterpret_cast<const float4*>(As + r * K + k0 + kk);
          float2 wv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) wv[j] = *reinterpret_cast<const float2*>(wt + (kk + j) * nc + cp * 2);
#pragma unroll
          for (int r = 0; r < kChRows; ++r) {
            const float ar[4] = {av[r].x, av[r].y, av[r].z, av[r].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[r][0] = fmaf(ar[j], wv[j].x, acc[r][0]);
              acc[r][1] = fmaf(ar[j], wv[j].y, acc[r][1]);
            }
          }
        }
      }
      __syncthreads();                        // everyone is done with ring slot c % stages before it is refilled
    }
    // the next stage's first weight chunks start streaming now, under the reduction / store / barrier
    if (s + 1 < a.n_stages) {
      const ChainStage& Sn = a.st[s + 1];
      const int ncn = Sn.N / (int)csize, n0n = (int)rank * ncn, KBn = chain_kb(ncn), nchn = (Sn.K + KBn - 1) / KBn;
      for (int c = 0; c < kChStages - 1; ++c) {
        if (c < nchn) chain_issue_chunk(Sn, ncn, n0n, c, ring + c * kChSlot);
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
    // ---- reduce the k-parts (fixed order) ----
    if (active) {
#pragma unroll
      for (int r = 0; r < kChRows; ++r)
        *reinterpret_cast<float2*>(red + ((kpart * kChRows + r) * nc) + cp * 2) = make_float2(acc[r][0], acc[r][1]);
    }
    __syncthreads();
    for (int o = tid; o < kChRows * ncp; o += kChThreads) {
      const int r = o / ncp, p2 = o % ncp;
      float2 v = make_float2(0.f, 0.f);
      for (int kp = 0; kp < KP; ++kp) {
        const float2 t = *reinterpret_cast<const float2*>(red + ((kp * kChRows + r) * nc) + p2 * 2);
        v.x += t.x; v.y += t.y;
      }
      const int m = row0 + r;
      if (m >= a.B) continue;
      const int n = n0 + p2 * 2;
      const float rs = S.norm ? s_rs[r] : 1.f;
      v.x *= rs; v.y *= rs;
      if (S.epi == EPI_GATED_GELU) {
        S.C[(long long)m * S.ldc + (n >> 1)] = gelu_tanh(v.x) * v.y;
        continue;
      }
      if (S.epi == EPI_RESIDUAL) {
        const float2 q = __ldcg(reinterpret_cast<const float2*>(S.C + (long long)m * S.ldc + n));
        v.x += q.x; v.y += q.y;
      }
      if (n < S.n_split) {
        *reinterpret_cast<float2*>(S.C + (long long)m * S.ldc + n) = v;
      } else {
        const int pos = S.kv_pos ? __ldcg(S.kv_pos) : 0;
        *reinterpret_cast<float2*>(S.kv + kv_dest(m, n - S.n_split, 1, S.kv_cap, S.kv_H, pos)) = v;
      }
    }
    // stage outputs (global) become visible to the whole cluster; also fences the reuse of As / red
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace mt3
