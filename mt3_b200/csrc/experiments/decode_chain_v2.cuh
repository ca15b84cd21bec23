// Cluster-owned GEMM chains for the decode step (sm_100a), exact fp32.
//
// Sequences are independent through the whole decoder.  The per-node decode GEMMs (decode.cuh) tile one GEMM over all 64
// rows, which makes every dependency between two consecutive GEMMs grid-wide: 41 kernel boundaries per step, each node
// paying ~2.5 us of load / exchange / reduce latency around a multiply that is bound by shared-memory bandwidth.  Here a
// thread-block CLUSTER of 16 CTAs owns 8 sequences and runs a CHAIN of dependent GEMM stages for them with only the
// hardware cluster barrier in between:
//     chain A (after self-attention)   y' = y + o.Wo   and   q_raw = [o | y].[Wo.Wq ; Wq]            (one stage, two GEMMs)
//     chain B (after cross-attention)  y'' = y' + o_c.Wo_c -> g = gated-GELU(rs y''.Wi) -> y''' = y'' + g.Wo2
//                                      -> next layer's q, k, v (+ KV-cache append) or the logits                (four stages)
// CTA r of a cluster computes output columns [r N/16, (r+1) N/16) of every stage for the cluster's 8 rows over the FULL K:
// no split-K, no partial tiles through DSMEM (which cost ~7 B/clk/SM in the tensor-core experiments).  The price is that
// every cluster streams every weight (8 x the L2 -> SM traffic of one GEMM over all rows), so the kernel is written for
// L2 streaming rate:
//   weights   re-laid-out at model creation as per-rank slices  Ws[rank][K][nc (+pad)]: a CTA's k-chunk of a stage is ONE
//             contiguous cp.async.bulk (12-18 KB) into an 8-slot shared-memory ring; a producer lane runs ahead across
//             stage boundaries (~128 KB in flight), so barriers and activation reloads hide under the stream
//   multiply  8 rows x 8 columns per thread (two groups of four contiguous columns), k split over KP thread groups inside
//             the CTA: per 4 k-steps 8 LDS.128 of activations + 8 LDS.128 of weights feed 256 FMAs (1 B/FMA; round 1's chain
//             kernel had 8 x 2 tiles at 2.5 B/FMA and cp.async per thread) -- packed FFMA2
//   reduce    the KP partial sums go through shared memory in k-part order (bit-reproducible), then the fused epilogue:
//             RMSNorm factor (statistic from the staged input rows), residual, gated GELU, KV-cache append (fp32 / fp16),
//             per-tile sums of squares; outputs go to global memory (L2) for the next stage / kernel
//   exchange  __threadfence + barrier.cluster (release / acquire); the next stage re-reads its 8 input rows from L2
// Summation order: k ascending inside a k-part, k-parts in order -- deterministic and independent of the batch composition.
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"
#include "decode.cuh"
#include "tc.cuh"

namespace mt3 {

constexpr int kC2Rows = 8;              // sequences per cluster
constexpr int kC2Cluster = 16;          // CTAs per cluster = column slices per stage
constexpr int kC2Compute = 160;         // compute threads (5 warps)
constexpr int kC2Threads = 192;         // + 1 producer warp
constexpr int kC2Slots = 8;
constexpr int kC2SlotBytes = 18432;     // one k-chunk of a rank's weight slice
constexpr int kC2KMax = 1024;
constexpr int kC2LdA = kC2KMax + 4;     // activation row stride (floats)
constexpr int kC2RedFloats = 9472;      // >= KP * 8 * nc of every stage shape used (72 x 16 x 8 = 9216)
inline size_t chain2_smem_bytes() {
  return (size_t)kC2Rows * kC2LdA * 4 + (size_t)kC2Slots * kC2SlotBytes + (size_t)kC2RedFloats * 4 + 64 + 2 * kC2Slots * 8 + 128;
}

struct C2Gemm {                          // one GEMM of a stage: 8 rows x nc columns of this rank, K deep
  const float* W;                        // rank slices: [16][K][ncp]
  int nc, ncp, K;                        // columns per rank, padded slice row stride, rows
  int KB, KP, CG;                        // rows per streamed chunk (= 4 KP), k-parts, column groups (= nc / 8)
  int norm, epi;
  float* C; int ldc;                     // output; EPI_GATED_GELU writes C[m][n / 2]
  const float* R; int ldr;               // EPI_RESIDUAL source (may alias C: a CTA only touches its own columns)
  int n_split; void* kv; int kv_half, kv_cap, kv_H;   // columns >= n_split: head-major KV-cache append at *pos
  float* ssq_out; int ssq_ld;            // nc == 32 only: per-row sum of squares of this rank's 32 output columns -> [m][rank]
};
struct C2Stage {
  const float* A; int lda; int K0;       // input rows [B][.]: columns [0, K0) from A ...
  const float* A2; int lda2;             // ... columns [K0, K) from A2 (or null)
  int K;
  int n_gemm;
  C2Gemm g[2];
};
struct C2Args {
  C2Stage st[4];
  int n_stages;
  int B;
  float eps;
  const int* pos;
  unsigned long long* trace;             // debug timeline slot or null: [0] min start, [1] max end (ns)
  int dbg_mode;                          // measurement only (mt3_debug_launch): 1 = stream without computing, 2 = compute without streaming
};

__device__ __forceinline__ void c2_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :
               : "r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(kC2Threads, 1)
dec_chain2_kernel(const C2Args a) {
  extern __shared__ __align__(128) unsigned char c2_raw[];
  float* As = reinterpret_cast<float*>(c2_raw);                                   // [8][kC2LdA]
  unsigned char* ring = c2_raw + (size_t)kC2Rows * kC2LdA * 4;                    // [slots][kC2SlotBytes]
  float* red = reinterpret_cast<float*>(ring + (size_t)kC2Slots * kC2SlotBytes);  // [KP][8][nc]
  float* s_rs = red + kC2RedFloats;                                               // [8] RMSNorm factors (+ pad)
  uint64_t* full = reinterpret_cast<uint64_t*>(s_rs + 16);
  uint64_t* empty = full + kC2Slots;

  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int cluster_id = blockIdx.x / kC2Cluster;
  const int row0 = cluster_id * kC2Rows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool tr0 = a.trace != nullptr && tid == 0 && blockIdx.x == 0;
  long long c0 = 0;
  if (a.trace && tid == 0) {
    atomicMin(a.trace, gtime_ns());
    c0 = clock64();
    if (blockIdx.x < 128 && (blockIdx.x & 15) == 0) {     // SM of rank 0 of each cluster, start offset of each cluster (ns / 8)
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      a.trace[8 + (blockIdx.x >> 4)] = smid;
    }
  }

  if (tid == 0) {
    for (int s = 0; s < kC2Slots; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], kC2Compute / 32);
    }
    tc::fence_barrier_init();
  }
  __syncthreads();

  if (warp == kC2Compute / 32) {
    // ---- producer warp: lane 0 walks the chunk schedule of the whole launch; the warp arrives at the cluster barrier of
    // stage s right after issuing its last chunk (so the consumers of every CTA can pass it) and runs ahead into stage
    // s + 1; the matching wait is deferred to just before the next arrive (by then it has usually completed) ----
    int it = 0;
    for (int s = 0; s < a.n_stages; ++s) {
      const C2Stage& S = a.st[s];
      if (lane == 0 && a.dbg_mode != 2) {
        for (int gi = 0; gi < S.n_gemm; ++gi) {
          const C2Gemm& G = S.g[gi];
          const int nch = G.K / G.KB;
          const uint32_t bytes = (uint32_t)(G.KB * G.ncp * 4);
          const float* src = G.W + (size_t)rank * G.K * G.ncp;
          for (int c = 0; c < nch; ++c, ++it) {
            const int slot = it % kC2Slots;
            tc::mbar_wait(&empty[slot], ((it / kC2Slots) & 1) ^ 1);
            tc::mbar_arrive_expect_tx(&full[slot], bytes);
            c2_bulk_g2s(ring + (size_t)slot * kC2SlotBytes, src + (size_t)c * G.KB * G.ncp, bytes, &full[slot]);
          }
        }
      }
      __syncwarp();
      if (s + 1 < a.n_stages) {
        if (s > 0) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      }
    }
    if (a.n_stages > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    return;
  }

  // ---- consumers: 160 threads ----
  int it = 0;
  for (int s = 0; s < a.n_stages; ++s) {
    const C2Stage& S = a.st[s];
    // the stage's 8 input rows from global memory (L2): rows past B are zero
    {
      const int q_per_row = S.K >> 2;
      for (int i = tid; i < kC2Rows * q_per_row; i += kC2Compute) {
        const int r = i / q_per_row, col = (i - r * q_per_row) << 2;
        const bool ok = row0 + r < a.B;
        const long long row = ok ? row0 + r : 0;
        const float* src = (S.A2 != nullptr && col >= S.K0) ? S.A2 + row * S.lda2 + (col - S.K0) : S.A + row * S.lda + col;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(As + r * kC2LdA + col)), "l"(src),
                     "r"(ok ? 16 : 0) : "memory");
      }
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kC2Compute) : "memory");
    if (tr0 && s == 0) a.trace[2] = (unsigned long long)(clock64() - c0);      // input rows staged
    if (S.g[0].norm || (S.n_gemm > 1 && S.g[1].norm)) {          // RMSNorm statistic of the input rows (layers.py:613-616)
      for (int r = warp; r < kC2Rows; r += kC2Compute / 32) {
        float ss = 0.f;
        for (int k = lane * 4; k < S.K; k += 128) {
          const float4 v = *reinterpret_cast<const float4*>(As + r * kC2LdA + k);
          ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        }
        ss = warp_sum(ss);
        if (lane == 0) s_rs[r] = 1.0f / sqrtf(ss / (float)S.K + a.eps);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kC2Compute) : "memory");
    }

    for (int gi = 0; gi < S.n_gemm; ++gi) {
      const C2Gemm& G = S.g[gi];
      const int nc = G.nc, ncp = G.ncp, half = nc >> 1;
      const int cg = tid % G.CG, kp = tid / G.CG;
      const bool active = kp < G.KP;
      const int nch = G.K / G.KB;
      float2 acc[kC2Rows][4];
#pragma unroll
      for (int r = 0; r < kC2Rows; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = make_float2(0.f, 0.f);

      for (int c = 0; c < nch; ++c, ++it) {
        const int slot = it % kC2Slots;
        if (a.dbg_mode != 2) tc::mbar_wait(&full[slot], (it / kC2Slots) & 1);
        if (tr0 && it == 0) a.trace[3] = (unsigned long long)(clock64() - c0);  // first weight chunk landed
        if (active && a.dbg_mode != 1) {
          const float* wt = reinterpret_cast<const float*>(ring + (size_t)slot * kC2SlotBytes) + (kp * 4) * ncp;
          const float* ap = As + c * G.KB + kp * 4;
          float4 av[kC2Rows], w0[4], w1[4];
#pragma unroll
          for (int r = 0; r < kC2Rows; ++r) av[r] = *reinterpret_cast<const float4*>(ap + r * kC2LdA);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            w0[j] = *reinterpret_cast<const float4*>(wt + j * ncp + cg * 4);
            w1[j] = *reinterpret_cast<const float4*>(wt + j * ncp + half + cg * 4);
          }
#pragma unroll
          for (int r = 0; r < kC2Rows; ++r) {
            const float ar[4] = {av[r].x, av[r].y, av[r].z, av[r].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              ffma2(acc[r][0], ar[j], make_float2(w0[j].x, w0[j].y));
              ffma2(acc[r][1], ar[j], make_float2(w0[j].z, w0[j].w));
              ffma2(acc[r][2], ar[j], make_float2(w1[j].x, w1[j].y));
              ffma2(acc[r][3], ar[j], make_float2(w1[j].z, w1[j].w));
            }
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[slot]);
      }

      if (tr0 && s == 0 && gi == 0) a.trace[4] = (unsigned long long)(clock64() - c0);   // multiply of stage 0 / GEMM 0 done
      // ---- k-parts through shared memory, summed in k-part order ----
      if (active) {
#pragma unroll
        for (int r = 0; r < kC2Rows; ++r) {
          float* dst = red + ((kp * kC2Rows + r) * nc);
          *reinterpret_cast<float4*>(dst + cg * 4) = make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y);
          *reinterpret_cast<float4*>(dst + half + cg * 4) = make_float4(acc[r][2].x, acc[r][2].y, acc[r][3].x, acc[r][3].y);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kC2Compute) : "memory");
      const int q_per_row = nc >> 2;
      const int n0 = (int)rank * nc;
      const int pos = (G.kv != nullptr && a.pos) ? *a.pos : 0;
      for (int o0 = 0; o0 < kC2Rows * q_per_row; o0 += kC2Compute) {
        const int o = o0 + tid;
        const bool in_range = o < kC2Rows * q_per_row;
        const int r = in_range ? o / q_per_row : 0, c4 = in_range ? (o - r * q_per_row) << 2 : 0;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in_range) {
          for (int p = 0; p < G.KP; ++p) {
            const float4 t = *reinterpret_cast<const float4*>(red + ((p * kC2Rows + r) * nc) + c4);
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
          }
        }
        const int m = row0 + r, n = n0 + c4;
        const bool valid = in_range && m < a.B;
        if (valid) {
          if (G.norm) {
            const float rs = s_rs[r];
            v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
          }
          if (G.epi == EPI_GATED_GELU) {
            *reinterpret_cast<float2*>(G.C + (long long)m * G.ldc + (n >> 1)) = make_float2(gelu_tanh(v.x) * v.y, gelu_tanh(v.z) * v.w);
          } else {
            if (G.epi == EPI_RESIDUAL) {
              const float4 q = *reinterpret_cast<const float4*>(G.R + (long long)m * G.ldr + n);
              v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
            }
            if (n < G.n_split) {
              *reinterpret_cast<float4*>(G.C + (long long)m * G.ldc + n) = v;
            } else {
              const long long d = kv_dest(m, n - G.n_split, 1, G.kv_cap, G.kv_H, pos);
              if (G.kv_half) {
                __half2* dst = reinterpret_cast<__half2*>(reinterpret_cast<__half*>(G.kv) + d);
                dst[0] = __floats2half2_rn(v.x, v.y);
                dst[1] = __floats2half2_rn(v.z, v.w);
              } else {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(G.kv) + d) = v;
              }
            }
          }
        }
        if (G.ssq_out) {             // nc == 32: eight consecutive lanes hold one row's 32 columns; fixed butterfly order
          float sq = valid ? (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w) : 0.f;
          sq += __shfl_xor_sync(0xffffffffu, sq, 1);
          sq += __shfl_xor_sync(0xffffffffu, sq, 2);
          sq += __shfl_xor_sync(0xffffffffu, sq, 4);
          if (valid && (tid & 7) == 0) G.ssq_out[(long long)m * G.ssq_ld + rank] = sq;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kC2Compute) : "memory");      // red / s_rs are reused by the next GEMM / stage
      if (tr0 && s == 0 && gi == 0) a.trace[5] = (unsigned long long)(clock64() - c0);   // reduce + epilogue of stage 0 / GEMM 0 done
    }
    if (s + 1 < a.n_stages) {
      // stage outputs (global) become visible to the whole cluster before anyone reloads its input rows
      __threadfence();
      asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
  }
  if (a.trace && tid == 0) {
    if (tr0) a.trace[6] = (unsigned long long)(clock64() - c0);
    atomicMax(a.trace + 1, gtime_ns());
  }
}

// ---- host side -------------------------------------------------------------------------------------------------
// thread mapping of a GEMM with nc columns per rank and K rows; false if the shape does not fit the kernel
inline bool chain2_plan(int N, int K, C2Gemm* g) {
  if (N % (kC2Cluster * 8) != 0 || K % 4 != 0 || K > kC2KMax) return false;
  const int nc = N / kC2Cluster;
  const int ncp = nc <= 32 ? nc + 4 : nc;          // padded rows spread the k-parts of a quarter-warp over the banks
  const int CG = nc / 8;
  for (int KP = 32; KP >= 4; KP >>= 1) {
    const int KB = 4 * KP;
    if (CG * KP > kC2Compute || K % KB != 0 || (size_t)KB * ncp * 4 > (size_t)kC2SlotBytes || KP * kC2Rows * nc > kC2RedFloats) continue;
    g->nc = nc; g->ncp = ncp; g->K = K; g->KB = KB; g->KP = KP; g->CG = CG;
    return true;
  }
  return false;
}
inline int64_t chain2_slice_floats(int N, int K) {
  C2Gemm g;
  if (!chain2_plan(N, K, &g)) return 0;
  return (int64_t)kC2Cluster * K * g.ncp;
}

// W [K][N] row-major -> slices [16][K][ncp] (pad columns zero)
__global__ void chain2_slice_kernel(const float* __restrict__ W, int K, int N, int nc, int ncp, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)kC2Cluster * K * ncp;
  if (i >= total) return;
  const int c = (int)(i % ncp);
  const int k = (int)((i / ncp) % K);
  const int r = (int)(i / ((long long)ncp * K));
  out[i] = c < nc ? W[(long long)k * N + r * nc + c] : 0.f;
}

inline int launch_chain2(const C2Args& a, cudaStream_t s) {
  static bool attr_done = false;
  if (!attr_done) {
    MT3_CUDA_CHECK(cudaFuncSetAttribute(dec_chain2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain2_smem_bytes()));
    MT3_CUDA_CHECK(cudaFuncSetAttribute(dec_chain2_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_done = true;
  }
  const int n_clusters = cdiv(a.B, kC2Rows);
  cudaLaunchConfig_t cfg;
  static bool printed = false;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(n_clusters * kC2Cluster);
  cfg.blockDim = dim3(kC2Threads);
  cfg.dynamicSmemBytes = chain2_smem_bytes();
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kC2Cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (!printed && getenv("MT3_CHAIN_INFO")) {
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dec_chain2_kernel, &cfg);
    fprintf(stderr, "[mt3] chain kernel: max active clusters of %d CTAs = %d (%s), smem %zu B, grid %d CTAs\n", kC2Cluster, n,
            cudaGetErrorString(e), chain2_smem_bytes(), n_clusters * kC2Cluster);
    printed = true;
  }
  MT3_CUDA_CHECK(cudaLaunchKernelEx(&cfg, dec_chain2_kernel, a));
  MT3_LAUNCH_CHECK();
  return MT3_OK;
}

}  // namespace mt3
