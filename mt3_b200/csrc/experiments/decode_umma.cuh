// Decode-step GEMM on the 5th-generation tensor cores (sm_100a), M = B <= 64 sequences.
//
// Why: the exact-fp32 cluster kernel (decode.cuh) is bound by shared-memory bandwidth -- a 4 x 4 register tile needs 8
// LDS.128 per 64 FMAs, which caps it at ~45 % of the fp32 pipe (profiles/r02_call1_*: FFMA2 and twice the warps change
// nothing) -- and the decode step spends more time in its 41 GEMM nodes than in attention.  tcgen05.mma reads its
// operands from shared memory through descriptors, not through the LSU.
//
// Orientation ("swap A/B"): the WEIGHTS are the M operand.  D^T[n, b] = sum_k W^T[n, k] X[b, k]: one CTA owns 128 output
// columns (n) of all 64 sequences (b) and ONE 32- or 64-deep slice of K; an instruction is M128 x N64 x K8, so a CTA issues
// 12 or 24 MMAs (round 1's first tcgen05 decode kernel, csrc/experiments/decode_tc.cuh, used M128 x N32 tiles with the
// batch on the M side: 4x less work per instruction at the same ~90-cycle issue cost, and lost to the FMA loop).
//
//   cluster   the S = 16 CTAs of one 128-column tile form a thread-block cluster and split K (rank r: k-blocks
//             [r nkb, (r+1) nkb) of 32; blocks past K are out of bounds for the tensor map and arrive as zeros)
//   load      TMA 2-D, 128B swizzle: W^T tile [128 n x 32 k] (16 KB) and X tile [64 b x 32 k] (8 KB) per k-block, one
//             mbarrier; weights are issued before the PDL wait
//   split     3xTF32: kind::tf32 reads the top 19 bits of an fp32 word, i.e. the raw tile IS the "hi" operand; only
//             lo = x - trunc_tf32(x) is materialised (second buffer, same swizzled offsets).  The RMSNorm sum of squares
//             of the X rows rides on the same pass.
//   multiply  one elected thread: per k8 step  lo.hi + hi.lo + hi.hi  into a 64-column fp32 TMEM accumulator
//   drain     tcgen05.ld: thread n holds D^T[n, 0..63]; it PUSHES the 64/S rows owned by rank r into rank r's shared
//             memory (st.shared::cluster), one cluster barrier, then rank r sums its rows over the S sources in rank
//             order (bit-reproducible) and runs the fused epilogue (RMSNorm factor, residual, gated GELU, KV-cache
//             append in fp32 or fp16, per-tile sums of squares) with row-contiguous, fully coalesced stores.
// Precision: 3xTF32 drops only the lo.lo term (~2^-22 relative): logits within 5e-6 of the float64 oracle
// (tests/test_gpu_parity.py), the same class as the encoder's MT3_GEMM_TF32X3 GEMMs.  MT3_GEMM_FP32_SIMT models keep
// the exact-fp32 cluster kernel.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "decode.cuh"
#include "gemm_tc.cuh"
#include "tc.cuh"

namespace mt3 {

constexpr int kDuS = 16;                     // cluster size (K split)
constexpr int kDuR = kDecBM / kDuS;          // rows of the batch each rank owns (4)
constexpr int kDuBN = 128;                   // output columns per CTA (the MMA's M)
constexpr int kDuWBytes = kDuBN * 128;       // one k-block of W^T: 128 rows x 32 fp32
constexpr int kDuXBytes = kDecBM * 128;      // one k-block of X: 64 rows x 32 fp32
constexpr int kDuRedFloats = kDuS * kDuBN * kDuR;   // [src rank][n][local row]: 32 KB
inline size_t dec_umma_smem(int nkb) {
  return (size_t)nkb * 2 * (kDuWBytes + kDuXBytes) + (size_t)(kDuRedFloats + 64) * sizeof(float) + 64 + 1024;
}

struct DecUmmaMaps { CUtensorMap w, xa, xb; };   // W^T [N, K]; X source 0 [M, K0 or K]; X source 1 [M, K - K0] (fused launches)

template <bool TRACE>
__device__ __forceinline__ void dec_umma_body(const CUtensorMap* tmW, const CUtensorMap* tmXa, const CUtensorMap* tmXb,
                                              const DecGemmArgs& p, const int nkb, const int tile_x) {
  constexpr int S = kDuS, R = kDuR, BN = kDuBN;
  extern __shared__ uint8_t du_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(du_raw) + 1023) & ~(uintptr_t)1023);
  const int w_bytes = nkb * kDuWBytes, x_bytes = nkb * kDuXBytes;
  uint8_t* W_hi = sm;
  uint8_t* W_lo = W_hi + w_bytes;
  uint8_t* X_hi = W_lo + w_bytes;
  uint8_t* X_lo = X_hi + x_bytes;
  float* Red = reinterpret_cast<float*>(X_lo + x_bytes);       // [S src][128 n][R rows]
  float* Rss = Red + kDuRedFloats;                             // [S src][R rows]
  uint64_t* full = reinterpret_cast<uint64_t*>(Rss + 64);
  uint64_t* done = full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = tile_x * BN;
  const int kb0 = (int)rank * nkb;                              // first k-block of this rank
  const bool tr = TRACE && p.trace != nullptr && tid == 0;
  const bool tr0 = tr && tile_x == 0 && rank == 0;
  long long c0 = 0;
  if (tr) {
    atomicMin(p.trace, gtime_ns());
    c0 = clock64();
  }
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");     // "I am running" (see decode.cuh)

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(tmW);
    tc::prefetch_tmap(tmXa);
    tc::mbar_init(full, 1);
    tc::mbar_init(done, 1);
    tc::fence_barrier_init();
    tc::mbar_arrive_expect_tx(full, (uint32_t)(w_bytes + x_bytes));
    for (int kb = 0; kb < nkb; ++kb)                 // weights first: they do not depend on the previous kernel
      tc::tma_load_2d(W_hi + kb * kDuWBytes, tmW, full, (kb0 + kb) * 32, n0);
    pdl_wait();
    for (int kb = 0; kb < nkb; ++kb) {
      const int k = (kb0 + kb) * 32;
      if (p.A2 != nullptr && k >= p.K0) tc::tma_load_2d(X_hi + kb * kDuXBytes, tmXb, full, k - p.K0, 0);
      else tc::tma_load_2d(X_hi + kb * kDuXBytes, tmXa, full, k, 0);
    }
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 64);
    tc::tmem_relinquish();
  }
  pdl_wait();
  pdl_trigger();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  tc::mbar_wait(full, 0);
  if (tr0) p.trace[2] = (unsigned long long)(clock64() - c0);      // loads landed

  // ---- split pass: lo = x - trunc_tf32(x) (the raw tile stays as the hi operand).  W: linear sweep.  X: thread ->
  // (row tid/2, 64-byte half of the row's 128-byte line) of every k-block; the swizzle only permutes 16-byte chunks
  // inside a line, so offsets are kept and the row's sum of squares does not care.
  float ss = 0.f;
  {
    const float4* wh = reinterpret_cast<const float4*>(W_hi);
    float4* wl = reinterpret_cast<float4*>(W_lo);
    const int nw = nkb * (kDuWBytes / 16);
#pragma unroll 4
    for (int i = tid; i < nw; i += 128) {
      const float4 v = wh[i];
      float4 hi, lo;
      split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
      wl[i] = lo;
    }
    const int r = tid >> 1, h = tid & 1;
    const int line = (r >> 3) * 1024 + (r & 7) * 128 + h * 64;
    for (int kb = 0; kb < nkb; ++kb) {
      const float4* ph = reinterpret_cast<const float4*>(X_hi + kb * kDuXBytes + line);
      float4* pl = reinterpret_cast<float4*>(X_lo + kb * kDuXBytes + line);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = ph[j];
        float4 hi, lo;
        split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        pl[j] = lo;
      }
    }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  }
  tc::fence_proxy_async();                    // generic-proxy writes above -> visible to the MMA's async-proxy reads
  __syncthreads();

  if (warp == 1) {
    if (tc::elect_one()) {
      tc::tc_fence_after();
      constexpr uint32_t idesc = tc::make_idesc(tc::kFmtTF32, BN, kDecBM, 0, 0);      // M = 128 (n), N = 64 (b)
      for (int kb = 0; kb < nkb; ++kb) {
        const uint64_t w_hi = tc::smem_desc_k_sw128(tc::smem_u32(W_hi + kb * kDuWBytes));
        const uint64_t w_lo = tc::smem_desc_k_sw128(tc::smem_u32(W_lo + kb * kDuWBytes));
        const uint64_t x_hi = tc::smem_desc_k_sw128(tc::smem_u32(X_hi + kb * kDuXBytes));
        const uint64_t x_lo = tc::smem_desc_k_sw128(tc::smem_u32(X_lo + kb * kDuXBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);       // 8 tf32 = 32 bytes along K inside the swizzle row
          tc::mma_tf32(tmem_base, w_lo + adv, x_hi + adv, idesc, (uint32_t)((kb | k) != 0));   // small terms first
          tc::mma_tf32(tmem_base, w_hi + adv, x_lo + adv, idesc, 1u);
          tc::mma_tf32(tmem_base, w_hi + adv, x_hi + adv, idesc, 1u);
        }
      }
      tc::mma_commit(done);
    }
    __syncwarp();
  }

  // ---- drain: thread n = tid holds D^T[n, b] for b = 0..63 ----
  uint32_t acc[64];
  tc::mbar_wait(done, 0);
  tc::tc_fence_after();
  {
    uint32_t lo32[32], hi32[32];
    tc::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), lo32);
    tc::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + 32u, hi32);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) { acc[j] = lo32[j]; acc[32 + j] = hi32[j]; }
  }
  if (tr0) p.trace[3] = (unsigned long long)(clock64() - c0);      // split + MMA + TMEM read done
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");      // every peer is running
  {
    const uint32_t mine = tc::smem_u32(Red) + (uint32_t)(((int)rank * BN + tid) * R * 4);   // [src = my rank][n = tid][..] in the owner
#pragma unroll
    for (int o = 0; o < S; ++o)                                                         // rows o R .. o R + R - 1 -> rank o
      st_cluster_f4(cluster_map(mine, (unsigned)o), make_float4(__uint_as_float(acc[o * R]), __uint_as_float(acc[o * R + 1]),
                                                                __uint_as_float(acc[o * R + 2]), __uint_as_float(acc[o * R + 3])));
    if (p.norm && (tid & 1) == 0) {
      const int row = tid >> 1;
      st_cluster_f1(cluster_map(tc::smem_u32(Rss) + (uint32_t)(((int)rank * R + (row % R)) * 4), (unsigned)(row / R)), ss);
    }
  }
  tc::tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (tr0) p.trace[4] = (unsigned long long)(clock64() - c0);      // partials exchanged
  if (warp == 1) tc::tmem_dealloc(tmem_base, 64);

  // ---- reduce (rank order) + fused epilogue: this rank's R rows, column n0 + tid ----
  float v[R] = {0.f, 0.f, 0.f, 0.f};
  float sst[R] = {0.f, 0.f, 0.f, 0.f};
  const float4* red4 = reinterpret_cast<const float4*>(Red);
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const float4 q = red4[s * BN + tid];
    v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w;
    if (p.norm) {
#pragma unroll
      for (int j = 0; j < R; ++j) sst[j] += Rss[s * R + j];
    }
  }
  const int n = n0 + tid;
  const int pos = (p.hm_pos && n >= p.n_split) ? *p.hm_pos : 0;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int m = (int)rank * R + j;
    const bool valid = m < p.M && n < p.N;
    float x = v[j];
    if (p.norm) x *= 1.0f / sqrtf(sst[j] / (float)p.K + p.eps);
    if (p.epi == EPI_GATED_GELU) {                                 // columns (2 i, 2 i + 1) = (wi_0[i], wi_1[i])
      const float other = __shfl_xor_sync(0xffffffffu, x, 1);
      if (valid && (tid & 1) == 0) p.C[(long long)m * p.ldc + (n >> 1)] = gelu_tanh(x) * other;
    } else if (valid) {
      if (p.epi == EPI_RESIDUAL) x += p.R[(long long)m * p.ldr + n];
      if (n < p.n_split) {
        p.C[(long long)m * p.ldc + n] = x;
      } else {
        const long long d = kv_dest(m, n - p.n_split, p.hm_rows_per_b, p.hm_cap, p.hm_H, pos);
        if (p.kv_half) reinterpret_cast<__half*>(p.C1)[d] = __float2half_rn(x);
        else reinterpret_cast<float*>(p.C1)[d] = x;
      }
    }
    if (p.ssq_out) {                                               // per (row, 32-column tile): fixed butterfly order
      float sq = valid ? x * x : 0.f;
      sq = warp_sum(sq);
      if (lane == 0 && m < p.M) p.ssq_out[(long long)m * p.ssq_ld + (n0 >> 5) + warp] = sq;
    }
  }
  if (tr) {
    if (tr0) p.trace[5] = (unsigned long long)(clock64() - c0);
    atomicMax(p.trace + 1, gtime_ns());
  }
}

template <bool TRACE>
__global__ void __launch_bounds__(128)
dec_gemm_umma_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmXa,
                     const __grid_constant__ CUtensorMap tmXb, const DecGemmArgs p, const int nkb) {
  dec_umma_body<TRACE>(&tmW, &tmXa, &tmXb, p, nkb, (int)blockIdx.x);
}

// Two GEMMs that read the same inputs in one launch (column tiles [0, tiles0) belong to p0): the self-attention
// out-projection and the cross-attention query projection through the precomposed [Wo.Wq ; Wq] block.
template <bool TRACE>
__global__ void __launch_bounds__(128)
dec_gemm_umma2_kernel(const __grid_constant__ CUtensorMap tmW0, const __grid_constant__ CUtensorMap tmX0,
                      const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmX1a,
                      const __grid_constant__ CUtensorMap tmX1b, const DecGemmArgs p0, const DecGemmArgs p1, const int nkb0,
                      const int nkb1, const int tiles0) {
  if ((int)blockIdx.x < tiles0) dec_umma_body<TRACE>(&tmW0, &tmX0, &tmX0, p0, nkb0, (int)blockIdx.x);
  else dec_umma_body<TRACE>(&tmW1, &tmX1a, &tmX1b, p1, nkb1, (int)blockIdx.x - tiles0);
}

inline bool dec_gemm_umma_supported(const DecGemmArgs& a) {
  return a.M <= kDecBM && a.K % 32 == 0 && a.K <= kDuS * 2 * 32 && a.N % 2 == 0 && a.lda % 4 == 0 &&
         (a.A2 == nullptr || (a.K0 % 32 == 0 && a.lda2 % 4 == 0));
}
inline int dec_umma_nkb(int K) { return cdiv(K / 32, kDuS); }

inline int dec_umma_attrs(size_t smem) {
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MT3_CUDA_CHECK(cudaFuncSetAttribute(dec_gemm_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MT3_CUDA_CHECK(cudaFuncSetAttribute(dec_gemm_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MT3_CUDA_CHECK(cudaFuncSetAttribute(dec_gemm_umma2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MT3_CUDA_CHECK(cudaFuncSetAttribute(dec_gemm_umma2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (attr_smem == 0) {
      MT3_CUDA_CHECK(cudaFuncSetAttribute(dec_gemm_umma_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      MT3_CUDA_CHECK(cudaFuncSetAttribute(dec_gemm_umma_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      MT3_CUDA_CHECK(cudaFuncSetAttribute(dec_gemm_umma2_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      MT3_CUDA_CHECK(cudaFuncSetAttribute(dec_gemm_umma2_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    attr_smem = smem;
  }
  return MT3_OK;
}

inline int launch_dec_gemm_umma(const DecUmmaMaps& maps, const DecGemmArgs& a, cudaStream_t s, bool pdl) {
  const int nkb = dec_umma_nkb(a.K);
  const size_t smem = dec_umma_smem(nkb);
  int r = dec_umma_attrs(dec_umma_smem(2));
  if (r != MT3_OK) return r;
  const dim3 grid(cdiv(a.N, kDuBN), kDuS);
  if (a.trace)
    MT3_CUDA_CHECK(launch_kernel_cluster(dec_gemm_umma_kernel<true>, grid, dim3(128), smem, s, pdl, (unsigned)kDuS, maps.w, maps.xa, maps.xb, a, nkb));
  else
    MT3_CUDA_CHECK(launch_kernel_cluster(dec_gemm_umma_kernel<false>, grid, dim3(128), smem, s, pdl, (unsigned)kDuS, maps.w, maps.xa, maps.xb, a, nkb));
  MT3_LAUNCH_CHECK();
  return MT3_OK;
}

inline int launch_dec_gemm_umma2(const DecUmmaMaps& m0, const DecUmmaMaps& m1, const DecGemmArgs& a0, const DecGemmArgs& a1,
                                 cudaStream_t s, bool pdl) {
  const int nkb0 = dec_umma_nkb(a0.K), nkb1 = dec_umma_nkb(a1.K);
  const size_t smem = dec_umma_smem(nkb0 > nkb1 ? nkb0 : nkb1);
  int r = dec_umma_attrs(dec_umma_smem(2));
  if (r != MT3_OK) return r;
  const int tiles0 = cdiv(a0.N, kDuBN), tiles1 = cdiv(a1.N, kDuBN);
  const dim3 grid(tiles0 + tiles1, kDuS);
  if (a0.trace)
    MT3_CUDA_CHECK(launch_kernel_cluster(dec_gemm_umma2_kernel<true>, grid, dim3(128), smem, s, pdl, (unsigned)kDuS, m0.w, m0.xa, m1.w,
                                         m1.xa, m1.xb, a0, a1, nkb0, nkb1, tiles0));
  else
    MT3_CUDA_CHECK(launch_kernel_cluster(dec_gemm_umma2_kernel<false>, grid, dim3(128), smem, s, pdl, (unsigned)kDuS, m0.w, m0.xa, m1.w,
                                         m1.xa, m1.xb, a0, a1, nkb0, nkb1, tiles0));
  MT3_LAUNCH_CHECK();
  return MT3_OK;
}

}  // namespace mt3
