// MEASURED ALTERNATIVE, not the default (MT3_DEC_GEMM_MODE=1 with MT3_DEC_TC=1): parity-green (logits within 5e-6 of
// the fp64 oracle) but 678 vs 521 ms per batch -- 24-48 DEPENDENT N=32 MMAs per CTA cost ~90 cycles each, more than
// the FMA loop they replace.
//
// Decode-step GEMM on the 5th-generation tensor cores (sm_100a): M = B <= 64 sequences, one 8-CTA cluster per
// 64 x 32 output tile, split-K across the cluster, push-style DSMEM reduction (see sgemm_dec_cluster_kernel in
// decode.cuh for the reduction / epilogue contract -- this kernel only replaces the load and multiply phases).
//
//   load      TMA 2-D (128B swizzle) of the CTA's K chunk: activations [64 rows x 32 fp32] per k-block and the
//             K-major weight copy W^T [32 output columns x 32 fp32] per k-block, one mbarrier.  Chunks past K
//             (K = 384 GEMMs: ranks 6, 7) are out of bounds for the tensor map and arrive as zeros.
//   split     x -> hi = x truncated to tf32 (in place), lo = x - hi (second buffer, same swizzled offset);
//             the RMSNorm sum of squares of the chunk rides on the same pass.
//   multiply  one elected thread issues tcgen05.mma.kind::tf32 M=128, N=32, K=8: lo*hi + hi*lo + hi*hi per k-step
//             (3xTF32, fp32-faithful) into a 32-column TMEM accumulator.  Only TMEM lanes 0..63 (the 64
//             sequences) are meaningful: the A descriptor's rows 64..127 read whatever follows the 64-row tile
//             in shared memory, which only ever lands in accumulator lanes 64..127 that nobody reads.
//   drain     warps 0/1 read their 32 lanes x 32 columns with tcgen05.ld and push 8-row slices to the owner ranks.
//
// The measured cost of the multiply phase of the FMA / mma.sync kernels was 1.0-3.6 us per node
// (profiles/r01_call13_trace_step_*.log); this kernel's came out at 1.7-3.2 us (profiles/r01_call14_*).
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "decode.cuh"
#include "tc.cuh"

namespace mt3 {

constexpr int kDtcABytes = 64 * 128;       // one k-block of activations: 64 rows x 32 fp32 (swizzled 128-byte rows)
constexpr int kDtcBBytes = 32 * 128;       // one k-block of W^T: 32 output columns x 32 fp32
inline size_t dec_tc_smem(int nkb) {
  return (size_t)nkb * 2 * (kDtcABytes + kDtcBBytes) + (size_t)(kDecRedFloats + 64) * sizeof(float) + 64 + 1024;
}

template <int MODE, bool TRACE>     // MODE 1: 3xTF32, MODE 2: 1xTF32
__global__ void __launch_bounds__(128)
dec_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const DecGemmArgs p,
                   const int nkb) {
  constexpr int BN = kDecBN;
  extern __shared__ uint8_t dtc_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dtc_raw) + 1023) & ~(uintptr_t)1023);
  const int a_bytes = nkb * kDtcABytes, b_bytes = nkb * kDtcBBytes;
  uint8_t* A_hi = sm;
  uint8_t* A_lo = A_hi + a_bytes;
  uint8_t* B_hi = A_lo + a_bytes;
  uint8_t* B_lo = B_hi + b_bytes;
  float* Red = reinterpret_cast<float*>(B_lo + b_bytes);       // [8 src][8 rows][BN]
  float* Rss = Red + kDecRedFloats;                            // [8 src][8 rows]
  uint64_t* full = reinterpret_cast<uint64_t*>(Rss + 64);
  uint64_t* done = full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.y * nkb * 32;
  const bool tr = TRACE && p.trace != nullptr && tid == 0;
  const bool tr0 = tr && blockIdx.x == 0 && blockIdx.y == 0;
  long long c0 = 0;
  if (tr) {
    atomicMin(p.trace, gtime_ns());
    c0 = clock64();
  }
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");     // "I am running" (see decode.cuh)

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
    tc::mbar_init(full, 1);
    tc::mbar_init(done, 1);
    tc::fence_barrier_init();
    tc::mbar_arrive_expect_tx(full, (uint32_t)(a_bytes + b_bytes));
    for (int kb = 0; kb < nkb; ++kb)                 // weights first: they do not depend on the previous kernel
      tc::tma_load_2d(B_hi + kb * kDtcBBytes, &tmB, full, kbeg + kb * 32, n0);
    pdl_wait();
    for (int kb = 0; kb < nkb; ++kb) tc::tma_load_2d(A_hi + kb * kDtcABytes, &tmA, full, kbeg + kb * 32, 0);
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 32);
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

  // ---- split pass.  A: thread -> (row tid/2, 64-byte half of the row's 128-byte line) of every k-block; the
  // swizzle only permutes 16-byte chunks inside a line, so hi/lo keep their offsets and the row's sum of squares
  // does not care.  B: plain linear sweep.
  float ss = 0.f;
  {
    const int r = tid >> 1, h = tid & 1;
    const int line = (r >> 3) * 1024 + (r & 7) * 128 + h * 64;
    for (int kb = 0; kb < nkb; ++kb) {
      float4* ph = reinterpret_cast<float4*>(A_hi + kb * kDtcABytes + line);
      float4* pl = reinterpret_cast<float4*>(A_lo + kb * kDtcABytes + line);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = ph[j];
        float4 hi, lo;
        split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        ph[j] = hi;
        if (MODE == 1) pl[j] = lo;
      }
    }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    float4* bh = reinterpret_cast<float4*>(B_hi);
    float4* bl = reinterpret_cast<float4*>(B_lo);
    for (int i = tid; i < nkb * (kDtcBBytes / 16); i += 128) {
      const float4 v = bh[i];
      float4 hi, lo;
      split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
      bh[i] = hi;
      if (MODE == 1) bl[i] = lo;
    }
  }
  tc::fence_proxy_async();                    // generic-proxy writes above -> visible to the MMA's async-proxy reads
  __syncthreads();

  if (warp == 1) {
    if (tc::elect_one()) {
      tc::tc_fence_after();
      constexpr uint32_t idesc = tc::make_idesc(tc::kFmtTF32, 128, BN, 0, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const uint64_t a_hi = tc::smem_desc_k_sw128(tc::smem_u32(A_hi + kb * kDtcABytes));
        const uint64_t a_lo = tc::smem_desc_k_sw128(tc::smem_u32(A_lo + kb * kDtcABytes));
        const uint64_t b_hi = tc::smem_desc_k_sw128(tc::smem_u32(B_hi + kb * kDtcBBytes));
        const uint64_t b_lo = tc::smem_desc_k_sw128(tc::smem_u32(B_lo + kb * kDtcBBytes));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);       // 8 tf32 = 32 bytes along K inside the swizzle row
          const uint32_t acc = (kb | k) != 0;
          if (MODE == 1) {
            tc::mma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, acc);
            tc::mma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
            tc::mma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, 1u);
          } else {
            tc::mma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, acc);
          }
        }
      }
      tc::mma_commit(done);
    }
    __syncwarp();
  }

  const uint32_t red_base = tc::smem_u32(Red) + rank * (8 * BN * 4);     // my slot [rank][..] in the owner's Red
  const uint32_t rss_base = tc::smem_u32(Rss) + rank * (8 * 4);
  uint32_t acc[32];
  if (warp < 2) {                           // TMEM lanes 0..63 = the 64 sequences
    tc::mbar_wait(done, 0);
    tc::tc_fence_after();
    tc::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), acc);
    tc::tmem_ld_wait();
  }
  if (tr0) p.trace[3] = (unsigned long long)(clock64() - c0);      // split + MMA + TMEM read done
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");      // every peer is running
  if (warp < 2) {
    const int row = warp * 32 + lane;
    const uint32_t dst = cluster_map(red_base + (uint32_t)((row & 7) * BN * 4), (unsigned)(row >> 3));
#pragma unroll
    for (int j = 0; j < 8; ++j)
      st_cluster_f4(dst + j * 16, make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                              __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3])));
  }
  if (p.norm && (tid & 1) == 0) {
    const int row = tid >> 1;
    st_cluster_f1(cluster_map(rss_base + (uint32_t)((row & 7) * 4), (unsigned)(row >> 3)), ss);
  }
  tc::tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (tr0) p.trace[4] = (unsigned long long)(clock64() - c0);      // partials exchanged
  if (warp == 1) tc::tmem_dealloc(tmem_base, 32);

  dec_reduce_epilogue<8>(p, Red, Rss, rank, n0);
  if (tr) {
    if (tr0) p.trace[5] = (unsigned long long)(clock64() - c0);
    atomicMax(p.trace + 1, gtime_ns());
  }
}

// Host side: tensor maps are built once per (buffer, shape) and cached on the model.
struct DecTcMaps { CUtensorMap a, b; };

template <int MODE>
inline int launch_dec_gemm_tc_mode(const DecTcMaps& maps, const DecGemmArgs& a, cudaStream_t s, bool pdl) {
  const int kc = ((a.K + 7) / 8 + 31) / 32 * 32;            // K chunk per rank, whole 32-float k-blocks
  const int nkb = kc / 32;
  const size_t smem = dec_tc_smem(nkb);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MT3_CUDA_CHECK(cudaFuncSetAttribute(dec_gemm_tc_kernel<MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MT3_CUDA_CHECK(cudaFuncSetAttribute(dec_gemm_tc_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  if (a.trace)
    MT3_CUDA_CHECK(launch_kernel_cluster(dec_gemm_tc_kernel<MODE, true>, dim3(cdiv(a.N, kDecBN), 8), dim3(128), smem, s, pdl, 8u,
                                         maps.a, maps.b, a, nkb));
  else
    MT3_CUDA_CHECK(launch_kernel_cluster(dec_gemm_tc_kernel<MODE, false>, dim3(cdiv(a.N, kDecBN), 8), dim3(128), smem, s, pdl, 8u,
                                         maps.a, maps.b, a, nkb));
  MT3_LAUNCH_CHECK();
  return MT3_OK;
}

inline bool dec_gemm_tc_supported(const DecGemmArgs& a) {
  return a.M <= kDecBM && a.N % kDecBN == 0 && a.K % 32 == 0 && a.n_split % 4 == 0 && a.lda % 4 == 0 && a.K <= 8 * 8 * 32;
}

inline int launch_dec_gemm_tc(const DecTcMaps& maps, const DecGemmArgs& a, int mode, cudaStream_t s, bool pdl) {
  return mode == 2 ? launch_dec_gemm_tc_mode<2>(maps, a, s, pdl) : launch_dec_gemm_tc_mode<1>(maps, a, s, pdl);
}

}  // namespace mt3
