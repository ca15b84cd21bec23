// tcgen05 / TMA GEMM for the MT3 encoder-sized matmuls (M = B*T rows), sm_100a.
//
//   C[M,N] = epilogue( row_scale[m] * (A[M,K] . W[K,N]) ),   A row-major (K contiguous),
//   W given TRANSPOSED as Wt[N,K] (K contiguous): both operands are "K-major" UMMA operands.
//
// Precision (reference computes in float32, gin/model.gin:50):
//   MT3_GEMM_TF32    one tcgen05.mma kind::tf32 pass (10-bit mantissa)
//   MT3_GEMM_TF32X3  error-compensated: x = hi + lo with hi = x truncated to tf32 (exactly what the
//                    tensor core reads), lo = x - hi (exact in fp32);  A.B ~= Ahi.Bhi + Ahi.Blo + Alo.Bhi,
//                    all three accumulated in the same fp32 TMEM accumulator.  Dropped term ~2^-22.
//                    Operands arrive pre-split (weights at model creation, activations by the
//                    producing kernel's epilogue), so the main loop is pure TMA -> tcgen05.
//
// Structure (one 128 x 128 output tile per CTA, 192 threads):
//   warp 0      TMA producer: per k-block (32 fp32 = one 128-byte swizzle row) loads the A and B tiles
//               (hi and lo) into a STAGES-deep shared-memory ring, completion on full[stage]
//   warp 1      MMA issuer: one elected thread issues 4 (x3) tcgen05.mma 128x128x8 per k-block into a
//               128-column fp32 TMEM accumulator; tcgen05.commit releases empty[stage]; allocs/frees TMEM
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> fused epilogue -> global
// Epilogues are the same set as the SIMT GEMM (gemm_simt.cuh): RMSNorm row scale, residual add,
// sinusoid add, gated-GELU, split store (KV-cache append), plus optional hi/lo split of the output.
#pragma once

#include <cuda_fp16.h>

#include "common.cuh"
#include "gemm_simt.cuh"   // EPI_* enums, gelu_tanh
#include "tc.cuh"

namespace mt3 {

struct TcGemmArgs {
  int M, N, K;
  const float* row_scale;
  int epi;
  const float* R_hi; const float* R_lo; int ldr;
  const float* pe; int pe_T; int pe_ld;
  float* C_hi; float* C_lo; int ldc;       // C_lo != null: write hi/lo split of the result
  int n_split; void* C1; int kv_fmt; int hm_rows_per_b; int hm_cap; int hm_H; const int* hm_pos;   // head-major K/V store, see kv_dest()
  // Transposed per-head store for columns >= vt_col0 (the V third of a fused QKV projection), enabled by VT_hi:
  // vt[((b*H + h)*64 + d)*vt_T + t] with row m = b*vt_T + t, column = vt_col0 + h*64 + d.  This is the K-major
  // (keys contiguous) operand the tcgen05 attention kernel reads for P.V.
  float* VT_hi; float* VT_lo; int vt_col0; int vt_T; int vt_H;
};

constexpr int kTcBM = 128, kTcBN = 128, kTcBK = 32;
constexpr int kTcTileBytes = kTcBM * kTcBK * 4;   // 16 KB (A and B tiles have the same size)

template <bool SPLIT3>
struct TcGemmCfg {
  static constexpr int kTilesPerStage = SPLIT3 ? 4 : 2;
  static constexpr int kStageBytes = kTilesPerStage * kTcTileBytes;
  static constexpr int kStages = SPLIT3 ? 3 : 6;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <bool SPLIT3>
__global__ void __launch_bounds__(192, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                 const TcGemmArgs p) {
  using Cfg = TcGemmCfg<SPLIT3>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tmem_full = empty + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kTcBM, n0 = blockIdx.x * kTcBN;
  const int nk = p.K / kTcBK;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA_hi);
    tc::prefetch_tmap(&tmB_hi);
    if (SPLIT3) {
      tc::prefetch_tmap(&tmA_lo);
      tc::prefetch_tmap(&tmB_lo);
    }
    for (int s = 0; s < Cfg::kStages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, kTcBN);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (tc::elect_one()) {
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        tc::mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + s * Cfg::kStageBytes;
        tc::mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
        tc::tma_load_2d(st, &tmA_hi, &full[s], kb * kTcBK, m0);
        tc::tma_load_2d(st + kTcTileBytes, &tmB_hi, &full[s], kb * kTcBK, n0);
        if (SPLIT3) {
          tc::tma_load_2d(st + 2 * kTcTileBytes, &tmA_lo, &full[s], kb * kTcBK, m0);
          tc::tma_load_2d(st + 3 * kTcTileBytes, &tmB_lo, &full[s], kb * kTcBK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc = tc::make_idesc(tc::kFmtTF32, kTcBM, kTcBN, 0, 0);
      for (int kb = 0; kb < nk; ++kb) {
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        const uint32_t st = tc::smem_u32(smem + s * Cfg::kStageBytes);
        const uint64_t a_hi = tc::smem_desc_k_sw128(st);
        const uint64_t b_hi = tc::smem_desc_k_sw128(st + kTcTileBytes);
        const uint64_t a_lo = tc::smem_desc_k_sw128(st + 2 * kTcTileBytes);
        const uint64_t b_lo = tc::smem_desc_k_sw128(st + 3 * kTcTileBytes);
#pragma unroll
        for (int k = 0; k < kTcBK / 8; ++k) {
          const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);   // 8 tf32 = 32 bytes along K inside the swizzle row
          const uint32_t acc = (kb | k) != 0;
          if (SPLIT3) {
            tc::mma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, acc);
            tc::mma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
            tc::mma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, 1u);
          } else {
            tc::mma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, acc);
          }
        }
        tc::mma_commit(&empty[s]);     // smem slot free once these MMAs have read it
      }
      tc::mma_commit(tmem_full);       // accumulator complete
    }
  } else {
    // ---- epilogue: warp (2..5) owns TMEM lanes [32*(warp%4), +32) ----
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
    const int wq = warp & 3;
    const int m = m0 + wq * 32 + lane;
    const bool row_ok = m < p.M;
    const float rs = (row_ok && p.row_scale) ? p.row_scale[m] : 1.f;
#pragma unroll 1
    for (int c0 = 0; c0 < kTcBN; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld_32x32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, r);
      tc::tmem_ld_wait();
      if (!row_ok) continue;
      const int n = n0 + c0;
      if (n >= p.N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * rs;
      if (p.VT_hi && n >= p.vt_col0) {        // lanes are consecutive rows (t): each store below is one coalesced line
        const int bb = m / p.vt_T, tt = m % p.vt_T;
        const int hh = (n - p.vt_col0) >> 6, d0 = (n - p.vt_col0) & 63;
        const long long base = (((long long)bb * p.vt_H + hh) * 64 + d0) * p.vt_T + tt;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (p.VT_lo) {
            float hi, lo;
            split_tf32(v[j], hi, lo);
            p.VT_hi[base + (long long)j * p.vt_T] = hi;
            p.VT_lo[base + (long long)j * p.vt_T] = lo;
          } else {
            p.VT_hi[base + (long long)j * p.vt_T] = v[j];
          }
        }
        continue;
      }
      if (p.epi == EPI_GATED_GELU) {
        float* ch = p.C_hi + (long long)m * p.ldc + (n >> 1);
        float* cl = p.C_lo ? p.C_lo + (long long)m * p.ldc + (n >> 1) : nullptr;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float o[4], oh[4], ol[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            o[j] = gelu_tanh(v[8 * q + 2 * j]) * v[8 * q + 2 * j + 1];
            split_tf32(o[j], oh[j], ol[j]);
          }
          if (cl) {
            *reinterpret_cast<float4*>(ch + 4 * q) = make_float4(oh[0], oh[1], oh[2], oh[3]);
            *reinterpret_cast<float4*>(cl + 4 * q) = make_float4(ol[0], ol[1], ol[2], ol[3]);
          } else {
            *reinterpret_cast<float4*>(ch + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        continue;
      }
      if (p.epi == EPI_RESIDUAL) {
        const float4* rh = reinterpret_cast<const float4*>(p.R_hi + (long long)m * p.ldr + n);
        const float4* rl = p.R_lo ? reinterpret_cast<const float4*>(p.R_lo + (long long)m * p.ldr + n) : nullptr;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 a = rh[q];
          if (rl) {
            const float4 b = rl[q];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
          }
          v[4 * q] += a.x; v[4 * q + 1] += a.y; v[4 * q + 2] += a.z; v[4 * q + 3] += a.w;
        }
      } else if (p.epi == EPI_ADD_PE) {
        const float4* pp = reinterpret_cast<const float4*>(p.pe + (long long)(m % p.pe_T) * p.pe_ld + n);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 a = __ldg(pp + q);
          v[4 * q] += a.x; v[4 * q + 1] += a.y; v[4 * q + 2] += a.z; v[4 * q + 3] += a.w;
        }
      }
      float* ch;
      float* cl = nullptr;
      if (n < p.n_split) {
        ch = p.C_hi + (long long)m * p.ldc + n;
        if (p.C_lo) cl = p.C_lo + (long long)m * p.ldc + n;
      } else {
        const int pos = p.hm_pos ? *p.hm_pos : 0;
        int d;
        char* row = reinterpret_cast<char*>(p.C1) + kv_dest(m, n - p.n_split, p.hm_rows_per_b, p.hm_cap, p.hm_H, pos, p.kv_fmt, d);
        if (p.kv_fmt == 1) {                 // 32 consecutive columns of one head: 64 contiguous bytes of fp16
          uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(row) + d);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            __half2 h0 = __floats2half2_rn(v[8 * q], v[8 * q + 1]), h1 = __floats2half2_rn(v[8 * q + 2], v[8 * q + 3]);
            __half2 h2 = __floats2half2_rn(v[8 * q + 4], v[8 * q + 5]), h3 = __floats2half2_rn(v[8 * q + 6], v[8 * q + 7]);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
            dst[q] = u;
          }
          continue;
        }
        if (p.kv_fmt == 2) {
#pragma unroll
          for (int q = 0; q < 8; ++q) kv_store<4>(row, 2, d + 4 * q, v + 4 * q);
          continue;
        }
        ch = reinterpret_cast<float*>(row) + d;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (cl) {
          float h0, h1, h2, h3, l0, l1, l2, l3;
          split_tf32(v[4 * q], h0, l0); split_tf32(v[4 * q + 1], h1, l1);
          split_tf32(v[4 * q + 2], h2, l2); split_tf32(v[4 * q + 3], h3, l3);
          *reinterpret_cast<float4*>(ch + 4 * q) = make_float4(h0, h1, h2, h3);
          *reinterpret_cast<float4*>(cl + 4 * q) = make_float4(l0, l1, l2, l3);
        } else {
          *reinterpret_cast<float4*>(ch + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, kTcBN);
}

// ---- host side ----------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// 2-D fp32 tensor [rows, cols] with row pitch ld (elements); box = [box_rows, 32 cols] (32 fp32 = 128 B),
// 128-byte swizzle, out-of-bounds reads return zeros.
inline int make_tmap_2d(CUtensorMap* out, const float* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  PFN_tmapEncodeTiled fn = tmap_encode_fn();
  MT3_REQUIRE(fn, MT3_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  MT3_REQUIRE(((uintptr_t)ptr & 15) == 0 && (ld * 4) % 16 == 0, MT3_ERR_BAD_ARG, "TMA operand must be 16-byte aligned");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)kTcBK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MT3_REQUIRE(r == CUDA_SUCCESS, MT3_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return MT3_OK;
}

struct TcOperand {           // an [rows, K] K-major operand, optionally hi/lo split
  CUtensorMap hi, lo;
  bool has_lo = false;
};

inline int make_operand(TcOperand* op, const float* hi, const float* lo, uint64_t rows, uint64_t K, uint64_t ld,
                        uint32_t box_rows = 128) {
  int r = make_tmap_2d(&op->hi, hi, rows, K, ld, box_rows);
  if (r != MT3_OK) return r;
  op->has_lo = lo != nullptr;
  if (lo) return make_tmap_2d(&op->lo, lo, rows, K, ld, box_rows);
  op->lo = op->hi;
  return MT3_OK;
}

inline int launch_tc_gemm(const TcOperand& A, const TcOperand& B, const TcGemmArgs& a, bool split3, cudaStream_t s) {
  MT3_REQUIRE(a.K % kTcBK == 0 && a.N % 32 == 0 && a.n_split % 32 == 0 && a.ldc % 4 == 0, MT3_ERR_UNSUPPORTED,
              "tc gemm: K=%d must be a multiple of 32, N=%d / n_split=%d of 32", a.K, a.N, a.n_split);
  MT3_REQUIRE(!split3 || (A.has_lo && B.has_lo), MT3_ERR_BAD_ARG, "tc gemm: TF32X3 needs hi/lo operands");
  static bool attr_done = false;
  if (!attr_done) {
    MT3_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        TcGemmCfg<true>::kSmemBytes));
    MT3_CUDA_CHECK(cudaFuncSetAttribute(gemm_tf32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        TcGemmCfg<false>::kSmemBytes));
    attr_done = true;
  }
  dim3 grid(cdiv(a.N, kTcBN), cdiv(a.M, kTcBM));
  if (split3)
    gemm_tf32_kernel<true><<<grid, 192, TcGemmCfg<true>::kSmemBytes, s>>>(A.hi, A.lo, B.hi, B.lo, a);
  else
    gemm_tf32_kernel<false><<<grid, 192, TcGemmCfg<false>::kSmemBytes, s>>>(A.hi, A.hi, B.hi, B.hi, a);
  MT3_LAUNCH_CHECK();
  return MT3_OK;
}

}  // namespace mt3
