// K4: encoder self-attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Reference: dot_product_attention (layers.py:85-157): softmax_k(q . k^T) . v, logits NOT scaled by
// 1/sqrt(d) (layers.py:230-234), mask all ones in the encoder (network.py:283-289).
//
// One CTA = 128 query rows of one (batch, head); T <= 256 keys, head_dim 64.
//   S[128 q, T k]  = Q . K^T   tcgen05.mma kind::tf32, M=128, N=128 per key chunk, K=64 (8 k-steps);
//                               error-compensated (Qhi.Khi + Qhi.Klo + Qlo.Khi) when hi/lo operands are given
//   P = exp(S - rowmax)         two threads per query row read the row from TMEM (tcgen05.ld) and write P (rounded to
//                               tf32, unnormalised) back over S in tensor memory (tcgen05.st): the next MMA reads
//                               its A operand straight from TMEM (kAtPTmem; the older route through 128B-swizzled
//                               shared-memory sub-tiles is kept behind the switch)
//   O[128 q, 64 d] = P . V      B operand = V^T tiles [64 d x 32 keys] (keys contiguous, K-major), written per head
//                               by the QKV GEMM epilogue (TcGemmArgs::VT_*): P . Vhi + P . Vlo; 1/rowsum is
//                               applied in the epilogue.  (An MN-major tf32 B operand returned zeros on sm_100a
//                               in bring-up, so V is transposed once at its producer instead.)
// TMEM: S = 256 columns, O = 64 columns (512 allocated).  Shared memory (192 KB):
//   [ Q hi/lo 64 KB | slot0 64 KB | slot1 64 KB ]   slots hold K chunk c (hi/lo) and later V chunk c (hi/lo);
//   the Q region is recycled as four 32-key P sub-tile buffers once the S MMAs have completed; slot 0 stages O.
// 320 threads: warp 0 TMA, warp 1 MMA issue, warps 2..9 softmax / epilogue (two warps per TMEM lane group).
// Operands come straight from the qkv buffers the QKV GEMM epilogue wrote ([B*T, 3*H*64] hi and lo)
// through two TMA tensor maps; output is [B*T, H*64] (hi/lo split for the out-projection GEMM).
#pragma once

#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc.cuh"

namespace mt3 {

#ifndef MT3_AT_INT_ROUND
#define MT3_AT_INT_ROUND 0
#endif
__device__ __forceinline__ float round_tf32(float x) {
#if MT3_AT_INT_ROUND
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
#else
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
#endif
}

constexpr int kAtQ = 128;                 // query rows per work item
constexpr int kAtKC = 128;                // keys per chunk
constexpr int kAtSub = 128 * 32 * 4;      // one [128 x 32] fp32 sub-tile = 16 KB
constexpr int kVSub = 64 * 32 * 4;        // one [64 d x 32 keys] V^T sub-tile = 8 KB
constexpr int kAtStage = 128 * 64 * 4;    // O staging tile (XOR-swizzled rows of 256 B) = 32 KB
constexpr int kAtSmem = 3 * 4 * kAtSub + kAtStage + 256 + 2 * 2 * 128 * 4;   // 231 680 B of the 232 448 a CTA may have: no
                                                                              // slack, the dynamic window must be 1024-aligned
constexpr int kAtPTmem = 1;               // 1: P stays in tensor memory (overwrites S in place) and is the P.V MMA's A operand;
                                          // 0: P goes through four swizzled shared-memory sub-tile buffers in the Q region
constexpr int kAtThreads = 320;           // warp 0 TMA, warp 1 MMA, warps 2..9 softmax / epilogue (two per TMEM lane group)

// PERSISTENT: one CTA per SM walks over work items (sequence b, head h, 128-query tile) with stride gridDim.x.  All
// mbarriers complete exactly once per item, so the wait parity is (item counter & 1).  While the softmax warps are
// still in item i, the TMA thread already streams item i+1: K chunk 0 as soon as the P.V MMAs that read V chunk 0 are
// done, Q and K chunk 1 as soon as the last P.V MMA is done (the Q region doubles as the P buffers); the MMA thread
// issues S(i+1) as soon as the softmax warps have finished reading S(i) out of TMEM, i.e. under the epilogue of i.
template <bool SPLIT3, bool KSPLIT>
__global__ void __launch_bounds__(kAtThreads, 1)
enc_attention_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                        const __grid_constant__ CUtensorMap tv_hi, const __grid_constant__ CUtensorMap tv_lo, int T, int H, int B,
                        float* __restrict__ out_hi, float* __restrict__ out_lo, int ldo, float* __restrict__ dbg_S,
                        int variant, unsigned long long* __restrict__ dbg_t, int TK_arg, int KS_arg, float* __restrict__ part_o,
                        float2* __restrict__ part_ml) {
  const int KS = KSPLIT ? KS_arg : 1, TK = KSPLIT ? TK_arg : T;       // compile-time 1 / T in the common single-pass form
  // TK = keys per work item (<= 256), KS = key parts per query tile (T = KS * TK).  KS == 1: the item covers all keys and
  // writes the normalised output.  KS > 1 (T = 512, ismir2021): an item covers keys [ks TK, (ks + 1) TK) and writes its
  // UNNORMALISED O tile plus (row max, row sum) to part_o / part_ml; enc_attention_combine_kernel merges the parts
  // (softmax is associative over key blocks: O = sum_k e^(m_k - m) O_k / sum_k e^(m_k - m) l_k).
  // dbg_S (bring-up tool only): raw S rows [B][H][T][T].  dbg_t (tool only): SM-clock stamps of CTA 0's first item ->
  // [0..31] and third item -> [32..63]: MMA thread in slots 0.., first softmax thread in slots 16..
  (void)variant;
  const long long t_start = clock64();
  unsigned long long* tslot = nullptr;
#define AT_STAMP(i) do { if (tslot) tslot[i] = (unsigned long long)(clock64() - t_start); } while (0)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();          // the swizzled tiles need 1024-byte alignment
  uint8_t* sQ = smem;                       // Q hi: 2 sub-tiles, Q lo: 2 sub-tiles (64 KB); later four P sub-tile buffers
  uint8_t* slot[2] = {smem + 4 * kAtSub, smem + 8 * kAtSub};
  float* stage = reinterpret_cast<float*>(smem + 12 * kAtSub);   // O tile on its way to global memory
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 12 * kAtSub + kAtStage);
  uint64_t* q_full = bars;                  // Q landed
  uint64_t* k_full = bars + 1;              // [2] K chunk landed
  uint64_t* v_full = bars + 3;              // [2] V chunk landed
  uint64_t* s_done = bars + 5;              // all S MMAs complete
  uint64_t* p_full = bars + 6;              // [8] P quarter-chunk (32 keys) written (128 arrivals: the 4 warps of one half)
  uint64_t* pv_done = bars + 14;            // [8] PV MMAs of quarter-chunk complete
  uint64_t* s_free = bars + 22;             // softmax warps are done reading S (256 arrivals)
  uint64_t* o_free = bars + 23;             // epilogue has read O out of TMEM (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  float* s_max = reinterpret_cast<float*>(bars + 26);      // [2 halves][128 rows] partial row maxima
  float* s_sum = s_max + 256;                              // [2 halves][128 rows] partial row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Q = H * 64;
  const int nchunk = (TK + kAtKC - 1) / kAtKC;             // 1 or 2
  const int ntile = (T + kAtQ - 1) / kAtQ;
  const int n_items = B * H * ntile * KS;
  const int last_qc = nchunk * 4 - 1;                       // the quarter-chunk whose P.V MMAs are issued last
  constexpr uint32_t kSubBytes = kAtSub;
  const uint32_t lo_tiles = SPLIT3 ? 2u : 1u;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tm_hi);
    tc::prefetch_tmap(&tv_hi);
    if (SPLIT3) {
      tc::prefetch_tmap(&tm_lo);
      tc::prefetch_tmap(&tv_lo);
    }
    tc::mbar_init(q_full, 1);
    for (int c = 0; c < 2; ++c) {
      tc::mbar_init(&k_full[c], 1);
      tc::mbar_init(&v_full[c], 1);
    }
    for (int c = 0; c < 8; ++c) {
      tc::mbar_init(&p_full[c], 128);
      tc::mbar_init(&pv_done[c], 1);
    }
    tc::mbar_init(s_done, 1);
    tc::mbar_init(s_free, 256);
    tc::mbar_init(o_free, 256);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_S = *tmem_slot;            // columns [0, 256)
  const uint32_t tmem_O = tmem_S + 256;          // columns [256, 320)

  if (warp == 0) {
    if (tc::elect_one()) {
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1, pph = ph ^ 1;
        const int ks = item % KS, it2 = item / KS;
        const int qt = it2 % ntile, h = (it2 / ntile) % H, b = it2 / (ntile * H);
        const int row0 = b * T, q0 = qt * kAtQ, koff = ks * TK;
        auto load_q = [&]() {
          tc::mbar_arrive_expect_tx(q_full, 2 * lo_tiles * kSubBytes);
          for (int sub = 0; sub < 2; ++sub) {
            tc::tma_load_2d(sQ + sub * kAtSub, &tm_hi, q_full, h * 64 + sub * 32, row0 + q0);
            if (SPLIT3) tc::tma_load_2d(sQ + (2 + sub) * kAtSub, &tm_lo, q_full, h * 64 + sub * 32, row0 + q0);
          }
        };
        // Q -> Q region.  With P in tensor memory the region is free as soon as the previous item's S MMAs have read Q
        // (s_done, which this thread waited for before it loaded that item's V): request it first.
        if (kAtPTmem) load_q();
        // K chunk 0 -> slot 0: free once the previous item's P.V MMAs over V chunk 0 are done
        if (it > 0) tc::mbar_wait(&pv_done[3], pph);
        tc::mbar_arrive_expect_tx(&k_full[0], 2 * lo_tiles * kSubBytes);
        for (int sub = 0; sub < 2; ++sub) {
          tc::tma_load_2d(slot[0] + sub * kAtSub, &tm_hi, &k_full[0], Q + h * 64 + sub * 32, row0 + koff);
          if (SPLIT3) tc::tma_load_2d(slot[0] + (2 + sub) * kAtSub, &tm_lo, &k_full[0], Q + h * 64 + sub * 32, row0 + koff);
        }
        // Without P in tensor memory the Q region holds the previous item's P buffers until its last P.V MMA is done.
        if (it > 0) tc::mbar_wait(&pv_done[last_qc], pph);
        if (!kAtPTmem) load_q();
        // K chunk 1 -> slot 1: free once the previous item's last P.V MMA (over V chunk 1) is done (waited for above)
        if (nchunk > 1) {
          tc::mbar_arrive_expect_tx(&k_full[1], 2 * lo_tiles * kSubBytes);
          for (int sub = 0; sub < 2; ++sub) {
            tc::tma_load_2d(slot[1] + sub * kAtSub, &tm_hi, &k_full[1], Q + h * 64 + sub * 32, row0 + koff + kAtKC);
            if (SPLIT3) tc::tma_load_2d(slot[1] + (2 + sub) * kAtSub, &tm_lo, &k_full[1], Q + h * 64 + sub * 32, row0 + koff + kAtKC);
          }
        }
        // V^T chunks reuse the slots once every S MMA of THIS item has read K
        tc::mbar_wait(s_done, ph);
        const int vrow = (b * H + h) * 64;
        for (int c = 0; c < nchunk; ++c) {
          tc::mbar_arrive_expect_tx(&v_full[c], 4 * lo_tiles * kVSub);
          for (int sub = 0; sub < 4; ++sub) {
            tc::tma_load_2d(slot[c] + sub * kVSub, &tv_hi, &v_full[c], koff + c * kAtKC + sub * 32, vrow);
            if (SPLIT3) tc::tma_load_2d(slot[c] + (4 + sub) * kVSub, &tv_lo, &v_full[c], koff + c * kAtKC + sub * 32, vrow);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      constexpr uint32_t idesc_s = tc::make_idesc(tc::kFmtTF32, 128, kAtKC, 0, 0);
      constexpr uint32_t idesc_o = tc::make_idesc(tc::kFmtTF32, 128, 64, 0, 0);
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1, pph = ph ^ 1;
        tslot = (dbg_t && blockIdx.x == 0 && (it == 0 || it == 2)) ? dbg_t + (it == 0 ? 0 : 32) : nullptr;
        // ---- S = Q K^T ----
        tc::mbar_wait(q_full, ph);
        AT_STAMP(0);
        if (it > 0) tc::mbar_wait(s_free, pph);              // the softmax warps no longer read the previous S
        tc::tc_fence_after();
        const uint32_t q_addr = tc::smem_u32(sQ);
        for (int c = 0; c < nchunk; ++c) {
          tc::mbar_wait(&k_full[c], ph);
          AT_STAMP(1 + c);
          tc::tc_fence_after();
          const uint32_t k_addr = tc::smem_u32(slot[c]);
          uint32_t acc = 0;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {                   // 64 d = 2 sub-tiles x 4 k-steps of 8
            const uint32_t off = (ks >> 2) * kSubBytes + (ks & 3) * 32;
            const uint64_t a_hi = tc::smem_desc_k_sw128(q_addr + off);
            const uint64_t b_hi = tc::smem_desc_k_sw128(k_addr + off);
            if (SPLIT3) {
              const uint64_t a_lo = tc::smem_desc_k_sw128(q_addr + 2 * kSubBytes + off);
              const uint64_t b_lo = tc::smem_desc_k_sw128(k_addr + 2 * kSubBytes + off);
              tc::mma_tf32(tmem_S + c * kAtKC, a_lo, b_hi, idesc_s, acc);
              tc::mma_tf32(tmem_S + c * kAtKC, a_hi, b_lo, idesc_s, 1u);
              tc::mma_tf32(tmem_S + c * kAtKC, a_hi, b_hi, idesc_s, 1u);
            } else {
              tc::mma_tf32(tmem_S + c * kAtKC, a_hi, b_hi, idesc_s, acc);
            }
            acc = 1u;
          }
        }
        tc::mma_commit(s_done);
        AT_STAMP(3);
        // ---- O = P V ----  quarter-chunks of 32 keys: A = P sub-tile (K-major, one of the 4 sub-tile buffers in the Q
        // region), B = V^T sub-tile [64 d x 32 keys].  The two softmax halves fill their buffers concurrently, so the
        // quarter-chunks become ready in the order 0,2,1,3 (chunk 0) 4,6,5,7 (chunk 1); issue in that order.
        if (it > 0) {
          tc::mbar_wait(o_free, pph);                        // the previous O has been read out of TMEM
          tc::tc_fence_after();
        }
        uint32_t acc = 0;
        const int nq = nchunk * 4;
        for (int i = 0; i < nq; ++i) {
          const int qc = (i & ~3) | ((i & 1) << 1) | ((i >> 1) & 1);      // 0,2,1,3,4,6,5,7
          const int c = qc >> 2, sub = qc & 3;
          if (i == (c << 2)) {
            tc::mbar_wait(&v_full[c], ph);
            AT_STAMP(4 + 3 * c);
          }
          tc::mbar_wait(&p_full[qc], ph);
          if (i == (c << 2)) AT_STAMP(5 + 3 * c);
          tc::tc_fence_after();
          const uint32_t p_addr = tc::smem_u32(sQ) + sub * kSubBytes;
          const uint32_t p_tmem = tmem_S + (uint32_t)(qc * 32);            // P quarter-chunk: columns [32 qc, +32), all 128 lanes
          const uint32_t v_addr = tc::smem_u32(slot[c]) + sub * kVSub;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {                   // 32 keys = 4 k-steps of 8
            if (kAtPTmem) {
              tc::mma_tf32_ts(tmem_O, p_tmem + ks * 8, tc::smem_desc_k_sw128(v_addr + ks * 32), idesc_o, acc);
              acc = 1u;
              if (SPLIT3) tc::mma_tf32_ts(tmem_O, p_tmem + ks * 8, tc::smem_desc_k_sw128(v_addr + 4 * kVSub + ks * 32), idesc_o, 1u);
            } else {
              const uint64_t a = tc::smem_desc_k_sw128(p_addr + ks * 32);
              tc::mma_tf32(tmem_O, a, tc::smem_desc_k_sw128(v_addr + ks * 32), idesc_o, acc);
              acc = 1u;
              if (SPLIT3) tc::mma_tf32(tmem_O, a, tc::smem_desc_k_sw128(v_addr + 4 * kVSub + ks * 32), idesc_o, 1u);
            }
          }
          tc::mma_commit(&pv_done[qc]);
          if ((i & 3) == 3) AT_STAMP(6 + 3 * c);
        }
      }
    }
  } else {
    // ---- softmax + epilogue: 8 warps; warp w reads TMEM lanes [32 (w % 4), +32), the two warps of a lane group split
    // the columns: of every 128-key chunk half 0 takes keys [0, 64), half 1 keys [64, 128); of O 32 columns each ----
    const int wq = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = wq * 32 + lane;                            // row inside the tile
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const int st = (int)threadIdx.x - 64;                    // 0..255
    int it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const uint32_t ph = it & 1;
      const int ks = item % KS, it2 = item / KS;
      const int qt = it2 % ntile, h = (it2 / ntile) % H, b = it2 / (ntile * H);
      const int row0 = b * T, q0 = qt * kAtQ;
      tslot = (dbg_t && blockIdx.x == 0 && threadIdx.x == 64 && (it == 0 || it == 2)) ? dbg_t + (it == 0 ? 0 : 32) : nullptr;
      tc::mbar_wait(s_done, ph);
      tc::tc_fence_after();
      AT_STAMP(16);
      float mx = -INFINITY;
      for (int c = 0; c < nchunk; ++c) {
        const int c0 = c * kAtKC + half * 64;                // this half's 64 columns of the chunk: both loads in flight
        if (c0 >= TK) continue;
        uint32_t v[32], w[32];
        tc::tmem_ld_32x32(tmem_S + lane_base + c0, v);
        tc::tmem_ld_32x32(tmem_S + lane_base + c0 + 32, w);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (c0 + j < TK) mx = fmaxf(mx, __uint_as_float(v[j]));
          if (c0 + 32 + j < TK) mx = fmaxf(mx, __uint_as_float(w[j]));
        }
        if (dbg_S && KS == 1 && q0 + r < T) {
          float* drow = dbg_S + (((long long)b * H + h) * T + q0 + r) * T + c0;
          for (int j = 0; j < 32; ++j) {
            if (c0 + j < T) drow[j] = __uint_as_float(v[j]);
            if (c0 + 32 + j < T) drow[32 + j] = __uint_as_float(w[j]);
          }
        }
      }
      s_max[half * 128 + r] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(s_max[r], s_max[128 + r]);
      AT_STAMP(17);
      float sum = 0.f;
      for (int c = 0; c < nchunk; ++c) {
        AT_STAMP(18 + 2 * c);
        for (int sub = half * 2; sub < half * 2 + 2; ++sub) {  // this half's two quarter-chunks of chunk c -> P buffers `sub`
          const int qc = c * 4 + sub;
          const int c0 = c * kAtKC + sub * 32;
          if (c > 0 && !kAtPTmem) {                          // the buffer was read by quarter-chunk qc - 4's MMAs
            tc::mbar_wait(&pv_done[qc - 4], ph);
            tc::tc_fence_after();
          }
          uint32_t v[32];
          tc::tmem_ld_32x32(tmem_S + lane_base + c0, v);
          tc::tmem_ld_wait();
          float p[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            // ex2.approx (2 ulp) is ample: the value is rounded to tf32 (2^-11) on the next line
            float e = (c0 + j < TK) ? exp2f((__uint_as_float(v[j]) - mx) * 1.4426950408889634f) : 0.f;
            e = round_tf32(e);                              // exactly what the tensor core will read
            p[j] = e;
            sum += e;
          }
          if (kAtPTmem) {
            // P overwrites S in place: same lanes, same 32 columns; the P.V MMA reads it as its A operand
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(p[j]);
            tc::tmem_st_32x32(tmem_S + lane_base + c0, v);
            tc::tmem_st_wait();
            tc::tc_fence_before();
          } else {
            // P sub-tile buffer `sub`, row r, 8 x 16-byte pieces XOR-swizzled by (r % 8)
            uint8_t* rowp = sQ + sub * kAtSub + r * 128;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              *reinterpret_cast<float4*>(rowp + ((j4 ^ (r & 7)) << 4)) = make_float4(p[4 * j4], p[4 * j4 + 1], p[4 * j4 + 2], p[4 * j4 + 3]);
            tc::fence_proxy_async();                         // generic-proxy writes -> visible to the MMA (async proxy)
          }
          tc::mbar_arrive(&p_full[qc]);
        }
        AT_STAMP(19 + 2 * c);
      }
      tc::tc_fence_before();
      tc::mbar_arrive(s_free);                               // this thread is done reading S: the next S may be issued
      s_sum[half * 128 + r] = sum;
      tc::mbar_wait(&pv_done[last_qc], ph);                  // issued last: covers every earlier MMA of the item
      tc::tc_fence_after();
      AT_STAMP(22);
      uint32_t v[32];
      tc::tmem_ld_32x32(tmem_O + lane_base + half * 32, v);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(o_free);                               // O is in registers: the next item may overwrite it
      asm volatile("bar.sync 1, 256;" ::: "memory");         // partial sums visible; previous item's staging reads done
      const float rsum = s_sum[r] + s_sum[128 + r];
      const float inv = KS == 1 ? 1.0f / rsum : 1.0f;       // key parts stay unnormalised: the combine kernel divides
      if (KS > 1 && half == 0 && q0 + r < T)
        part_ml[(((long long)ks * B + b) * H + h) * T + q0 + r] = make_float2(mx, rsum);
      // O row -> staging tile (rows of 256 B, 16-byte chunk c4 of row r stored at chunk c4 ^ (r & 7)), then coalesced stores
      {
        uint8_t* srow = reinterpret_cast<uint8_t*>(stage) + r * 256;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const int c4 = half * 8 + j4;
          *reinterpret_cast<float4*>(srow + ((c4 ^ (r & 7)) << 4)) =
              make_float4(__uint_as_float(v[4 * j4]) * inv, __uint_as_float(v[4 * j4 + 1]) * inv,
                          __uint_as_float(v[4 * j4 + 2]) * inv, __uint_as_float(v[4 * j4 + 3]) * inv);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
      for (int i8 = 0; i8 < 8; ++i8) {
        const int idx = i8 * 256 + st;
        const int row = idx >> 4, c4 = idx & 15;
        const int q = q0 + row;
        if (q < T) {
          const float4 o = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(stage) + row * 256 + ((c4 ^ (row & 7)) << 4));
          const long long off = (long long)(row0 + q) * ldo + h * 64 + c4 * 4;
          if (KS > 1) {
            *reinterpret_cast<float4*>(part_o + (long long)ks * B * T * ldo + off) = o;
          } else if (out_lo) {
            float4 oh, ol;
            split_tf32(o.x, oh.x, ol.x); split_tf32(o.y, oh.y, ol.y); split_tf32(o.z, oh.z, ol.z); split_tf32(o.w, oh.w, ol.w);
            *reinterpret_cast<float4*>(out_hi + off) = oh;
            *reinterpret_cast<float4*>(out_lo + off) = ol;
          } else {
            *reinterpret_cast<float4*>(out_hi + off) = o;
          }
        }
      }
      AT_STAMP(23);
    }
  }
#undef AT_STAMP
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_S, 512);
}

// Merge of KS key parts: out[row, h, :] = sum_k e^(m_k - m) O_k / sum_k e^(m_k - m) l_k.  One warp per (row, head).
__global__ void enc_attention_combine_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml, int B, int T, int H,
                                             int KS, float* __restrict__ out_hi, float* __restrict__ out_lo) {
  const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // (b, t, h)
  const int lane = threadIdx.x & 31;
  if (item >= (long long)B * T * H) return;
  const int h = (int)(item % H);
  const long long row = item / H;                      // b * T + t
  const int b = (int)(row / T), t = (int)(row % T);
  const int ldo = H * 64;
  float m = -INFINITY;
  for (int k = 0; k < KS; ++k) m = fmaxf(m, part_ml[(((long long)k * B + b) * H + h) * T + t].x);
  float den = 0.f;
  float2 acc = make_float2(0.f, 0.f);
  for (int k = 0; k < KS; ++k) {
    const float2 ml = part_ml[(((long long)k * B + b) * H + h) * T + t];
    const float w = expf(ml.x - m);
    den = fmaf(w, ml.y, den);
    const float2 o = *reinterpret_cast<const float2*>(part_o + (long long)k * B * T * ldo + row * ldo + h * 64 + lane * 2);
    acc.x = fmaf(w, o.x, acc.x);
    acc.y = fmaf(w, o.y, acc.y);
  }
  const float inv = 1.0f / den;
  const float x = acc.x * inv, y = acc.y * inv;
  const long long off = row * ldo + h * 64 + lane * 2;
  if (out_lo) {
    float xh, xl, yh, yl;
    split_tf32(x, xh, xl);
    split_tf32(y, yh, yl);
    *reinterpret_cast<float2*>(out_hi + off) = make_float2(xh, yh);
    *reinterpret_cast<float2*>(out_lo + off) = make_float2(xl, yl);
  } else {
    *reinterpret_cast<float2*>(out_hi + off) = make_float2(x, y);
  }
}

// floats of scratch the key-split form needs (T > 256): KS unnormalised O tiles + KS (max, sum) pairs per (row, head)
inline int64_t enc_attention_tc_scratch_floats(int B, int T, int H) {
  if (T <= 2 * kAtKC) return 0;
  const int64_t KS = T / (2 * kAtKC);
  return KS * (int64_t)B * T * H * 64 + KS * (int64_t)B * H * T * 2;
}
inline bool enc_attention_tc_supported(int T) { return T % 8 == 0 && (T <= 2 * kAtKC || T % (2 * kAtKC) == 0); }

// qkv_hi / qkv_lo: [B*T, 3*H*64] (q and k are read); vt: V^T [B*H*64, T] (box rows 64); out: [B*T, H*64]
// (out_lo optional).  T <= 256 (T % 8 == 0), or a multiple of 256 (ismir2021: 512) with `scratch`
// (enc_attention_tc_scratch_floats) for the key-split form.
inline int launch_enc_attention_tc(const TcOperand& qkv, const TcOperand& vt, int B, int T, int H, float* out_hi, float* out_lo,
                                   bool split3, cudaStream_t s, float* dbg_S = nullptr, int variant = 0,
                                   unsigned long long* dbg_t = nullptr, float* scratch = nullptr) {
  MT3_REQUIRE(enc_attention_tc_supported(T), MT3_ERR_UNSUPPORTED, "tc attention: T=%d (needs T <= 256 and a multiple of 8, or a multiple of 256)", T);
  MT3_REQUIRE(!split3 || qkv.has_lo, MT3_ERR_BAD_ARG, "tc attention: TF32X3 needs hi/lo qkv");
  const int KS = T <= 2 * kAtKC ? 1 : T / (2 * kAtKC);
  const int TK = T / KS;
  MT3_REQUIRE(KS == 1 || scratch != nullptr, MT3_ERR_WORKSPACE, "tc attention: T=%d needs the key-split scratch buffer", T);
  float* part_o = scratch;
  float2* part_ml = KS > 1 ? reinterpret_cast<float2*>(scratch + (int64_t)KS * B * T * H * 64) : nullptr;
  static bool attr_done = false;
  if (!attr_done) {
    MT3_CUDA_CHECK(cudaFuncSetAttribute(enc_attention_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    MT3_CUDA_CHECK(cudaFuncSetAttribute(enc_attention_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    MT3_CUDA_CHECK(cudaFuncSetAttribute(enc_attention_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    MT3_CUDA_CHECK(cudaFuncSetAttribute(enc_attention_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
    attr_done = true;
  }
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    MT3_CUDA_CHECK(cudaGetDevice(&dev));
    MT3_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const int n_items = B * H * cdiv(T, kAtQ) * KS;
  dim3 grid(n_items < sm_count ? n_items : sm_count);       // one persistent CTA per SM (198 KB of shared memory each)
#define MT3_AT_LAUNCH(S3, KSP, LO_Q, LO_V)                                                                                       \
  enc_attention_tc_kernel<S3, KSP><<<grid, kAtThreads, kAtSmem, s>>>(qkv.hi, LO_Q, vt.hi, LO_V, T, H, B, out_hi, out_lo, H * 64, dbg_S, variant, \
                                                                     dbg_t, TK, KS, part_o, part_ml)
  if (split3 && KS > 1) MT3_AT_LAUNCH(true, true, qkv.lo, vt.lo);
  else if (split3) MT3_AT_LAUNCH(true, false, qkv.lo, vt.lo);
  else if (KS > 1) MT3_AT_LAUNCH(false, true, qkv.hi, vt.hi);
  else MT3_AT_LAUNCH(false, false, qkv.hi, vt.hi);
#undef MT3_AT_LAUNCH
  MT3_LAUNCH_CHECK();
  if (KS > 1) {
    const long long items = (long long)B * T * H;
    enc_attention_combine_kernel<<<(unsigned)((items + 7) / 8), 256, 0, s>>>(part_o, part_ml, B, T, H, KS, out_hi, out_lo);
    MT3_LAUNCH_CHECK();
  }
  return MT3_OK;
}

}  // namespace mt3
