// Decode-step kernels (one new token per sequence, M = batch rows), sm_100a.
//
// The decode step is bandwidth/latency bound (DESIGN.md "decode step"): per layer it streams 11.8 MB of fp32 weights
// (L2-resident across steps) and, per sequence, the self-attention K/V rows written so far plus the 256 hoisted
// cross-attention K/V rows (HBM).  The kernels:
//
//  sgemm_dec_cluster_kernel   exact-fp32 GEMM for M <= 64 rows, the default.  One thread-block CLUSTER (8 CTAs; 16 for
//                      the long-K MLP-out projection) per 64 x 32 output tile splits K; every CTA pushes the rows
//                      owned by rank r into rank r's shared memory and, after one cluster barrier, each rank sums its
//                      rows in rank order (deterministic) and runs the fused epilogue: RMSNorm factor (statistic
//                      computed from the A tiles the GEMM loads anyway, layers.py:613-616), residual, gated GELU,
//                      head-major KV-cache append, per-tile sums of squares of the output.
//  sgemm_dec_cluster2_kernel  two such GEMMs that read the same inputs in one launch (out-projection + the
//                      cross-attention query projection through a precomposed weight block).
//  sgemm_dec_kernel    the same GEMM without clusters (MT3_DEC_CLUSTER=0 / shapes the cluster kernel does not
//                      tile): K split over gridDim.y CTAs, partial tiles through an L2-resident scratch, the LAST
//                      CTA to arrive for a tile (atomic ticket) sums them in fixed order.
//  dec_attention_bulk_kernel  one query per (sequence, head) over contiguous head-major K/V rows
//                      (kv[b][K|V][head][cap][64]).  A producer warp streams 32-key tiles (8 KB) with
//                      cp.async.bulk (TMA 1-D) into a 6-stage shared-memory ring, signalled through
//                      mbarriers; four consumer warps compute scores (pass 1: K tiles), the softmax,
//                      and P.V (pass 2: V tiles).  Slots >= len are never read -- identical to the
//                      reference's -1e10 mask bias (layers.py:297-322: exp underflows to exactly 0).
#pragma once

#include "common.cuh"
#include "gemm_simt.cuh"
#include "tc.cuh"

namespace mt3 {

struct DecGemmArgs {
  const float* A; int lda;        // [M, K]
  const float* W; int ldw;        // [K, N] row-major
  int M, N, K;
  int norm;                       // 1: scale row m by rsqrt(mean(A[m,:]^2) + eps); needs K == row length
  float eps;
  int epi;
  const float* R; int ldr;
  float* C; int ldc;
  int n_split; float* C1; int hm_rows_per_b; int hm_cap; int hm_H; const int* hm_pos;
  float* partial;                 // [splits][n_tiles][64*32 + 64] scratch
  int* counters;                  // [n_tiles], zero on entry, left zero on exit
  // optional second activation source: columns [K0, K) of the virtual A come from A2[:, k - K0] (fused launches
  // multiply the concatenation [o | y] by a precomposed weight block); A2 == nullptr -> single source
  const float* A2; int lda2; int K0;
  // optional: per (row, 32-column tile) sum of squares of the OUTPUT row slice, [M][ssq_ld]; the consumer of the
  // output sums the N/32 partials in order and gets the RMSNorm statistic without re-reading the row
  float* ssq_out; int ssq_ld;
  unsigned long long* trace;      // debug timeline slot (mt3_debug_trace_step) or null: [0] min start, [1] max end (ns,
                                  // %globaltimer); [2..6] clock64 deltas of CTA (0,0) at its phase boundaries
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Volatile PTX loads: emitted in program order, so a run of them stays a run of independent loads in
// flight (the compiler otherwise pairs each load with its shared-memory store and serialises the latency).
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {      // read-once weights: no L1 allocation
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
#define MT3_COMPILER_BARRIER() asm volatile("" ::: "memory")

constexpr int kDecBM = 64, kDecBN = 32, kDecKC = 64;     // output tile 64 x 32, K chunk per CTA
constexpr int kDecBK = kDecKC;                               // (host-side divisibility checks)
constexpr int kDecTileFloats = kDecBM * kDecBN + kDecBM;   // partial tile + per-row sum of squares

// Each CTA handles ONE K chunk of 64: all of its global loads (8 KB of weights, 16 KB of activations)
// are issued before anything is consumed -- one memory latency per CTA instead of one per k-tile --
// then the 64 k-steps run out of shared memory with no further synchronisation.
__global__ void __launch_bounds__(128)
sgemm_dec_kernel(const DecGemmArgs p) {
  constexpr int BM = kDecBM, BN = kDecBN, KC = kDecKC, NT = 128, APAD = 4;
  __shared__ __align__(16) float As[KC][BM + APAD];      // transposed: [k][m]
  __shared__ __align__(16) float Bs[KC][BN];
  __shared__ float s_ss[BM];
  __shared__ int s_last;

  const int tid = threadIdx.x;
  const int tx = tid % 8, ty = tid / 8;                  // thread tile: rows ty*4.., cols tx*4..
  const int n0 = blockIdx.x * BN;
  const int splits = gridDim.y, ks = blockIdx.y;
  const int kbeg = ks * KC;                              // host guarantees K == splits * KC

  // ---- load phase: 4 weight + 8 activation 16-byte loads per thread, all in flight together ----
  float4 rb[4], ra[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * NT;                        // 512 float4: [64 k][8 column quads]
    const int kr = idx >> 3, nq = idx & 7;
    rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + nq * 4 < p.N) rb[i] = __ldg(reinterpret_cast<const float4*>(p.W + (long long)(kbeg + kr) * p.ldw + n0 + nq * 4));
  }
  // weights do not depend on the previous kernel: they are in flight while it drains (PDL)
  pdl_wait();
  pdl_trigger();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = tid + i * NT;                        // 1024 float4: [64 m][16 k quads]
    const int row = idx >> 4, kq = idx & 15;
    ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < p.M) ra[i] = *reinterpret_cast<const float4*>(p.A + (long long)row * p.lda + kbeg + kq * 4);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * NT;
    *reinterpret_cast<float4*>(&Bs[idx >> 3][(idx & 7) * 4]) = rb[i];
  }
  // rows handled by this thread in the A load: (tid >> 4) + 8 i ; 16 consecutive lanes share a row
  float ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int idx = tid + i * NT;
    const int row = idx >> 4, kq = idx & 15;
    As[kq * 4 + 0][row] = ra[i].x;
    As[kq * 4 + 1][row] = ra[i].y;
    As[kq * 4 + 2][row] = ra[i].z;
    As[kq * 4 + 3][row] = ra[i].w;
    float v = ra[i].x * ra[i].x + ra[i].y * ra[i].y + ra[i].z * ra[i].z + ra[i].w * ra[i].w;
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    ss[i] = v;                                           // sum of squares of row (tid>>4) + 8 i over this K chunk
  }
  __syncthreads();

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 16
  for (int k = 0; k < KC; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
    const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }

  const int n_tiles = gridDim.x;
  if (splits > 1) {
    float* mine = p.partial + ((long long)ks * n_tiles + blockIdx.x) * kDecTileFloats;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(mine + (ty * 4 + i) * BN + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if ((tid & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) mine[BM * BN + (tid >> 4) + 8 * i] = ss[i];
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int ticket = atomicAdd(&p.counters[blockIdx.x], 1);
      s_last = (ticket == splits - 1);
      if (s_last) p.counters[blockIdx.x] = 0;            // re-arm for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // fixed-order reduction of all partials (including our own) -> bit-reproducible
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float sst = 0.f;
    for (int s0 = 0; s0 < splits; s0 += 4) {
      float4 v[4][4];
      float sv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {                      // 4 partial tiles in flight per round
        const int s = s0 + u;
        const float* src = p.partial + ((long long)min(s, splits - 1) * n_tiles + blockIdx.x) * kDecTileFloats;
#pragma unroll
        for (int i = 0; i < 4; ++i) v[u][i] = __ldcg(reinterpret_cast<const float4*>(src + (ty * 4 + i) * BN + tx * 4));
        sv[u] = (tid < BM) ? __ldcg(src + BM * BN + tid) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (s0 + u < splits) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i][0] += v[u][i].x; acc[i][1] += v[u][i].y; acc[i][2] += v[u][i].z; acc[i][3] += v[u][i].w;
          }
          sst += sv[u];
        }
      }
    }
    if (tid < BM) s_ss[tid] = sst;
  } else {
    if ((tid & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s_ss[(tid >> 4) + 8 * i] = ss[i];
    }
  }
  __syncthreads();

  // ---- fused epilogue (same set as gemm_simt.cuh) ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = ty * 4 + i;
    if (m >= p.M) continue;
    const int n = n0 + tx * 4;
    if (n >= p.N) continue;
    const float rs = p.norm ? 1.0f / sqrtf(s_ss[m] / (float)p.K + p.eps) : 1.f;
    float4 v = make_float4(acc[i][0] * rs, acc[i][1] * rs, acc[i][2] * rs, acc[i][3] * rs);
    if (p.epi == EPI_GATED_GELU) {
      *reinterpret_cast<float2*>(p.C + (long long)m * p.ldc + (n >> 1)) = make_float2(gelu_tanh(v.x) * v.y, gelu_tanh(v.z) * v.w);
      continue;
    }
    if (p.epi == EPI_RESIDUAL) {
      const float4 q = *reinterpret_cast<const float4*>(p.R + (long long)m * p.ldr + n);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    if (n < p.n_split) {
      *reinterpret_cast<float4*>(p.C + (long long)m * p.ldc + n) = v;
    } else {
      const int pos = p.hm_pos ? *p.hm_pos : 0;
      *reinterpret_cast<float4*>(p.C1 + kv_dest(m, n - p.n_split, p.hm_rows_per_b, p.hm_cap, p.hm_H, pos)) = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Cluster variant (the default on sm_100a): the 8 CTAs of one 64 x 32 output tile form a thread-block
// cluster (1, 8, 1) and split K.  Each CTA computes the partial tile of its K chunk, then PUSHES the 8 rows
// owned by rank r into rank r's shared memory (st.shared::cluster); after ONE cluster barrier every rank sums
// the 8 partials of its own rows from LOCAL shared memory in rank order (bit-reproducible), applies the
// fused epilogue and stores.  No global scratch, no atomics, no second barrier (nobody reads remote memory
// after the barrier, so a CTA may exit as soon as it is done).
//
// MODE 0: exact fp32 FMA (parity anchor).  MODE 1: 3xTF32 on the tensor cores (mma.sync m16n8k8: x = hi + lo,
// hi*hi + hi*lo + lo*hi, fp32 accumulate: fp32-faithful, ~3e-6 relative) -- the measured FMA loop was
// 1.4-3.6 us of a 4.6-8.4 us node (scripts/trace_step.py), the MMA version is ~0.3 us.  MODE 2: 1xTF32.
// mma.sync and not tcgen05 on purpose: the tile is 64 x 32 x KC (KC <= 128) per CTA, 0.5 MFLOP; a TMEM
// allocation + descriptor set-up + commit/ld round trip costs more than the whole MMA loop.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t cluster_map(uint32_t smem_addr, unsigned rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f2(uint32_t addr, float x, float y) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cluster_f1(uint32_t addr, float x) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(x) : "memory");
}

constexpr int kDecLDB = kDecBN + 8;            // padded weight rows: conflict-free mma B-fragment reads
constexpr int kDecRedFloats = 8 * 8 * kDecBN;  // [src rank][8 local rows][32 cols] partials pushed to this CTA
template <int KC>
constexpr size_t dec_cluster_smem() {
  return (size_t)(kDecBM * (KC + 4) + KC * kDecLDB + kDecRedFloats + 64) * sizeof(float);
}

// Sum this rank's 64/S rows over the S source ranks (rank order, local shared memory) and apply the fused epilogue:
// RMSNorm row factor, gated-GELU / residual, plain store or head-major KV-cache append.  128 threads; S = 8: 8 rows x
// 16 column pairs, S = 16: 4 rows x 32 single columns (no gated-GELU, whose inputs are column pairs).
template <int S>
__device__ __forceinline__ void dec_reduce_epilogue(const DecGemmArgs& p, const float* Red, const float* Rss, unsigned rank, int n0) {
  constexpr int BN = kDecBN, R = kDecBM / S, CPT = (R * BN) / 128, LPR = BN / CPT;   // rows per rank, columns per thread, lanes per row
  static_assert(CPT == 1 || CPT == 2, "cluster size must be 8 or 16");
  const int tid = threadIdx.x;
  const int rl = tid / LPR, c0 = (tid % LPR) * CPT;
  const int m = (int)rank * R + rl, n = n0 + c0;
  float v[2] = {0.f, 0.f};
  float sst = 0.f;
#pragma unroll
  for (int s = 0; s < S; ++s) {
    if (CPT == 2) {
      const float2 q = *reinterpret_cast<const float2*>(&Red[(s * R + rl) * BN + c0]);
      v[0] += q.x; v[1] += q.y;
    } else {
      v[0] += Red[(s * R + rl) * BN + c0];
    }
    if (p.norm) sst += Rss[s * R + rl];
  }
  const bool valid = m < p.M && n < p.N;
  if (valid) {
    const float rs = p.norm ? 1.0f / sqrtf(sst / (float)p.K + p.eps) : 1.f;
    v[0] *= rs; v[1] *= rs;
    if (CPT == 2 && p.epi == EPI_GATED_GELU) {
      p.C[(long long)m * p.ldc + (n >> 1)] = gelu_tanh(v[0]) * v[1];
    } else {
      if (p.epi == EPI_RESIDUAL) {
        const float* r = p.R + (long long)m * p.ldr + n;
        v[0] += r[0];
        if (CPT == 2) v[1] += r[1];
      }
      float* dst = n < p.n_split ? p.C + (long long)m * p.ldc + n
                                 : p.C1 + kv_dest(m, n - p.n_split, p.hm_rows_per_b, p.hm_cap, p.hm_H, p.hm_pos ? *p.hm_pos : 0);
      if (CPT == 2) *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
      else dst[0] = v[0];
    }
  }
  if (p.ssq_out) {                      // LPR lanes share a row: fixed butterfly order
    float sq = valid ? fmaf(v[0], v[0], CPT == 2 ? v[1] * v[1] : 0.f) : 0.f;
#pragma unroll
    for (int o = LPR / 2; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((tid % LPR) == 0 && m < p.M) p.ssq_out[(long long)m * p.ssq_ld + n0 / BN] = sq;
  }
}

template <int KC, int MODE, bool TRACE, int S = 8>
__device__ __forceinline__ void dec_cluster_body(const DecGemmArgs& p, const int tile_x) {
  constexpr int BM = kDecBM, BN = kDecBN, NT = 128, LDA = KC + 4, LDB = kDecLDB;
  constexpr int R = BM / S;                   // rows of the tile each rank owns (8 for clusters of 8, 4 for clusters of 16)
  constexpr int WQ = KC * 8 / NT;            // weight 16-byte copies per thread
  constexpr int AQ = KC * 16 / NT;           // activation 16-byte copies per thread
  extern __shared__ __align__(16) float dsm[];
  float* As = dsm;                            // [BM][LDA] activations, row-major
  float* Bs = As + BM * LDA;                  // [KC][LDB] weights
  float* Red = Bs + KC * LDB;                 // [8 src][8 rows][BN] partial tiles of MY rows (written by all ranks)
  float* Rss = Red + kDecRedFloats;           // [8 src][8 rows] partial sums of squares of my rows

  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n0 = tile_x * BN;
  const int kbeg = blockIdx.y * KC;
  const bool tr = TRACE && p.trace != nullptr && tid == 0;
  const bool tr0 = tr && tile_x == 0 && blockIdx.y == 0;
  long long c0 = 0;
  if (tr) {
    atomicMin(p.trace, gtime_ns());
    c0 = clock64();
    if (tile_x == 0) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      if (blockIdx.y < 8) p.trace[8 + blockIdx.y] = smid;
    }
  }
  // "I am running": peers may write into my shared memory once every CTA of the cluster has arrived here
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

  // ---- load phase: cp.async (global -> shared, no register staging): all of the CTA's weights and activations
  // are in flight at once; one wait.  Weights are issued BEFORE the PDL wait (they do not depend on the
  // previous kernel).
#pragma unroll
  for (int i = 0; i < WQ; ++i) {
    const int idx = tid + i * NT;
    const int kr = idx >> 3, nq = idx & 7;
    const bool ok = n0 + nq * 4 < p.N;
    const float* src = p.W + (long long)(kbeg + kr) * p.ldw + (ok ? n0 + nq * 4 : 0);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(&Bs[kr * LDB + nq * 4])), "l"(src),
                 "r"(ok ? 16 : 0) : "memory");
  }
  pdl_wait();
  pdl_trigger();
#pragma unroll
  for (int i = 0; i < AQ; ++i) {
    const int idx = tid + i * NT;
    const int row = idx / (KC / 4), kq = idx % (KC / 4);
    const bool ok = row < p.M;
    const int col = kbeg + kq * 4;
    const float* src = (p.A2 != nullptr && col >= p.K0) ? p.A2 + (long long)(ok ? row : 0) * p.lda2 + (col - p.K0)
                                                         : p.A + (long long)(ok ? row : 0) * p.lda + col;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(&As[row * LDA + kq * 4])), "l"(src),
                 "r"(ok ? 16 : 0) : "memory");
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (tr0) p.trace[2] = (unsigned long long)(clock64() - c0);      // loads landed

  // sum of squares of this CTA's K chunk: thread -> (row tid/2, half of the chunk), fixed order
  float ss = 0.f;
  if (p.norm) {
    const float* ar = As + (tid >> 1) * LDA + (tid & 1) * (KC / 2);
#pragma unroll
    for (int k = 0; k < KC / 2; k += 4) {
      const float4 v = *reinterpret_cast<const float4*>(ar + k);
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  }

  const uint32_t red_base = tc::smem_u32(Red) + rank * (R * BN * 4);     // my slot [rank][..] in the owner's Red
  const uint32_t rss_base = tc::smem_u32(Rss) + rank * (R * 4);
  if (MODE == 0) {
    // ---- exact fp32: thread tile rows ty + 16 i (i < 4), columns tx*4 .. +3; k ascending ----
    const int tx = tid % 8, ty = tid / 8;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int k = 0; k < KC; k += 4) {
      float4 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(&As[(ty + 16 * i) * LDA + k]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) b[kk] = *reinterpret_cast<const float4*>(&Bs[(k + kk) * LDB + tx * 4]);
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
    if (tr0) p.trace[3] = (unsigned long long)(clock64() - c0);    // FMA loop done
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");    // every peer is running
#pragma unroll
    for (int i = 0; i < 4; ++i) {                                  // row ty + 16 i -> owner rank, local row
      const int row = ty + 16 * i;
      st_cluster_f4(cluster_map(red_base + (uint32_t)(((row % R) * BN + tx * 4) * 4), (unsigned)(row / R)),
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
  } else {
    // ---- tensor cores: a warp owns a 16-row group and nj of the four n8 column tiles.  With M <= 32 (<= 16) live
    // rows the four warps regroup as 2 row groups x 2 column halves (1 x 4), so the multiply time scales with M.
    const int g = lane >> 2, t = lane & 3;
    int rg, j0, nj;
    if (p.M > 32) { rg = warp; j0 = 0; nj = 4; }
    else if (p.M > 16) { rg = warp & 1; j0 = (warp >> 1) * 2; nj = 2; }
    else { rg = 0; j0 = warp; nj = 1; }
    float acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
    const float* a_lo_row = As + (rg * 16 + g) * LDA + t;
    const float* a_hi_row = a_lo_row + 8 * LDA;
#pragma unroll 2
    for (int k = 0; k < KC; k += 8) {
      const float af[4] = {a_lo_row[k], a_hi_row[k], a_lo_row[k + 4], a_hi_row[k + 4]};   // a0..a3
      uint32_t ah[4], al[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float hi, lo;
        split_tf32(af[e], hi, lo);
        ah[e] = __float_as_uint(hi);
        al[e] = __float_as_uint(lo);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j < nj) {
          const int col = (j0 + j) * 8 + g;
          const float bf[2] = {Bs[(k + t) * LDB + col], Bs[(k + t + 4) * LDB + col]};
          uint32_t bh[2], bl[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float hi, lo;
            split_tf32(bf[e], hi, lo);
            bh[e] = __float_as_uint(hi);
            bl[e] = __float_as_uint(lo);
          }
          if (MODE == 1) {                     // small terms first
            mma_tf32_16x8x8(acc[j], al, bh);
            mma_tf32_16x8x8(acc[j], ah, bl);
          }
          mma_tf32_16x8x8(acc[j], ah, bh);
        }
      }
    }
    if (tr0) p.trace[3] = (unsigned long long)(clock64() - c0);    // MMA loop done
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");    // every peer is running
    // c0,c1: row 16 rg + g -> rank 2 rg, local row g;  c2,c3: row 16 rg + g + 8 -> rank 2 rg + 1, local row g
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < nj) {
        const int ra = rg * 16 + g, rb = ra + 8;                   // rows of c0,c1 and of c2,c3
        const uint32_t col = (uint32_t)(((j0 + j) * 8 + 2 * t) * 4);
        st_cluster_f2(cluster_map(red_base + (uint32_t)((ra % R) * BN * 4) + col, (unsigned)(ra / R)), acc[j][0], acc[j][1]);
        st_cluster_f2(cluster_map(red_base + (uint32_t)((rb % R) * BN * 4) + col, (unsigned)(rb / R)), acc[j][2], acc[j][3]);
      }
    }
  }
  if (p.norm && (tid & 1) == 0) {
    const int row = tid >> 1;
    st_cluster_f1(cluster_map(rss_base + (uint32_t)((row % R) * 4), (unsigned)(row / R)), ss);
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (tr0) p.trace[4] = (unsigned long long)(clock64() - c0);      // partials exchanged

  dec_reduce_epilogue<S>(p, Red, Rss, rank, n0);
  if (tr) {
    if (tr0) p.trace[5] = (unsigned long long)(clock64() - c0);    // reduce + epilogue stores issued
    atomicMax(p.trace + 1, gtime_ns());
  }
}

template <int KC, int MODE, bool TRACE, int S = 8>
__global__ void __launch_bounds__(128)
sgemm_dec_cluster_kernel(const DecGemmArgs p) {
  dec_cluster_body<KC, MODE, TRACE, S>(p, (int)blockIdx.x);
}

// Two independent GEMMs that read the same inputs in ONE launch (column tiles [0, tiles0) belong to p0, the rest to
// p1): used for  y' = y + o.Wo  together with  q_raw = [o | y].[Wo.Wq ; Wq]  (the out-projection folded into the
// next projection with a precomposed weight block), which removes a kernel from the dependency chain.
template <int KC0, int KC1, int MODE, bool TRACE>
__global__ void __launch_bounds__(128)
sgemm_dec_cluster2_kernel(const DecGemmArgs p0, const DecGemmArgs p1, const int tiles0) {
  if ((int)blockIdx.x < tiles0) dec_cluster_body<KC0, MODE, TRACE>(p0, (int)blockIdx.x);
  else dec_cluster_body<KC1, MODE, TRACE>(p1, (int)blockIdx.x - tiles0);
}

template <int KC0, int KC1, int MODE>
inline int launch_dec_gemm_cluster2(const DecGemmArgs& a0, const DecGemmArgs& a1, cudaStream_t s, bool pdl) {
  constexpr size_t smem = dec_cluster_smem<(KC0 > KC1 ? KC0 : KC1)>();
  if (a0.M > kDecBM || a1.M > kDecBM || a0.K != 8 * KC0 || a1.K != 8 * KC1 || a0.N % kDecBN != 0 || a1.N % kDecBN != 0)
    return MT3_ERR_UNSUPPORTED;
  static bool attr_done = false;
  if (!attr_done) {
    MT3_CUDA_CHECK(cudaFuncSetAttribute(sgemm_dec_cluster2_kernel<KC0, KC1, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MT3_CUDA_CHECK(cudaFuncSetAttribute(sgemm_dec_cluster2_kernel<KC0, KC1, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int tiles0 = a0.N / kDecBN, tiles1 = a1.N / kDecBN;
  if (a0.trace)
    MT3_CUDA_CHECK(launch_kernel_cluster(sgemm_dec_cluster2_kernel<KC0, KC1, MODE, true>, dim3(tiles0 + tiles1, 8), dim3(128), smem, s,
                                         pdl, 8u, a0, a1, tiles0));
  else
    MT3_CUDA_CHECK(launch_kernel_cluster(sgemm_dec_cluster2_kernel<KC0, KC1, MODE, false>, dim3(tiles0 + tiles1, 8), dim3(128), smem, s,
                                         pdl, 8u, a0, a1, tiles0));
  MT3_LAUNCH_CHECK();
  return MT3_OK;
}

inline int launch_dec_gemm_out_q(const DecGemmArgs& a0, const DecGemmArgs& a1, int mode, cudaStream_t s, bool pdl) {
  if (mode == 1) return launch_dec_gemm_cluster2<48, 112, 1>(a0, a1, s, pdl);
  if (mode == 2) return launch_dec_gemm_cluster2<48, 112, 2>(a0, a1, s, pdl);
  return launch_dec_gemm_cluster2<48, 112, 0>(a0, a1, s, pdl);
}

template <int KC, int MODE, int S = 8>
inline int launch_dec_gemm_cluster_kc(const DecGemmArgs& a, cudaStream_t s, bool pdl) {
  constexpr size_t smem = dec_cluster_smem<KC>();
  static bool attr_done = false;
  if (!attr_done) {
    MT3_CUDA_CHECK(cudaFuncSetAttribute(sgemm_dec_cluster_kernel<KC, MODE, false, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MT3_CUDA_CHECK(cudaFuncSetAttribute(sgemm_dec_cluster_kernel<KC, MODE, true, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (S > 8) {
      MT3_CUDA_CHECK(cudaFuncSetAttribute(sgemm_dec_cluster_kernel<KC, MODE, false, S>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      MT3_CUDA_CHECK(cudaFuncSetAttribute(sgemm_dec_cluster_kernel<KC, MODE, true, S>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    }
    attr_done = true;
  }
  if (a.trace)
    MT3_CUDA_CHECK(launch_kernel_cluster(sgemm_dec_cluster_kernel<KC, MODE, true, S>, dim3(cdiv(a.N, kDecBN), S), dim3(128), smem, s, pdl, (unsigned)S, a));
  else
    MT3_CUDA_CHECK(launch_kernel_cluster(sgemm_dec_cluster_kernel<KC, MODE, false, S>, dim3(cdiv(a.N, kDecBN), S), dim3(128), smem, s, pdl, (unsigned)S, a));
  MT3_LAUNCH_CHECK();
  return MT3_OK;
}

template <int MODE>
inline int launch_dec_gemm_cluster_mode(const DecGemmArgs& a, cudaStream_t s, bool pdl) {
  // the long-K, narrow-N GEMM (MLP out: K = 1024, N = 512) has only N/32 x 8 = 128 CTAs of 4 warps at cluster size 8: a
  // cluster of 16 halves every CTA's K chunk and fills the machine (MT3_DEC_CLUSTER16=0 keeps clusters of 8)
  static const bool c16 = [] { const char* e = getenv("MT3_DEC_CLUSTER16"); return !(e && e[0] == '0'); }();
  if (c16 && a.K == 1024 && a.epi != EPI_GATED_GELU && a.N <= 512) return launch_dec_gemm_cluster_kc<64, MODE, 16>(a, s, pdl);
  switch (a.K / 8) {
    case 48: return launch_dec_gemm_cluster_kc<48, MODE>(a, s, pdl);
    case 64: return launch_dec_gemm_cluster_kc<64, MODE>(a, s, pdl);
    case 128: return launch_dec_gemm_cluster_kc<128, MODE>(a, s, pdl);
    default: return MT3_ERR_UNSUPPORTED;
  }
}

// mode: 0 exact fp32 FMA, 1 3xTF32 mma, 2 1xTF32 mma.  Returns MT3_ERR_UNSUPPORTED (without launching) when K
// does not split into 8 chunks of 48/64/128.
inline int launch_dec_gemm_cluster(const DecGemmArgs& a, int mode, cudaStream_t s, bool pdl) {
  if (a.M > kDecBM || a.N % 4 != 0 || a.lda % 4 != 0 || a.ldw % 4 != 0 || a.n_split % 4 != 0 || a.K % 8 != 0) return MT3_ERR_UNSUPPORTED;
  if (mode == 1) return launch_dec_gemm_cluster_mode<1>(a, s, pdl);
  if (mode == 2) return launch_dec_gemm_cluster_mode<2>(a, s, pdl);
  return launch_dec_gemm_cluster_mode<0>(a, s, pdl);
}

// One CTA per (32-column tile, 64-deep K chunk).
inline int dec_gemm_splits(int N, int K, int sm_count) {
  (void)N; (void)sm_count;
  return K / kDecKC;
}

inline int launch_dec_gemm(const DecGemmArgs& a, int splits, cudaStream_t s, bool pdl = false) {
  MT3_REQUIRE(a.M <= kDecBM, MT3_ERR_UNSUPPORTED, "decode gemm: M=%d > %d rows", a.M, kDecBM);
  MT3_REQUIRE(a.K == splits * kDecKC && a.N % 4 == 0 && a.lda % 4 == 0 && a.ldw % 4 == 0 && a.n_split % 4 == 0,
              MT3_ERR_UNSUPPORTED, "decode gemm: K=%d must be splits(%d) x %d; N=%d", a.K, splits, kDecKC, a.N);
  MT3_CUDA_CHECK(launch_kernel(sgemm_dec_kernel, dim3(cdiv(a.N, kDecBN), splits), dim3(128), 0, s, pdl, a));
  MT3_LAUNCH_CHECK();
  return MT3_OK;
}

// ---------------------------------------------------------------------------------------------
// Decode attention over head-major K/V, bulk-copy pipelined.
// ---------------------------------------------------------------------------------------------
constexpr int kAttKT = 32;                       // keys per tile (8 KB)
constexpr int kAttStages = 6;
constexpr int kAttTileFloats = kAttKT * 64;
constexpr int kAttThreads = 160;                 // 4 consumer warps + 1 producer warp

// K/V rows are read once per step and are 10x the size of L2: stream them with an evict-first policy so
// that the 104 MB of decoder weights (re-read every step) stay L2-resident.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               :
               : "r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(tc::smem_u32(bar)),
                 "l"(policy)
               : "memory");
}

inline size_t dec_attention_smem(int max_len) {
  return (size_t)(kAttStages * kAttTileFloats + 4 * ((max_len + 3) & ~3) + 8 * 64 + 64) * sizeof(float) + 2 * kAttStages * 8 + 64;
}

// q [B, ldq], head h at column q_off + h*64.  kv: head-major [b][2][H][cap][64].  out [B, ldo].
// len = (len_ptr ? *len_ptr : 0) + len_add.
template <bool TRACE>
__global__ void __launch_bounds__(kAttThreads)
dec_attention_bulk_kernel(const float* __restrict__ q, int ldq, int q_off, const float* __restrict__ kv, int H, int cap,
                          const int* __restrict__ len_ptr, int len_add, int max_len, float* __restrict__ out, int ldo,
                          const float* __restrict__ q_ssq, int q_ssq_n, int q_ssq_ld, float q_dim, float q_eps,
                          unsigned long long* trace) {
  extern __shared__ __align__(128) float sm[];
  const int ml4 = (max_len + 3) & ~3;
  float* ring = sm;                                              // [stages][32*64]
  float* sP = ring + kAttStages * kAttTileFloats;                // [4][ml4]: per-warp partial scores; row 0 becomes P
  float* sRed = sP + 4 * ml4;                                    // [8][64]
  float* sQ = sRed + 8 * 64;                                     // [64] the (scaled) query
  uint64_t* full = reinterpret_cast<uint64_t*>(sQ + 64);         // [stages]
  uint64_t* empty = full + kAttStages;
  __shared__ float s_stat[8];

  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int len = (len_ptr ? *len_ptr : 0) + len_add;
  const int nt = (len + kAttKT - 1) / kAttKT;
  const float* kbase = kv + (((long long)b * 2 + 0) * H + h) * (long long)cap * 64;
  const float* vbase = kv + (((long long)b * 2 + 1) * H + h) * (long long)cap * 64;
  const bool tr = TRACE && trace != nullptr && tid == 0;
  const bool tr0 = tr && blockIdx.x == 0 && blockIdx.y == 0;
  long long c0 = 0;
  if (tr) {
    atomicMin(trace, gtime_ns());
    c0 = clock64();
  }

  if (tid == 0) {
    for (int s = 0; s < kAttStages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 4);
    }
    tc::fence_barrier_init();
  }
  __syncthreads();

  if (warp == 4) {
    // ---- producer: K tiles 0..nt-1, then V tiles 0..nt-1 ----
    // self-attention (len_ptr set): the newest cache row comes from the preceding QKV GEMM -> wait;
    // the hoisted cross K/V is independent of every decode-step kernel -> stream it right away (PDL)
    // self-attention: every cache row but the newest (tile nt-1) was written by earlier steps, so those tiles are
    // streamed before the wait as well and only the last K tile is held back until the QKV GEMM has completed.
    pdl_trigger();
    if (lane == 0) {
      const uint64_t policy = l2_evict_first_policy();
      bool waited = len_ptr == nullptr;
      for (int j = 0; j < 2 * nt; ++j) {
        const int s = j % kAttStages;
        const uint32_t ph = (j / kAttStages) & 1;
        const int t = j < nt ? j : j - nt;
        if (!waited && t == nt - 1) {
          pdl_wait();
          waited = true;
        }
        tc::mbar_wait(&empty[s], ph ^ 1);
        const int keys = min(kAttKT, len - t * kAttKT);
        const uint32_t bytes = (uint32_t)keys * 64 * 4;
        const float* src = (j < nt ? kbase : vbase) + (long long)t * kAttTileFloats;
        tc::mbar_arrive_expect_tx(&full[s], bytes);
        bulk_g2s(ring + s * kAttTileFloats, src, bytes, &full[s], policy);
      }
    }
    return;
  }

  // ---- consumers (128 threads) ----
  pdl_wait();                                                    // q comes from the preceding GEMM
  pdl_trigger();
  if (tid < 16) {
    float4 q4 = *reinterpret_cast<const float4*>(q + (long long)b * ldq + q_off + h * 64 + tid * 4);
    if (q_ssq) {
      // q arrives un-normalised from a fused projection: scale by rsqrt(mean(y^2) + eps) of its input row, whose sum
      // of squares the producing GEMM left as q_ssq_n per-tile partials (summed in order)
      float ssq = 0.f;
      for (int i = 0; i < q_ssq_n; ++i) ssq += __ldg(q_ssq + (long long)b * q_ssq_ld + i);
      const float rs = 1.0f / sqrtf(ssq / q_dim + q_eps);
      q4.x *= rs; q4.y *= rs; q4.z *= rs; q4.w *= rs;
    }
    reinterpret_cast<float4*>(sQ)[tid] = q4;
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  // pass 1: scores, no cross-lane traffic.  Lane <-> key of the tile (a 256-byte row); warp w covers four of the
  // row's sixteen 16-byte chunks, rotated by the lane so that the eight lanes of a quarter-warp hit eight distinct
  // bank groups (rows are 64 words apart, i.e. bank-aligned); the four per-warp partial dot products of a key are
  // added in warp order by the softmax pass.
  {
    const float4* q4s = reinterpret_cast<const float4*>(sQ);
    float* part = sP + warp * ml4;
    for (int j = 0; j < nt; ++j) {
      const int s = j % kAttStages;
      tc::mbar_wait(&full[s], (j / kAttStages) & 1);
      if (tr0 && j == 0) trace[2] = (unsigned long long)(clock64() - c0);      // first K tile landed
      const float4* row = reinterpret_cast<const float4*>(ring + s * kAttTileFloats + lane * 64);
      const int k0 = j * kAttKT;
      float d = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int c = (4 * warp + jj + (lane & 7)) & 15;
        const float4 kx = row[c];
        const float4 qx = q4s[c];
        d = fmaf(qx.x, kx.x, d); d = fmaf(qx.y, kx.y, d); d = fmaf(qx.z, kx.z, d); d = fmaf(qx.w, kx.w, d);
      }
      if (k0 + lane < len) part[k0 + lane] = d;
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&empty[s]);
    }
  }
  if (tr0) trace[3] = (unsigned long long)(clock64() - c0);       // pass 1 (K tiles) consumed
  asm volatile("bar.sync 1, 128;" ::: "memory");                // all four partial rows are complete
  float lmax = -INFINITY;
  for (int k = tid; k < len; k += 128) {
    const float sc = ((sP[k] + sP[ml4 + k]) + sP[2 * ml4 + k]) + sP[3 * ml4 + k];
    sP[k] = sc;
    lmax = fmaxf(lmax, sc);
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
  if (tr0) trace[4] = (unsigned long long)(clock64() - c0);       // softmax done

  // pass 2: O = P V.  thread -> key group (tid / 16: 8 groups) x 4 dims (tid % 16)
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
  if (tr0) trace[5] = (unsigned long long)(clock64() - c0);       // pass 2 (V tiles) consumed
  *reinterpret_cast<float4*>(sRed + kg * 64 + d4 * 4) = acc;
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (tid < 64) {
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) s += sRed[g * 64 + tid];
    out[(long long)b * ldo + h * 64 + tid] = s * inv;
  }
  if (tr) {
    if (tr0) trace[6] = (unsigned long long)(clock64() - c0);
    atomicMax(trace + 1, gtime_ns());
  }
}

}  // namespace mt3
