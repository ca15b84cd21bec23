// Compile-only probe for DESIGN.md section 6, item 1 (not part of libmt3b200.so, never run on a GPU so far):
// which SASS forms ptxas (12.9, sm_100a) gives a GEMM inner loop whose A operand is warp-uniform (lane = output column,
// 16 rows x 4 columns of accumulators per lane).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -c uniform_fma_probe.cu && cuobjdump -sass uniform_fma_probe.o
// Result (profiles/sass_uniform_fma_probe.txt): with A in constant memory the loop is LDCU (uniform-datapath load) +
// `FFMA R, R.reuse, UR, R` -- the uniform-register operand form exists and needs two vector-register reads per FMA;
// with A in shared memory at a thread-independent address ptxas keeps LDS + the three-register FFMA (no automatic R2UR).
#include <cstdint>
__constant__ float cA[4096];
// lane = column; 16 rows x 4 cols per lane; A from constant memory (warp-uniform operand)
__global__ void k_const(const float4* __restrict__ W, float* __restrict__ C, int K) {
  float acc[16][4];
#pragma unroll
  for (int m = 0; m < 16; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float4 w = W[k * 32 + threadIdx.x];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const float a = cA[m * 256 + k];
      acc[m][0] = fmaf(a, w.x, acc[m][0]); acc[m][1] = fmaf(a, w.y, acc[m][1]);
      acc[m][2] = fmaf(a, w.z, acc[m][2]); acc[m][3] = fmaf(a, w.w, acc[m][3]);
    }
  }
#pragma unroll
  for (int m = 0; m < 16; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) C[(m * 4 + j) * 32 + threadIdx.x] = acc[m][j];
}
// same with A in shared memory at a thread-independent address
__global__ void k_smem(const float4* __restrict__ W, const float* __restrict__ A, float* __restrict__ C, int K) {
  __shared__ float sA[16 * 64];
  for (int i = threadIdx.x; i < 16 * 64; i += 32) sA[i] = A[i];
  __syncwarp();
  float acc[16][4];
#pragma unroll
  for (int m = 0; m < 16; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[m][j] = 0.f;
  for (int k = 0; k < 64; ++k) {
    const float4 w = W[k * 32 + threadIdx.x];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const float a = sA[m * 64 + k];
      acc[m][0] = fmaf(a, w.x, acc[m][0]); acc[m][1] = fmaf(a, w.y, acc[m][1]);
      acc[m][2] = fmaf(a, w.z, acc[m][2]); acc[m][3] = fmaf(a, w.w, acc[m][3]);
    }
  }
#pragma unroll
  for (int m = 0; m < 16; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j) C[(m * 4 + j) * 32 + threadIdx.x] = acc[m][j];
}
