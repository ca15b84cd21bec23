// Standalone bring-up test of the tcgen05/TMA GEMM (gemm_tc.cuh) against a double-precision CPU
// reference and the exact-fp32 SIMT GEMM.  Run on the B200 box:  ./gemm_tc_test [quick]
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../gemm_tc.cuh"

using namespace mt3;

namespace mt3 {
std::string& last_error() { static thread_local std::string e; return e; }
int fail(int code, const char* fmt, ...) {
  char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
  last_error() = buf; fprintf(stderr, "FAIL(%d): %s\n", code, buf); return code;
}
std::atomic<uint64_t> g_launch_count{0};
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

__global__ void split_kernel(const float* x, float* hi, float* lo, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float h, l; split_tf32(x[i], h, l); hi[i] = h; lo[i] = l; }
}

struct Case { int M, N, K, epi; bool rs; };

static int run_case(const Case& c, bool split3, bool timing) {
  const int M = c.M, N = c.N, K = c.K;
  const int Nout = c.epi == EPI_GATED_GELU ? N / 2 : N;
  std::vector<float> A((size_t)M * K), W((size_t)K * N), Wt((size_t)N * K), R((size_t)M * N), rs(M), pe((size_t)64 * N);
  for (auto& v : A) v = frand();
  for (auto& v : W) v = frand() * 0.1f;
  for (auto& v : R) v = frand();
  for (auto& v : pe) v = frand();
  for (auto& v : rs) v = 0.5f + 0.5f * fabsf(frand());
  for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) Wt[(size_t)n * K + k] = W[(size_t)k * N + n];
  float *dA, *dAh, *dAl, *dWt, *dWh, *dWl, *dW, *dR, *dRh, *dRl, *drs, *dpe, *dC, *dCl, *dCs;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dAh, A.size() * 4)); CK(cudaMalloc(&dAl, A.size() * 4));
  CK(cudaMalloc(&dWt, Wt.size() * 4)); CK(cudaMalloc(&dWh, Wt.size() * 4)); CK(cudaMalloc(&dWl, Wt.size() * 4));
  CK(cudaMalloc(&dW, W.size() * 4));
  CK(cudaMalloc(&dR, R.size() * 4)); CK(cudaMalloc(&dRh, R.size() * 4)); CK(cudaMalloc(&dRl, R.size() * 4));
  CK(cudaMalloc(&drs, M * 4)); CK(cudaMalloc(&dpe, pe.size() * 4));
  CK(cudaMalloc(&dC, (size_t)M * Nout * 4)); CK(cudaMalloc(&dCl, (size_t)M * Nout * 4)); CK(cudaMalloc(&dCs, (size_t)M * Nout * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dWt, Wt.data(), Wt.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dR, R.data(), R.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(drs, rs.data(), M * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dpe, pe.data(), pe.size() * 4, cudaMemcpyHostToDevice));
  split_kernel<<<(unsigned)((A.size() + 255) / 256), 256>>>(dA, dAh, dAl, (long long)A.size());
  split_kernel<<<(unsigned)((Wt.size() + 255) / 256), 256>>>(dWt, dWh, dWl, (long long)Wt.size());
  split_kernel<<<(unsigned)((R.size() + 255) / 256), 256>>>(dR, dRh, dRl, (long long)R.size());
  CK(cudaMemset(dC, 0xFF, (size_t)M * Nout * 4));
  CK(cudaMemset(dCl, 0, (size_t)M * Nout * 4));
  CK(cudaDeviceSynchronize());

  TcOperand opA, opB;
  if (make_operand(&opA, split3 ? dAh : dA, split3 ? dAl : nullptr, M, K, K) != MT3_OK) return 1;
  if (make_operand(&opB, split3 ? dWh : dWt, split3 ? dWl : nullptr, N, K, K) != MT3_OK) return 1;
  TcGemmArgs a; memset(&a, 0, sizeof(a));
  a.M = M; a.N = N; a.K = K; a.row_scale = c.rs ? drs : nullptr; a.epi = c.epi;
  a.R_hi = split3 ? dRh : dR; a.R_lo = split3 ? dRl : nullptr; a.ldr = N;
  a.pe = dpe; a.pe_T = 64; a.pe_ld = N;
  a.C_hi = dC; a.C_lo = split3 ? dCl : nullptr; a.ldc = Nout; a.n_split = N;
  if (launch_tc_gemm(opA, opB, a, split3, 0) != MT3_OK) return 1;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); return 2; }

  // SIMT exact-fp32 reference on the device
  GemmArgs g; memset(&g, 0, sizeof(g));
  g.A = dA; g.lda = K; g.B = dW; g.ldb = N; g.M = M; g.N = N; g.K = K; g.row_scale = c.rs ? drs : nullptr; g.epi = c.epi;
  g.R = dR; g.ldr = N; g.pe = dpe; g.pe_T = 64; g.pe_ld = N; g.C = dCs; g.ldc = Nout; g.n_split = N;
  if (launch_sgemm(g, 0) != MT3_OK) return 1;
  CK(cudaDeviceSynchronize());

  std::vector<float> C((size_t)M * Nout), Cl((size_t)M * Nout), Cs((size_t)M * Nout);
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(Cl.data(), dCl, C.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(Cs.data(), dCs, C.size() * 4, cudaMemcpyDeviceToHost));
  // CPU double reference on a sample of rows
  double max_err = 0, max_err_simt = 0, max_ref = 0;
  const int row_step = M > 512 ? M / 97 : 1;
  for (int m = 0; m < M; m += row_step) {
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)A[(size_t)m * K + k] * (double)W[(size_t)k * N + n];
      if (c.rs) acc *= rs[m];
      if (c.epi == EPI_RESIDUAL) acc += R[(size_t)m * N + n];
      if (c.epi == EPI_ADD_PE) acc += pe[(size_t)(m % 64) * N + n];
      if (c.epi == EPI_GATED_GELU) {
        if (n & 1) continue;
        double acc1 = 0;
        for (int k = 0; k < K; ++k) acc1 += (double)A[(size_t)m * K + k] * (double)W[(size_t)k * N + n + 1];
        if (c.rs) acc1 *= rs[m];
        const double x = acc;
        const double ge = 0.5 * x * (1.0 + tanh(0.7978845608028654 * (x + 0.044715 * x * x * x)));
        acc = ge * acc1;
        const size_t idx = (size_t)m * Nout + n / 2;
        const double got = (double)C[idx] + (double)Cl[idx];
        max_err = fmax(max_err, fabs(got - acc)); max_err_simt = fmax(max_err_simt, fabs(Cs[idx] - acc)); max_ref = fmax(max_ref, fabs(acc));
        continue;
      }
      const size_t idx = (size_t)m * Nout + n;
      const double got = (double)C[idx] + (double)Cl[idx];
      max_err = fmax(max_err, fabs(got - acc)); max_err_simt = fmax(max_err_simt, fabs(Cs[idx] - acc)); max_ref = fmax(max_ref, fabs(acc));
    }
  }
  const double rel = max_err / max_ref, rel_simt = max_err_simt / max_ref;
  // tf32x3: the dropped lo.lo term is ~2^-22, the rest is the tensor core's truncating fp32 accumulation
  // over K terms (measured 3e-6 .. 9e-6 of max|C| for K = 384 .. 1024 vs 6e-7 .. 1.3e-6 for fp32 FMA)
  const double tol = split3 ? 2e-5 : 3e-3;
  double ms = 0;
  if (timing) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch_tc_gemm(opA, opB, a, split3, 0);
    cudaEventRecord(e0);
    const int it = 20;
    for (int i = 0; i < it; ++i) launch_tc_gemm(opA, opB, a, split3, 0);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float t; cudaEventElapsedTime(&t, e0, e1); ms = t / it;
  }
  const double tflops = ms > 0 ? 2.0 * M * N * K / (ms * 1e-3) / 1e12 : 0;
  printf("M=%5d N=%4d K=%4d epi=%d rs=%d %s  rel_err=%.3e (simt %.3e) tol=%.0e %s", M, N, K, c.epi, (int)c.rs,
         split3 ? "tf32x3" : "tf32  ", rel, rel_simt, tol, rel <= tol ? "OK  " : "FAIL");
  if (timing) printf("  %.3f ms  %.1f TFLOP/s (algorithmic)", ms, tflops);
  printf("\n");
  cudaFree(dA); cudaFree(dAh); cudaFree(dAl); cudaFree(dWt); cudaFree(dWh); cudaFree(dWl); cudaFree(dW); cudaFree(dR);
  cudaFree(dRh); cudaFree(dRl); cudaFree(drs); cudaFree(dpe); cudaFree(dC); cudaFree(dCl); cudaFree(dCs);
  return rel <= tol ? 0 : 3;
}

int main(int argc, char** argv) {
  srand(1);
  int bad = 0;
  const Case small[] = {
      {128, 128, 32, EPI_STORE, false},   {128, 128, 64, EPI_STORE, false},  {128, 128, 512, EPI_STORE, false},
      {256, 256, 384, EPI_STORE, true},   {200, 384, 512, EPI_RESIDUAL, true}, {512, 512, 1024, EPI_ADD_PE, false},
      {512, 2048, 512, EPI_GATED_GELU, true}, {384, 1152, 512, EPI_STORE, true},
  };
  for (const auto& c : small)
    for (int s3 = 0; s3 < 2; ++s3) {
      int r = run_case(c, s3 != 0, false);
      if (r == 2) { printf("aborting after kernel error\n"); return 2; }
      bad += r != 0;
    }
  if (argc < 2) {
    const Case big[] = {{16384, 1152, 512, EPI_STORE, true}, {16384, 512, 384, EPI_RESIDUAL, false},
                        {16384, 2048, 512, EPI_GATED_GELU, true}, {16384, 512, 1024, EPI_RESIDUAL, false},
                        {16384, 768, 512, EPI_STORE, false}};
    for (const auto& c : big)
      for (int s3 = 0; s3 < 2; ++s3) bad += run_case(c, s3 != 0, true) != 0;
  }
  printf("%s (%d failing cases)\n", bad ? "GEMM_TC_TEST FAILED" : "GEMM_TC_TEST PASSED", bad);
  return bad ? 1 : 0;
}
