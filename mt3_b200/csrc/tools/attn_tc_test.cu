// Standalone bring-up test of the tcgen05 encoder-attention kernel (attention_tc.cuh): dumps the raw
// S = Q K^T accumulator from TMEM and compares S and O with a double-precision CPU reference on random
// and on structured inputs (uniform P, V = f(d), V = f(k)) so that an operand-layout bug is localised.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../attention_tc.cuh"

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

__global__ void split_k(const float* x, float* hi, float* lo, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float h, l; split_tf32(x[i], h, l); hi[i] = h; lo[i] = l; }
}

// mode: 0 random, 1 Q=0 & V random, 2 Q=0 & V[k][d]=d, 3 Q=0 & V[k][d]=(k==5), 4 Q=0 & V[k][d]=k
static int run(int B, int T, int H, int mode, bool split3, int variant) {
  const int Q = H * 64, ld = 3 * Q, M = B * T;
  std::vector<float> qkv((size_t)M * ld);
  for (int m = 0; m < M; ++m)
    for (int c = 0; c < ld; ++c) {
      const int sel = c / Q, d = c % 64, k = m % T;
      float v;
      if (sel == 0) v = mode == 0 ? frand() * 0.5f : 0.f;
      else if (sel == 1) v = frand() * 0.5f;
      else v = mode <= 1 ? frand() : mode == 2 ? (float)d : mode == 3 ? (k == 5 ? 1.f : 0.f) : (float)k;
      qkv[(size_t)m * ld + c] = v;
    }
  // per-head V^T [B*H*64, T]
  std::vector<float> vt((size_t)B * H * 64 * T);
  for (int b = 0; b < B; ++b) for (int h = 0; h < H; ++h) for (int d = 0; d < 64; ++d) for (int t = 0; t < T; ++t)
    vt[(((size_t)b * H + h) * 64 + d) * T + t] = qkv[(size_t)(b * T + t) * ld + 2 * Q + h * 64 + d];
  float *d_x, *d_hi, *d_lo, *d_o, *d_ol, *d_S, *d_vt, *d_vth, *d_vtl;
  CK(cudaMalloc(&d_vt, vt.size() * 4)); CK(cudaMalloc(&d_vth, vt.size() * 4)); CK(cudaMalloc(&d_vtl, vt.size() * 4));
  CK(cudaMemcpy(d_vt, vt.data(), vt.size() * 4, cudaMemcpyHostToDevice));
  split_k<<<(unsigned)((vt.size() + 255) / 256), 256>>>(d_vt, d_vth, d_vtl, (long long)vt.size());
  CK(cudaMalloc(&d_x, qkv.size() * 4)); CK(cudaMalloc(&d_hi, qkv.size() * 4)); CK(cudaMalloc(&d_lo, qkv.size() * 4));
  CK(cudaMalloc(&d_o, (size_t)M * Q * 4)); CK(cudaMalloc(&d_ol, (size_t)M * Q * 4));
  CK(cudaMalloc(&d_S, (size_t)B * H * T * T * 4));
  CK(cudaMemcpy(d_x, qkv.data(), qkv.size() * 4, cudaMemcpyHostToDevice));
  split_k<<<(unsigned)((qkv.size() + 255) / 256), 256>>>(d_x, d_hi, d_lo, (long long)qkv.size());
  CK(cudaMemset(d_o, 0xFF, (size_t)M * Q * 4)); CK(cudaMemset(d_ol, 0, (size_t)M * Q * 4));
  CK(cudaMemset(d_S, 0xFF, (size_t)B * H * T * T * 4));
  TcOperand op;
  if (make_operand(&op, split3 ? d_hi : d_x, split3 ? d_lo : nullptr, M, ld, ld) != MT3_OK) return 1;
  TcOperand opv;
  if (make_operand(&opv, split3 ? d_vth : d_vt, split3 ? d_vtl : nullptr, (uint64_t)B * H * 64, T, T, 64) != MT3_OK) return 1;
  if (launch_enc_attention_tc(op, opv, B, T, H, d_o, split3 ? d_ol : nullptr, split3, 0, d_S, variant) != MT3_OK) return 1;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  kernel error: %s\n", cudaGetErrorString(e)); return 2; }
  if (B == 64 && split3) {   // phase timeline of the production launch (no S dump): CTA (0,0,0) and a mid-grid CTA
    unsigned long long* d_t;
    CK(cudaMalloc(&d_t, 64 * 8));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaMemset(d_t, 0, 64 * 8));
      if (launch_enc_attention_tc(op, opv, B, T, H, d_o, d_ol, split3, 0, nullptr, 0, d_t) != MT3_OK) return 1;
      CK(cudaDeviceSynchronize());
    }
    unsigned long long t[64];
    CK(cudaMemcpy(t, d_t, sizeof(t), cudaMemcpyDeviceToHost));
    const char* mma_names[10] = {"Q landed", "K0 landed", "K1 landed", "S issued", "V0 landed", "P0 ready", "PV0 issued", "V1 landed", "P1 ready", "PV1 issued"};
    const char* sm_names[8] = {"S done", "row max done", "P0 start", "P0 written", "P1 start (PV0 done)", "P1 written", "PV1 done", "epilogue done"};
    for (int c = 0; c < 2; ++c) {
      printf("  timeline CTA 0 %s item (us since kernel start @1.965 GHz): MMA thread:", c == 0 ? "first" : "third");
      for (int i = 0; i < 10; ++i) printf(" %s %.2f |", mma_names[i], t[c * 32 + i] / 1965.0);
      printf("\n      softmax thread:");
      for (int i = 0; i < 8; ++i) printf(" %s %.2f |", sm_names[i], t[c * 32 + 16 + i] / 1965.0);
      printf("\n");
    }
    cudaFree(d_t);
  }
  std::vector<float> o((size_t)M * Q), ol((size_t)M * Q), S((size_t)B * H * T * T);
  CK(cudaMemcpy(o.data(), d_o, o.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ol.data(), d_ol, o.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(S.data(), d_S, S.size() * 4, cudaMemcpyDeviceToHost));
  double s_err = 0, s_ref = 0, o_err = 0, o_ref = 0;
  std::vector<double> srow(T);
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h)
      for (int q = 0; q < T; q += (T > 64 ? 7 : 1)) {
        double mx = -1e300;
        for (int k = 0; k < T; ++k) {
          double s = 0;
          for (int d = 0; d < 64; ++d)
            s += (double)qkv[(size_t)(b * T + q) * ld + h * 64 + d] * (double)qkv[(size_t)(b * T + k) * ld + Q + h * 64 + d];
          srow[k] = s; mx = fmax(mx, s);
          const double got = S[(((size_t)b * H + h) * T + q) * T + k];
          s_err = fmax(s_err, fabs(got - s)); s_ref = fmax(s_ref, fabs(s));
        }
        double sum = 0;
        for (int k = 0; k < T; ++k) { srow[k] = exp(srow[k] - mx); sum += srow[k]; }
        for (int d = 0; d < 64; ++d) {
          double acc = 0;
          for (int k = 0; k < T; ++k) acc += srow[k] * (double)qkv[(size_t)(b * T + k) * ld + 2 * Q + h * 64 + d];
          acc /= sum;
          const size_t idx = (size_t)(b * T + q) * Q + h * 64 + d;
          o_err = fmax(o_err, fabs((double)o[idx] + (double)ol[idx] - acc)); o_ref = fmax(o_ref, fabs(acc));
        }
      }
  printf("B=%d T=%3d H=%d mode=%d %s variant=%d  S rel_err=%.3e  O rel_err=%.3e", B, T, H, mode, split3 ? "tf32x3" : "tf32  ", variant,
         s_err / fmax(s_ref, 1e-30), o_err / fmax(o_ref, 1e-30));
  if (mode >= 2) {
    printf("  O[b0,h1,q37,d0..5]=");
    for (int d = 0; d < 6; ++d) printf("%.4f ", o[(size_t)37 * Q + 64 + d] + ol[(size_t)37 * Q + 64 + d]);
    printf(" d32..34=");
    for (int d = 32; d < 35; ++d) printf("%.4f ", o[(size_t)37 * Q + 64 + d] + ol[(size_t)37 * Q + 64 + d]);
  }
  printf("\n");
  cudaFree(d_vt); cudaFree(d_vth); cudaFree(d_vtl);
  cudaFree(d_x); cudaFree(d_hi); cudaFree(d_lo); cudaFree(d_o); cudaFree(d_ol); cudaFree(d_S);
  // P is rounded (to nearest) to tf32 and used as a single term: measured 2e-4 of max|O| on random inputs
  const double tol = split3 ? 5e-4 : 5e-3;
  return (o_err / fmax(o_ref, 1e-30) <= tol) ? 0 : 3;
}

int main() {
  srand(3);
  int bad = 0;
  for (int variant = 0; variant < 1; ++variant)
    for (int mode = 0; mode < 5; ++mode) {
      int r = run(2, 256, 6, mode, true, variant);
      if (r == 2) return 2;
      bad += (r != 0 && variant == 0);
    }
  for (int s3 = 0; s3 < 2; ++s3) {
    bad += run(1, 128, 2, 0, s3 != 0, 0) != 0;
    bad += run(3, 200, 6, 0, s3 != 0, 0) != 0;
    bad += run(64, 256, 6, 0, s3 != 0, 0) != 0;
  }
  printf("%s (%d failing cases with variant 0)\n", bad ? "ATTN_TC_TEST FAILED" : "ATTN_TC_TEST PASSED", bad);
  return bad ? 1 : 0;
}
