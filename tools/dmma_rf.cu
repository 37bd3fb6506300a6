// Does DMMA throughput depend on operand register reuse / ordering?  (5 A-frags x 4 B-frags like the GEMM warp tile)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ORDER, int MT, int NT>
__global__ void __launch_bounds__(256) k(double* out, int iters, const double* in) {
  double a[MT], b[NT], c[MT][NT][2];
  for (int i = 0; i < MT; i++) a[i] = in[threadIdx.x + i];
  for (int j = 0; j < NT; j++) b[j] = in[threadIdx.x + 8 + j];
  for (int i = 0; i < MT; i++) for (int j = 0; j < NT; j++) { c[i][j][0] = i; c[i][j][1] = j; }
  for (int it = 0; it < iters; ++it) {
    if (ORDER == 0) {
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
    } else {
#pragma unroll
      for (int j = 0; j < NT; j++)
#pragma unroll
        for (int i = 0; i < MT; i++) dmma(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
    double t = a[0];
#pragma unroll
    for (int i = 0; i + 1 < MT; i++) a[i] = a[i + 1];
    a[MT - 1] = t;
  }
  double s = 0;
  for (int i = 0; i < MT; i++) for (int j = 0; j < NT; j++) s += c[i][j][0] + c[i][j][1];
  if (s == 123.456) out[0] = s;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); f(); cudaDeviceSynchronize(); float best = 1e30f;
  for (int r = 0; r < 3; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
  return best;
}
int main() {
  double *out, *in; cudaMalloc(&out, 1024); cudaMalloc(&in, 8192); cudaMemset(in, 0, 8192);
  int sms = 148, iters = 4000;
  for (int bps : {1, 2}) {
    int grid = sms * bps;
    float m0 = timeit([&] { k<0, 5, 4><<<grid, 256>>>(out, iters, in); });
    float m1 = timeit([&] { k<1, 5, 4><<<grid, 256>>>(out, iters, in); });
    float m2 = timeit([&] { k<0, 4, 4><<<grid, 256>>>(out, iters, in); });
    float m3 = timeit([&] { k<0, 8, 4><<<grid, 256>>>(out, iters, in); });
    auto tf = [&](float ms, int n) { return (double)grid * 8 * iters * n * 512.0 / ms * 1e-9; };
    printf("blocks/SM=%d: 5x4 A-outer %.2f TF, 5x4 B-outer %.2f TF, 4x4 %.2f TF, 8x4 %.2f TF\n", bps, tf(m0, 20), tf(m1, 20), tf(m2, 16), tf(m3, 32));
  }
  return 0;
}
