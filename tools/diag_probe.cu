// Development probe: per-phase clock64() stamps of gpx::diag_block_kernel (one CTA, one 64 x 64 block).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/diag_probe tools/diag_probe.cu \
//        gpitch_b200/csrc/{api,composite,gemm,gemm_tma,ops,builder,grad_lag}.cu -lcuda && tools/diag_probe
// (it #includes chol.cu with GPX_DIAG_STAMP defined; -DGPX_RSQRT_NEWTON selects the two-step Newton reciprocal square root)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ long long g_stamp[32];
#define GPX_DIAG_STAMP(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_stamp[i] = clock64(); } while (0)
#include "../gpitch_b200/csrc/chol.cu"

__global__ void acc_kernel(double* maxerr) {
  // relative error of rsqrt_fast against correctly rounded 1/sqrt in units of 2^-53, over 2^22 arguments
  double worst = 0.0;
  for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < (1 << 22); i += blockDim.x * gridDim.x) {
    const double d = ldexp(1.0 + (double)i / (1 << 22) * 3.0, (i % 41) - 20);
    const double y = gpx::rsqrt_fast(d);
    const double ref = 1.0 / sqrt(d);
    const double e = fabs(y - ref) / ref * 9007199254740992.0;
    worst = fmax(worst, e);
  }
  for (int o = 16; o; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  if ((threadIdx.x & 31) == 0) atomicMax((unsigned long long*)maxerr, __double_as_longlong(worst));
}

__global__ void lat_kernel(double* out, long long* cyc, double seed) {
  __shared__ double sh[64];
  double a = seed, b = seed * 0.5;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 256; i++) a = fma(a, b, b);
  long long t1 = clock64();
  double r = a * 1e-300 + 2.0;
#pragma unroll
  for (int i = 0; i < 32; i++) r = gpx::rsqrt_fast(r) + 1.5;
  long long t2 = clock64();
  double q = r;
#pragma unroll
  for (int i = 0; i < 32; i++) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q)); q = y + 1.5; }
  long long t3 = clock64();
  double w = q;
#pragma unroll
  for (int i = 0; i < 32; i++) { sh[(threadIdx.x + i) & 63] = w; __syncwarp(); w = sh[(threadIdx.x + i + 1) & 63] + w; __syncwarp(); }
  long long t4 = clock64();
  double m = w;
#pragma unroll
  for (int i = 0; i < 64; i++) m = m * b;
  long long t5 = clock64();
  out[threadIdx.x] = a + r + q + w + m;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
}

int main() {
  {
    double* me; cudaMalloc(&me, 8); cudaMemset(me, 0, 8);
    acc_kernel<<<64, 256>>>(me);
    double h; cudaMemcpy(&h, me, 8, cudaMemcpyDeviceToHost);
    printf("rsqrt_fast max error vs 1/sqrt: %.3f x 2^-53\n", h);
  }
  {
    double* o; long long* c; cudaMalloc(&o, 8 * 64); cudaMalloc(&c, 8 * 8);
    lat_kernel<<<1, 32>>>(o, c, 0.999); lat_kernel<<<1, 32>>>(o, c, 0.999);
    long long h[8]; cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
    printf("dependent DFMA: %.1f cyc   rsqrt_fast + DADD: %.1f cyc   rsqrt.approx + DADD: %.1f cyc   STS/sync/LDS/DADD/sync: %.1f cyc  DMUL: %.1f\n",
           h[0] / 256.0, h[1] / 32.0, h[2] / 32.0, h[3] / 32.0, h[4] / 64.0);
  }
  const int M = 64;
  double *hA = (double*)malloc(sizeof(double) * M * M);
  srand(1);
  // SPD: B B^T + M I
  double* B = (double*)malloc(sizeof(double) * M * M);
  for (int i = 0; i < M * M; i++) B[i] = rand() / (double)RAND_MAX - 0.5;
  for (int i = 0; i < M; i++)
    for (int j = 0; j < M; j++) {
      double s = (i == j) ? 1.0 : 0.0;
      for (int k = 0; k < M; k++) s += B[i * M + k] * B[j * M + k];
      hA[i * M + j] = s;
    }
  double *A, *X; int* info;
  cudaMalloc(&A, sizeof(double) * M * M); cudaMalloc(&X, sizeof(double) * M * M); cudaMalloc(&info, 4);
  cudaMemset(info, 0, 4);
  cudaFuncSetAttribute(gpx::diag_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gpx::DIAG_SMEM);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int it = 0; it < 20; it++) {
    cudaMemcpy(A, hA, sizeof(double) * M * M, cudaMemcpyHostToDevice);
    cudaEventRecord(e0);
    gpx::diag_block_kernel<<<1, 256, gpx::DIAG_SMEM>>>(A, 0, M, X, 0, M, 0, M, info);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  printf("launch-to-end (events, best of 20): %.2f us  err=%s\n", best * 1e3, cudaGetErrorString(cudaGetLastError()));
  long long st[32];
  cudaMemcpyFromSymbol(st, g_stamp, sizeof(st));
  const char* names[25] = {"load", "sync", "p0 start", "p0 chol16", "p0 sync", "p0 panel", "p1 start(upd)", "p1 chol16",
                           "p1 sync", "p1 panel", "p2 start(upd)", "p2 chol16", "p2 sync", "p2 panel", "p3 start(upd)",
                           "p3 chol16", "p3 sync", "-", "factor done", "L written", "diag inverses", "inv i=1", "inv i=2", "inv i=3",
                           "X written"};
  for (int i = 1; i < 25; i++)
    if (st[i] && i != 17) {
      int j = i - 1; while (j > 0 && (st[j] == 0 || j == 17)) j--;
      printf("%2d %-16s +%6lld cycles  (t=%lld)\n", i, names[i], st[i] - st[j], st[i] - st[0]);
    }
  // check L L^T = A
  double* hL = (double*)malloc(sizeof(double) * M * M); double* hX = (double*)malloc(sizeof(double) * M * M);
  cudaMemcpy(hL, A, sizeof(double) * M * M, cudaMemcpyDeviceToHost); cudaMemcpy(hX, X, sizeof(double) * M * M, cudaMemcpyDeviceToHost);
  double e1m = 0, e2m = 0;
  for (int i = 0; i < M; i++)
    for (int j = 0; j <= i; j++) {
      double s = 0, t = 0;
      for (int k = 0; k < M; k++) { s += hL[i * M + k] * hL[j * M + k]; t += hX[i * M + k] * hL[k * M + j]; }
      e1m = fmax(e1m, fabs(s - hA[i * M + j])); e2m = fmax(e2m, fabs(t - (i == j)));
    }
  printf("max |L L^T - A| = %.2e   max |X L - I| = %.2e\n", e1m, e2m);
  return 0;
}
