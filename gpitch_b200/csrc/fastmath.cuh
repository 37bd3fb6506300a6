// FP64 exp(-r) and sqrt with fewer FP64-pipe slots than libm (the builder / gradient kernels are bound by the FP64
// pipe, which DMMA and DFMA share on B200: tools/fp64_peaks.cu).  Accuracy: <= 2 ulp (tests/test_gpu_primitives.py
// checks both against libm over the full working range).
#pragma once
#include "common.cuh"
#include <cmath>

namespace gpx {

// 2^(j/64), j = 0..63, filled once per device by init_fastmath() (computed on the host in long double).
// One copy per translation unit that includes this header (no relocatable device code in this build).
static __constant__ double c_exp2_64[64];

static int init_fastmath() {
  // __constant__ memory is per device (per context): one upload per device this translation unit is used on
  static bool done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return GPX_ERR_LAUNCH;
  if (done[dev]) return GPX_OK;
  double tab[64];
  for (int j = 0; j < 64; j++) tab[j] = (double)exp2l((long double)j / 64.0L);
  if (cudaMemcpyToSymbol(c_exp2_64, tab, sizeof(tab)) != cudaSuccess) return GPX_ERR_LAUNCH;
  done[dev] = true;
  return GPX_OK;
}

// Copy the table into shared memory (lane-divergent index -> shared, not constant, memory).  tab >= 64 doubles.
__device__ __forceinline__ void load_exp_table(double* tab) {
  for (int i = threadIdx.x; i < 64; i += blockDim.x) tab[i] = c_exp2_64[i];
}

// exp(-r) for r >= 0:  -r = (64 k + j) ln2/64 + y, |y| <= ln2/128;  exp(-r) = 2^k * 2^(j/64) * exp(y).
// 11 FP64-pipe instructions (libm exp: ~21).  Returns 0 for r > 700 (true value < 1e-304).
// CB: take the constants from the constant bank -- an FP64 instruction accepts a c[bank][offset] operand directly, whereas
// a 64-bit literal costs two UMOV issue slots whenever the compiler re-materialises it.  Measured: -11 % on the issue-bound
// single-component lag-histogram pass, but +8 % on its 4-component variant and on the builders, so it is opt-in.
static __constant__ double c_fm[8] = {
    92.332482616893656758, -1.0830424696223417e-02, -2.572804622327669e-14, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0,
    6755399441055744.0};

template <bool CB = false>
__device__ __forceinline__ double exp_neg(double r, const double* __restrict__ tab) {
  const double MAGIC = CB ? c_fm[7] : 6755399441055744.0;       // 1.5 * 2^52: low word of (x + MAGIC) = rint(x)
  const double INV = CB ? c_fm[0] : 92.332482616893656758;      // 64 / ln 2
  const double NL_HI = CB ? c_fm[1] : -1.0830424696223417e-02;  // -(ln2 / 64) with 17 trailing zero bits (0x3F862E42FEFA0000):
  const double NL_LO = CB ? c_fm[2] : -2.572804622327669e-14;   // n * L_HI is exact for |n| < 2^17;  L_LO = ln2/64 - L_HI
  double t = fma(-r, INV, MAGIC);
  const int n = __double2loint(t);
  t -= MAGIC;
  double y = fma(t, NL_HI, -r);
  y = fma(t, NL_LO, y);
  double p = fma(y, CB ? c_fm[3] : 1.0 / 720.0, CB ? c_fm[4] : 1.0 / 120.0);
  p = fma(p, y, CB ? c_fm[5] : 1.0 / 24.0);
  p = fma(p, y, CB ? c_fm[6] : 1.0 / 6.0);
  p = fma(p, y, 0.5);
  p = fma(p * y, y, y);                                 // exp(y) - 1
  const double T = tab[n & 63];
  double res = fma(T, p, T);
  const int k = n >> 6;                                 // floor division (n <= 0)
  res = __hiloint2double(__double2hiint(res) + (k << 20), __double2loint(res));
  return (r > 700.0) ? 0.0 : res;
}

// sqrt(s) for s > 0 (normal range): hardware reciprocal-sqrt seed (~2^-22) + coupled Goldschmidt/Newton steps.
// 7 FP64-pipe instructions + 1 MUFU (libm sqrt: ~14).
__device__ __forceinline__ double sqrt_pos(double s) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  double g = s * y, h = 0.5 * y;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-g, g, s);
  return fma(e, h, g);
}

// sqrt(s) and 1/sqrt(s) together (the gradient kernels need s / r): rinv is accurate to ~1e-14 relative.
__device__ __forceinline__ double sqrt_pos_rinv(double s, double& rinv) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  double g = s * y, h = 0.5 * y;
  double e = fma(-h, g, 0.5);
  g = fma(g, e, g);
  h = fma(h, e, h);
  e = fma(-g, g, s);
  rinv = h + h;
  return fma(e, h, g);
}

}  // namespace gpx
