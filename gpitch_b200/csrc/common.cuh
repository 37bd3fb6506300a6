// gpitch_b200 -- common device helpers (sm_100a).  fp64 throughout: gpitch's float_type is float64
// (reference: gpitch/matern12_spectral_mixture.py:8, gpitch/pdgp.py:172).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

#define GPX_OK 0
#define GPX_ERR_ARG (-1)
#define GPX_ERR_LAUNCH (-2)

namespace gpx { extern std::atomic<unsigned long long> g_launches; }   // kernels launched by this library (bench.py reports it)

#define GPX_CHECK_LAUNCH()                                  \
  do {                                                      \
    ++gpx::g_launches;                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return GPX_ERR_LAUNCH;          \
  } while (0)

namespace gpx {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// D(8x8) += A(8x4,row) * B(4x8,col), fp64 tensor pipe (SASS: DMMA.8x8x4).
// lane = 4*g + t:  a = A[g][t], b = B[t][g], c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0 (and broadcast to all if BCAST).  `red` = >= 32 doubles of smem.
template <bool BCAST>
__device__ __forceinline__ double block_sum(double v, double* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double s = (lane < nw) ? red[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) red[0] = s;
  }
  if (BCAST) {
    __syncthreads();
    v = red[0];
  } else {
    v = red[0];  // only thread 0's copy is guaranteed (after the w==0 branch)
  }
  return v;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// thread-block-cluster primitives (split-K GEMM, row-chunked lag binning): cluster barrier and loads from another CTA's
// shared memory (distributed shared memory)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ double ld_dsmem_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];\n" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

}  // namespace gpx
