// Common device-side definitions for the qsp_b200 kernels (complex128 arithmetic,
// block reductions, launch abstraction).  Compiled by nvcc for sm_100a; the CPU
// test-suite compiles the same sources against tests/emu/cuda_emu.h (-DQSP_EMU) to check
// kernel logic without a GPU -- that build is test infrastructure, never a fallback.
#pragma once

#ifdef QSP_EMU
#include "cuda_emu.h"
#define QSP_SMEM_BASE() emu_dynamic_smem()
typedef void* qsp_stream_t;
#define QSP_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
extern __shared__ __align__(16) char qsp_dyn_smem[];
#define QSP_SMEM_BASE() (qsp_dyn_smem)
typedef cudaStream_t qsp_stream_t;
#define QSP_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<dim3(grid), dim3(block), (size_t)(smem), (stream)>>>(__VA_ARGS__)
#endif

#include <stdint.h>

// A pointer handed to an out-of-line device function is generic to the compiler even when every
// caller passes shared memory.  QSP_ASSUME_SHARED(flag, p) re-derives p from the dynamic shared
// window (base + 32-bit offset) when the compile-time flag is set, so its accesses become LDS/STS.
#ifdef QSP_EMU
#define QSP_ASSUME_SHARED(flag, p) ((void)0)
#else
#define QSP_ASSUME_SHARED(flag, p)                                                                        \
  do {                                                                                                    \
    if (flag)                                                                                             \
      p = (decltype(p))(qsp_dyn_smem +                                                                    \
                        ((unsigned)__cvta_generic_to_shared(p) - (unsigned)__cvta_generic_to_shared(qsp_dyn_smem))); \
  } while (0)
#endif

#ifdef QSP_EMU
inline long long qsp_clock() { return 0; }
inline double qsp_rsqrt(double x) { return 1.0 / sqrt(x); }
inline void qsp_spin_hint() { emu_spin_yield(); }
inline void qsp_named_barrier(int id, int count) { emu_named_barrier(id, count); }
#else
__device__ __forceinline__ long long qsp_clock() { return clock64(); }
__device__ __forceinline__ double qsp_rsqrt(double x) { return rsqrt(x); }
__device__ __forceinline__ void qsp_spin_hint() {}
// bar.sync on one of the 15 named hardware barriers (id 1..15), `count` threads (multiple of 32)
__device__ __forceinline__ void qsp_named_barrier(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
#endif

#define QSP_FULL_MASK 0xffffffffu

struct __attribute__((aligned(16))) cplx {
  double x, y;
};

__device__ __forceinline__ cplx cmake(double x, double y) {
  cplx r;
  r.x = x;
  r.y = y;
  return r;
}
__device__ __forceinline__ cplx cconj(cplx a) { return cmake(a.x, -a.y); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cmake(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cmake(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return cmake(a.x * s, a.y * s); }
__device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
// acc += a * b
__device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}
// acc += conj(a) * b
__device__ __forceinline__ void cfma_conja(cplx& acc, cplx a, cplx b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ cplx cdiv(cplx a, cplx b) {
  double d = 1.0 / cabs2(b);
  cplx n = cmul(a, cconj(b));
  return cmake(n.x * d, n.y * d);
}

__host__ __device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(QSP_FULL_MASK, v, m); }

// Sum over the whole block; every thread gets the result.  red: >= 33 doubles of smem.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
  __syncthreads();  // protect red[] from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nw; ++w) t += red[w];
  return t;
}
__device__ __forceinline__ double block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, shfl_xor_d(v, o));
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = red[0];
  for (int w = 1; w < nw; ++w) t = fmax(t, red[w]);
  return t;
}

// status bits reported per instance
#define QSP_ST_JACOBI_NOCONV 1   // Jacobi SVD hit the sweep limit
#define QSP_ST_BOND_OVERFLOW 2   // a bond outgrew the padded buffer
#define QSP_ST_NONUNITARY 4      // a generated gate failed the unitarity check
#define QSP_ST_DEGENERATE 8      // D=2 Schmidt pair degenerate: completion gauge ill-defined
#define QSP_ST_BADNORM 16        // zero / non-finite norm
// relative gap (s_a - s_b) / s_0 below which two Schmidt values at the D=2 cut count as degenerate
#define QSP_DEGENERATE_GAP 1e-6
