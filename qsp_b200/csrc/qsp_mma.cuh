// FP64 tensor-core and TMA primitives of the large-bond path.
//
//   dmma_8x8x4     warp-level DMMA  D(8x8) += A(8x4) * B(4x8)  (mma.sync.aligned.m8n8k4.row.col.f64: the only
//                  FP64 tensor instruction of sm_100a -- tcgen05 has no f64 kind).  Fragment layout (lane l):
//                  a = A[l / 4][l % 4],  b = B[l % 4][l / 4],  (c0, c1) = C[l / 4][2 * (l % 4) + {0, 1}].
//   tma_load_1d    bulk asynchronous copy global -> shared through the TMA unit (cp.async.bulk, SASS UBLKCP),
//                  completion counted in bytes on an mbarrier.
//
// The CPU test emulation (QSP_EMU) restates both with warp shuffles / memcpy so that the very same kernels run on
// the fiber emulator; that build is test infrastructure only.
#pragma once
#include "qsp_common.cuh"

#ifdef QSP_EMU
#include <string.h>
inline void dmma_8x8x4(double& c0, double& c1, double a, double b) { emu_dmma(c0, c1, a, b); }
struct qsp_mbar {
  long long tx;    // bytes still expected in this phase (may go negative before expect_tx is seen)
  int arrivals;    // arrivals still missing in this phase
  int count;       // arrivals per phase
  int phase;
  int pad;
};
inline void mbar_check(qsp_mbar* m) {
  if (m->arrivals == 0 && m->tx == 0) {
    m->phase ^= 1;
    m->arrivals = m->count;
  }
}
inline void mbar_init(qsp_mbar* m, int count = 1) {
  m->tx = 0;
  m->arrivals = m->count = count;
  m->phase = 0;
}
inline void mbar_arrive(qsp_mbar* m) {
  m->arrivals -= 1;
  mbar_check(m);
}
inline void mbar_expect_tx(qsp_mbar* m, unsigned bytes) {
  m->tx += bytes;
  mbar_arrive(m);
}
inline void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, qsp_mbar* m) {
  memcpy(dst_smem, src_gmem, bytes);
  m->tx -= bytes;
  mbar_check(m);
}
inline void mbar_wait(qsp_mbar* m, int parity) {
  while (m->phase == parity) emu_spin_yield();
}
inline void fence_mbar_init() {}
inline void fence_proxy_async() {}
#else
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
typedef unsigned long long qsp_mbar;  // 8-byte shared-memory mbarrier object
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(qsp_mbar* m, int count = 1) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(m)), "r"(count) : "memory");
}
// plain arrival (consumer side of a ring: "this warp has finished reading the stage")
__device__ __forceinline__ void mbar_arrive(qsp_mbar* m) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(m)) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this thread's generic-proxy shared-memory writes before later async-proxy (TMA) accesses to the same bytes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one arrival (the barrier was initialised with count 1) + the byte count this phase has to receive
__device__ __forceinline__ void mbar_expect_tx(qsp_mbar* m, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(m)), "r"(bytes) : "memory");
}
// dst / src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, qsp_mbar* m) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(m))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(qsp_mbar* m, int parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(m)),
      "r"(parity)
      : "memory");
}
#endif

// A complex 8x8 accumulator tile (real and imaginary parts as two DMMA C fragments).
struct ctile {
  double r0, r1, i0, i1;
};
__device__ __forceinline__ ctile ctile_zero() {
  ctile t;
  t.r0 = t.r1 = t.i0 = t.i1 = 0.0;
  return t;
}
// acc += A * B for complex 8x4 / 4x8 fragments a = (ar, ai), b = (br, bi): four real DMMAs
__device__ __forceinline__ void cdmma(ctile& acc, double ar, double ai, double br, double bi) {
  dmma_8x8x4(acc.r0, acc.r1, ar, br);
  dmma_8x8x4(acc.r0, acc.r1, -ai, bi);
  dmma_8x8x4(acc.i0, acc.i1, ar, bi);
  dmma_8x8x4(acc.i0, acc.i1, ai, br);
}
// acc += A * B^H with b the fragment of B's ROWS (b = B[n][k] loaded like an A fragment): (ar + i ai)(br - i bi)
__device__ __forceinline__ void cdmma_bh(ctile& acc, double ar, double ai, double br, double bi) {
  dmma_8x8x4(acc.r0, acc.r1, ar, br);
  dmma_8x8x4(acc.r0, acc.r1, ai, bi);
  dmma_8x8x4(acc.i0, acc.i1, ai, br);
  dmma_8x8x4(acc.i0, acc.i1, -ar, bi);
}
