// Tensor memory (TMEM, 256 KB per SM: 128 lanes x 512 columns x 32 bit) used as a second on-chip row store.
// tcgen05 has no FP64 MMA kind, but its load / store instructions move plain 32-bit words between the registers
// of a warp and the 32 TMEM lanes of the warp's quadrant (lanes 32 (warp % 4) .. +31; thread i of the warp owns
// lane base + i).  A complex128 element is four consecutive columns of the owning thread's lane.
// Warps w, w+4, w+8, ... of a CTA see the same quadrant: a row slice written by one of them can be read by
// another at the same lane index after a CTA barrier bracketed by tcgen05.fence.
#pragma once
#include "qsp_common.cuh"

#ifndef QSP_EMU
__device__ __forceinline__ uint32_t tmem_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one warp: allocate `ncols` columns (power of two >= 32) and publish the base address in *slot (shared memory)
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_smem_addr(slot)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// two complex128 (8 columns) per thread
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, cplx& a, cplx& b) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "r"(taddr)
               : "memory");
  a.x = __hiloint2double(r1, r0);
  a.y = __hiloint2double(r3, r2);
  b.x = __hiloint2double(r5, r4);
  b.y = __hiloint2double(r7, r6);
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, cplx a, cplx b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__double2loint(a.x)), "r"(__double2hiint(a.x)), "r"(__double2loint(a.y)), "r"(__double2hiint(a.y)),
               "r"(__double2loint(b.x)), "r"(__double2hiint(b.x)), "r"(__double2loint(b.y)), "r"(__double2hiint(b.y))
               : "memory");
}
// one complex128 (4 columns) per thread
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, cplx& a) {
  uint32_t r0, r1, r2, r3;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(taddr)
               : "memory");
  a.x = __hiloint2double(r1, r0);
  a.y = __hiloint2double(r3, r2);
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, cplx a) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__double2loint(a.x)),
               "r"(__double2hiint(a.x)), "r"(__double2loint(a.y)), "r"(__double2hiint(a.y))
               : "memory");
}
#else  // CPU emulation (tests/emu): a per-CTA array, the same addressing (the lane field is implied by the thread)
template <int NCOLS>
inline void tmem_alloc(uint32_t* slot) { *slot = 0; }
template <int NCOLS>
inline void tmem_dealloc(uint32_t) {}
inline void tmem_fence_before() {}
inline void tmem_fence_after() {}
inline void tmem_wait_ld() {}
inline void tmem_wait_st() {}
inline void tmem_ld1(uint32_t taddr, cplx& a) { memcpy(&a, emu_tmem_cell(taddr & 0xffffu), 16); }
inline void tmem_st1(uint32_t taddr, cplx a) { memcpy(emu_tmem_cell(taddr & 0xffffu), &a, 16); }
inline void tmem_ld2(uint32_t taddr, cplx& a, cplx& b) {
  tmem_ld1(taddr, a);
  tmem_ld1(taddr + 4, b);
}
inline void tmem_st2(uint32_t taddr, cplx a, cplx b) {
  tmem_st1(taddr, a);
  tmem_st1(taddr + 4, b);
}
#endif
