// TEST INFRASTRUCTURE: a minimal CUDA-on-CPU emulator (fibers), so that the very same
// kernel sources under qsp_b200/csrc can be exercised by the CPU test-suite in a
// container without a GPU.  It is NOT a fallback: the product library (libqsp_b200.so) is
// built by nvcc only, the python package never loads the emulated build, and only
// tests/ link against it (tests/emu/build_emu.py -> tests/emu/_build/libqsp_emu.so).
//
// Model: every CUDA thread of a block is a ucontext fiber; blocks run one after another.
// __syncthreads() and the warp shuffles are "group barriers" on which a fiber yields
// until the whole group has arrived.  Execution is deterministic (fibers run in thread
// order), so missing barriers usually show up as wrong numbers rather than flakiness.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)

namespace emu {
struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  bool done = false;
  unsigned tid = 0;
};
struct Group {
  unsigned count = 0;
  unsigned gen = 0;
};
struct Block {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  std::vector<char> smem;
  Group bar;
  Group named[16];
  std::vector<Group> warp_bar;
  std::vector<double> xchg, xchg2;  // 32 doubles per warp each
  std::vector<uint32_t> tmem;       // tensor memory: 128 lanes x 512 columns (allocated on first use)
  std::function<void()> body;
  unsigned nthreads = 0;
  int cur = -1;
};
extern thread_local Block* g_blk;
extern thread_local dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;

inline void yield() { swapcontext(&g_blk->fibers[g_blk->cur].ctx, &g_blk->sched); }

inline void group_sync(Group& g, unsigned size) {
  unsigned my = g.gen;
  if (++g.count == size) {
    g.count = 0;
    g.gen++;
  } else {
    while (g.gen == my) yield();
  }
}

inline void trampoline() {
  Block* b = g_blk;
  b->body();
  b->fibers[b->cur].done = true;
  swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& kernel_call) {
  const unsigned nt = block.x;
  if (nt % 32 != 0) {
    fprintf(stderr, "emu: blockDim must be a multiple of 32\n");
    abort();
  }
  for (unsigned bx = 0; bx < grid.x; ++bx) {
    Block blk;
    blk.nthreads = nt;
    blk.smem.assign(smem_bytes + 64, 0);
    blk.warp_bar.resize(nt / 32);
    blk.xchg.assign(nt, 0.0);
    blk.xchg2.assign(nt, 0.0);
    blk.fibers.resize(nt);
    blk.body = [&]() { kernel_call(); };
    g_blk = &blk;
    g_blockIdx = dim3(bx);
    g_blockDim = block;
    g_gridDim = grid;
    for (unsigned t = 0; t < nt; ++t) {
      Fiber& f = blk.fibers[t];
      f.tid = t;
      f.stack.resize(256 * 1024);
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack.data();
      f.ctx.uc_stack.ss_size = f.stack.size();
      f.ctx.uc_link = &blk.sched;
      makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    unsigned alive = nt;
    while (alive) {
      for (unsigned t = 0; t < nt; ++t) {
        Fiber& f = blk.fibers[t];
        if (f.done) continue;
        blk.cur = (int)t;
        g_threadIdx = dim3(t);
        swapcontext(&blk.sched, &f.ctx);
        if (f.done) --alive;
      }
    }
    g_blk = nullptr;
  }
}
}  // namespace emu

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

inline void __syncthreads() { emu::group_sync(emu::g_blk->bar, emu::g_blk->nthreads); }
inline void __syncwarp(unsigned = 0xffffffffu) {
  emu::group_sync(emu::g_blk->warp_bar[emu::g_threadIdx.x / 32], 32);
}
inline double __shfl_sync(unsigned, double v, int src, int width = 32) {
  emu::Block* b = emu::g_blk;
  unsigned t = emu::g_threadIdx.x, w = t / 32, lane = t % 32;
  b->xchg[t] = v;
  emu::group_sync(b->warp_bar[w], 32);
  unsigned base = lane - (lane % width);
  double r = b->xchg[w * 32 + base + ((unsigned)src % width)];
  emu::group_sync(b->warp_bar[w], 32);
  return r;
}
inline double __shfl_xor_sync(unsigned m, double v, int mask, int width = 32) {
  unsigned lane = emu::g_threadIdx.x % 32;
  return __shfl_sync(m, v, (int)((lane ^ (unsigned)mask) % (unsigned)width), width);
}
inline int __shfl_sync(unsigned m, int v, int src, int width = 32) {
  return (int)__shfl_sync(m, (double)v, src, width);
}
inline int __any_sync(unsigned, int pred) {
  emu::Block* b = emu::g_blk;
  unsigned t = emu::g_threadIdx.x, w = t / 32;
  b->xchg[t] = pred ? 1.0 : 0.0;
  emu::group_sync(b->warp_bar[w], 32);
  int r = 0;
  for (unsigned l = 0; l < 32; ++l) r |= (b->xchg[w * 32 + l] != 0.0);
  emu::group_sync(b->warp_bar[w], 32);
  return r;
}
// tensor memory cell of the calling thread (lane 32 (warp % 4) + lane of the warp), column `col`
inline uint32_t* emu_tmem_cell(unsigned col) {
  emu::Block* b = emu::g_blk;
  if (b->tmem.empty()) b->tmem.assign(128 * 512, 0u);
  const unsigned t = emu::g_threadIdx.x, lane = 32 * ((t / 32) % 4) + t % 32;
  if (col >= 512) {
    fprintf(stderr, "emu: tensor-memory column %u out of range\n", col);
    abort();
  }
  return &b->tmem[lane * 512 + col];
}
inline int __shfl_xor_sync(unsigned m, int v, int mask, int width = 32) {
  return (int)__shfl_xor_sync(m, (double)v, mask, width);
}
// warp-level D(8x8) += A(8x4) B(4x8) with the fragment layout of mma.sync.m8n8k4.f64 (qsp_mma.cuh)
inline void emu_dmma(double& c0, double& c1, double a, double b) {
  emu::Block* blk = emu::g_blk;
  const unsigned t = emu::g_threadIdx.x, w = t / 32, lane = t % 32;
  blk->xchg[t] = a;
  blk->xchg2[t] = b;
  emu::group_sync(blk->warp_bar[w], 32);
  const unsigned row = lane >> 2, q = lane & 3;
  for (unsigned k = 0; k < 4; ++k) {
    const double av = blk->xchg[w * 32 + row * 4 + k];
    c0 = std::fma(av, blk->xchg2[w * 32 + (2 * q) * 4 + k], c0);
    c1 = std::fma(av, blk->xchg2[w * 32 + (2 * q + 1) * 4 + k], c1);
  }
  emu::group_sync(blk->warp_bar[w], 32);
}
inline void __threadfence_block() {}
inline int atomicAdd(int* p, int v) {
  int o = *p;
  *p = o + v;
  return o;
}
inline void emu_named_barrier(int id, int count) { emu::group_sync(emu::g_blk->named[id & 15], (unsigned)count); }
inline void emu_spin_yield() { emu::yield(); }
inline char* emu_dynamic_smem() {
  uintptr_t p = (uintptr_t)emu::g_blk->smem.data();
  return (char*)((p + 15) & ~(uintptr_t)15);
}
