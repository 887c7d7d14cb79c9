// libqsp_b200_dev.so (include/qsp_b200_dev.h): the product ABI compiled with its development hooks
// (qsp_debug_flags) plus measurement / development entry points that do not belong in the product
// library: FP64 pipe probes (bench.py takes its roofline denominators from them), the Jacobi
// micro-benchmarks, the shared-memory probe.  The test emulation (tests/emu) is built from this file too.
#define QSP_DEV_BUILD 1
#include "qsp_capi.cu"
#include <vector>
#include "../../include/qsp_b200_dev.h"

// FP64 throughput probe (bench.py measures the DFMA roofline with it; MEASURED_PEAKS.json
// carries no FP64 figure).  16 independent FMA chains per thread.
__global__ void k_dfma_probe(double* out, int iters) {
  double a[16];
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-9 * blockIdx.x;
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = k * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = fma(a[k], x, y);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  if (s == 12345.678) out[0] = s;  // never true; keeps the loop alive
}
extern "C" int qsp_dfma_probe(double* out, int blocks, int threads, int iters, void* stream) {
  if (!out || blocks <= 0 || threads <= 0 || threads % 32) return QSP_ERR_ARG;
  QSP_LAUNCH(k_dfma_probe, blocks, threads, 0, (qsp_stream_t)stream, out, iters);
  ++g_launches;
  return check_launch();
}

// FP64 tensor-pipe probe: warp-level DMMA (mma.sync m8n8k4, the only FP64 tensor instruction on
// sm_100a -- tcgen05 has no f64 kind).  8 independent accumulator pairs per thread, 512 flops per
// warp instruction.  bench.py reports it next to the DFMA figure ("why no tensor cores").
__global__ void k_dmma_probe(double* out, int iters) {
#ifndef QSP_EMU
  double c[8][2];
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (blockIdx.x + 1);
#pragma unroll
  for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = k * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[k][0]), "+d"(c[k][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
  if (s == 12345.678) out[0] = s;
#else
  (void)out;
  (void)iters;
#endif
}
extern "C" int qsp_dmma_probe(double* out, int blocks, int threads, int iters, void* stream) {
  if (!out || blocks <= 0 || threads <= 0 || threads % 32) return QSP_ERR_ARG;
#ifdef QSP_EMU
  return QSP_ERR_UNSUPPORTED;
#else
  QSP_LAUNCH(k_dmma_probe, blocks, threads, 0, (qsp_stream_t)stream, out, iters);
  ++g_launches;
  return check_launch();
#endif
}

// Jacobi micro-benchmark (development aid): every CTA orthogonalises its own copy of an
// nrows x ldw matrix (W[B][nrows][ldw]) in place; sweeps[B] receives the sweep counts.
template <int NTMAX>
__global__ void __launch_bounds__(NTMAX) k_jacobi_bench(cplx* W, int nrows, int ndotp, int ldw, int B, int cap_elems,
                                                        int* sweeps, int* ok, int use_tmem, int presort) {
  char* sm = QSP_SMEM_BASE();
  cplx* Wsm = (cplx*)sm;
  double* red = (double*)(sm + (size_t)cap_elems * sizeof(cplx));
  uint32_t* tslot = (uint32_t*)(red + 58);
  use_tmem = use_tmem && blockDim.x == 512;
  if (use_tmem && (threadIdx.x >> 5) == 0) tmem_alloc<256>(tslot);
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  JacobiIO io;
  io.src = nullptr;
  io.src_ld = io.src_cols = 0;
  io.norms_off = (int)((char*)(red + 64) - sm);  // double[nrows] behind red: scratch of the norm ordering
  io.norms_done = false;
  io.tmem = use_tmem ? *tslot : 0u;
  io.tmem_ok = use_tmem != 0;
  io.presort = presort != 0;
  for (int inst = blockIdx.x; inst < B; inst += gridDim.x) {
    int ns = 0;
    const double tol = 4e-16 * sqrt((double)ndotp);
    bool c = jacobi_svd_rows(W + (long)inst * nrows * ldw, nrows, ndotp, ldw, Wsm, cap_elems, 8, tol * tol, 60, red, &ns, &io);
    if (threadIdx.x == 0) {
      sweeps[inst] = ns;
      ok[inst] = c ? 1 : 0;
    }
    __syncthreads();
  }
  if (use_tmem) {
    tmem_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc<256>(io.tmem);
  }
}
// Experiment: "fat thread" geometry -- 256 threads, teams of 4 lanes, 32 columns per lane (row half
// in 128 registers), so that latency is hidden by instruction-level parallelism instead of warps.
__global__ void __launch_bounds__(256) k_jacobi_bench_fat(cplx* W, int nrows, int ldw, int B, int* sweeps, int* ok) {
  char* sm = QSP_SMEM_BASE();
  const int ne_ = (nrows + 1) / 2;
  const int cap_elems = ne_ * ldw + (ne_ + 1) / 2 + 1;
  double* red = (double*)(sm + (size_t)cap_elems * sizeof(cplx));
  const int tpp = 4;
  for (int inst = blockIdx.x; inst < B; inst += gridDim.x) {
    int ns = 0;
    const double tol = 4e-16 * sqrt((double)ldw);
    bool c = jacobi_rows_oe<32, 1, false, 4, true, 1>(W + (long)inst * nrows * ldw, nrows, ldw / tpp, ldw / tpp, 0, tpp, tol * tol, 60,
                                          (int)((char*)red - sm), &ns);
    if (threadIdx.x == 0) {
      sweeps[inst] = ns;
      ok[inst] = c ? 1 : 0;
    }
    __syncthreads();
  }
}
extern "C" int qsp_jacobi_bench_fat(qsp_c128* W, int nrows, int ldw, int B, int* sweeps, int* ok, void* stream) {
  DevInfo d = dev_info();
  if (nrows / 2 > 64 || ldw != 128) return QSP_ERR_ARG;
  size_t smem = ((size_t)((nrows + 1) / 2) * ldw + (nrows + 3) / 4 + 1) * sizeof(cplx) + 64 * sizeof(double) + 64;
  int grid = B < d.sms ? B : d.sms;
  QSP_PREP_KERNEL(k_jacobi_bench_fat, smem);
  QSP_LAUNCH(k_jacobi_bench_fat, grid, 256, smem, (qsp_stream_t)stream, (cplx*)W, nrows, ldw, B, sweeps, ok);
  ++g_launches;
  return check_launch();
}
// Shared-memory bandwidth probe with the Jacobi access pattern (teams of tpp lanes, 16-byte
// elements, 16 per lane, rows of 2 KB): mode 0 = load only, 1 = store only, 2 = load + store.
__global__ void __launch_bounds__(512) k_smem_probe(double* out, int iters, int mode) {
  cplx* S = (cplx*)QSP_SMEM_BASE();
  const int tid = threadIdx.x, tpp = 8, team = tid / tpp, tl = tid % tpp;
  for (int i = tid; i < 64 * 128; i += blockDim.x) S[i] = cmake(i, -i);
  __syncthreads();
  cplx R[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) R[j] = cmake(j, tid);
  cplx acc = cmake(0, 0);
  for (int it = 0; it < iters; ++it) {
    cplx* xs = S + ((team + (it & 1)) & 63) * 128 + tl;
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) acc = cadd(acc, xs[j * tpp]);
    } else if (mode == 1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) xs[j * tpp] = R[j];
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        cplx x = xs[j * tpp];
        xs[j * tpp] = R[j];
        R[j] = x;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) acc = cadd(acc, R[j]);
  if (acc.x == 1.2345e300) out[0] = acc.y;
}
extern "C" int qsp_smem_probe(double* out, int blocks, int iters, int mode, void* stream) {
  size_t smem = 64 * 128 * sizeof(cplx);
  QSP_PREP_KERNEL(k_smem_probe, smem);
  QSP_LAUNCH(k_smem_probe, blocks, 512, smem, (qsp_stream_t)stream, out, iters, mode);
  ++g_launches;
  return check_launch();
}
// Tensor-memory bandwidth probe with the Jacobi access pattern: every thread moves 16 complex128 (64 columns of its
// TMEM lane) per iteration, two per instruction.  mode 0 = load only (wait per chunk), 1 = store only,
// 2 = load + store, 3 = load only with one wait per 16 elements.  out[0] = cycles per iteration (block 0).
#include "qsp_tmem.cuh"
__global__ void __launch_bounds__(512) k_tmem_probe(double* out, int iters, int mode) {
#ifndef QSP_EMU
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<256>(&tslot);
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t base = tslot;
  const uint32_t mine = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  cplx R[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) R[j] = cmake(j, tid);
#pragma unroll
  for (int j = 0; j < 16; j += 2) tmem_st2(mine + 4 * j, R[j], R[j + 1]);
  tmem_wait_st();
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  cplx acc = cmake(0, 0);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        cplx a, b;
        tmem_ld2(mine + 4 * j, a, b);
        tmem_wait_ld();
        acc = cadd(acc, cadd(a, b));
      }
    } else if (mode == 1) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) tmem_st2(mine + 4 * j, R[j], R[j + 1]);
      tmem_wait_st();
    } else if (mode == 2) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        cplx a, b;
        tmem_ld2(mine + 4 * j, a, b);
        tmem_wait_ld();
        tmem_st2(mine + 4 * j, R[j], R[j + 1]);
        R[j] = a;
        R[j + 1] = b;
      }
      tmem_wait_st();
    } else {
      cplx a[16];
#pragma unroll
      for (int j = 0; j < 16; j += 2) tmem_ld2(mine + 4 * j, a[j], a[j + 1]);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc = cadd(acc, a[j]);
    }
  }
  const long long t1 = clock64();
#pragma unroll
  for (int j = 0; j < 16; ++j) acc = cadd(acc, R[j]);
  if (acc.x == 1.2345e300) out[1] = acc.y;
  if (blockIdx.x == 0 && tid == 0) out[0] = (double)(t1 - t0) / iters;
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(base);
#endif
}
extern "C" int qsp_tmem_probe(double* out, int blocks, int iters, int mode, void* stream) {
  QSP_LAUNCH(k_tmem_probe, blocks, 512, 0, (qsp_stream_t)stream, out, iters, mode);
  ++g_launches;
  return check_launch();
}
extern "C" void qsp_jacobi_variant(int v) {
#ifdef QSP_JACOBI_ABLATION
  cudaMemcpyToSymbol(g_jvar, &v, sizeof(int));
#else
  (void)v;
#endif
}
extern "C" int qsp_jacobi_bench(qsp_c128* W, int nrows, int ndotp, int ldw, int B, int nt, int* sweeps, int* ok,
                                void* stream) {
  DevInfo d = dev_info();
  long cap = ((long)d.max_smem - 4096 - (long)nrows * 8) / (long)sizeof(cplx);
  long need = (long)nrows * ldw;
  if (need < cap) cap = need;
  size_t smem = (size_t)cap * sizeof(cplx) + (64 + (size_t)nrows + 2) * sizeof(double) + 64;
  int grid = B < d.sms ? B : d.sms;
  if (nt > 512) {
    QSP_PREP_KERNEL(k_jacobi_bench<1024>, smem);
    QSP_LAUNCH(k_jacobi_bench<1024>, grid, nt, smem, (qsp_stream_t)stream, (cplx*)W, nrows, ndotp, ldw, B, (int)cap, sweeps,
               ok, 0, g_nosort ? 0 : 1);
  } else {
    QSP_PREP_KERNEL(k_jacobi_bench<512>, smem);
    QSP_LAUNCH(k_jacobi_bench<512>, grid, nt, smem, (qsp_stream_t)stream, (cplx*)W, nrows, ndotp, ldw, B, (int)cap, sweeps,
               ok, g_tmem, g_nosort ? 0 : 1);
  }
  ++g_launches;
  return check_launch();
}


// Block-Jacobi SVD of the large-bond path on its own (tests, micro-benchmarks, ncu): the rows of the B matrices
// W[B][nrows][ldw] are orthogonalised in place (ldw, ndotp multiples of 32, zero padded); sweeps[B] receives the sweep
// counts, ok[B] 1 on convergence.  work: qsp_big_jacobi_work_bytes(nrows, ldw, B) bytes of device memory.
extern "C" size_t qsp_big_jacobi_work_bytes(int nrows, int ldw, int B) {
  return align_up((size_t)B * sizeof(BigJac)) + align_up((size_t)(B + 1) * sizeof(int)) + big_jacobi_scratch_bytes(B, nrows, ldw);
}
extern "C" int qsp_big_jacobi(qsp_c128* W, int nrows, int ndotp, int ldw, int B, int* sweeps, int* ok, void* work, void* stream) {
  if (!W || !sweeps || !ok || !work || B <= 0 || nrows < 1 || ldw % QSP_BIG_CK || ndotp % QSP_BIG_CK || ndotp > ldw) return QSP_ERR_ARG;
  qsp_stream_t st = (qsp_stream_t)stream;
  BigJac* jd = (BigJac*)work;
  int* flags = (int*)((char*)work + align_up((size_t)B * sizeof(BigJac)));
  int* any = flags + B;
  void* jw = (char*)work + align_up((size_t)B * sizeof(BigJac)) + align_up((size_t)(B + 1) * sizeof(int));
  std::vector<BigJac> h(B);
  const double tol = 4e-16 * sqrt((double)ndotp);
  for (int b = 0; b < B; ++b) {
    memset(&h[b], 0, sizeof(BigJac));
    h[b].W = (cplx*)W + (long)b * nrows * ldw;
    h[b].nrows = nrows, h[b].ndotp = ndotp, h[b].len = ldw, h[b].ldw = ldw;
    h[b].nblk = nrows >= 2 ? (((nrows + QSP_BIG_RB - 1) / QSP_BIG_RB + 1) & ~1) : 0;
    h[b].inner = h[b].nblk == 2 ? 8 : 1;
    h[b].tol2 = tol * tol;
  }
#ifdef QSP_EMU
  memcpy(jd, h.data(), (size_t)B * sizeof(BigJac));
#else
  cudaMemcpyAsync(jd, h.data(), (size_t)B * sizeof(BigJac), cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
#endif
  dev_zero(sweeps, (size_t)B * sizeof(int), st);
  int rc = big_jacobi_run(jd, B, nrows, ldw, flags, any, sweeps, jw, st);
  if (rc) return rc;
#ifdef QSP_EMU
  for (int b = 0; b < B; ++b) ok[b] = jd[b].nblk == 0;
#else
  cudaMemcpyAsync(h.data(), jd, (size_t)B * sizeof(BigJac), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  std::vector<int> hok(B);
  for (int b = 0; b < B; ++b) hok[b] = h[b].nblk == 0;
  cudaMemcpyAsync(ok, hok.data(), (size_t)B * sizeof(int), cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
#endif
  return check_launch();
}

// Batched tensor-core GEMM of the large-bond path on its own: C[b] (m x n, row-major) = op(A[b]) op(B[b]) for B
// problems of one shape; A is m x k row-major (transA: stored k x m), B is k x n row-major (transB: stored n x k);
// conj flags conjugate the operand.  work: B * 128 bytes of device memory.
extern "C" int qsp_big_gemm(const qsp_c128* A, int transA, int conjA, const qsp_c128* Bm, int transB, int conjB, qsp_c128* C, int m,
                            int n, int k, int batch, void* work, void* stream) {
  if (!A || !Bm || !C || !work || m <= 0 || n <= 0 || k < 0 || batch <= 0) return QSP_ERR_ARG;
  qsp_stream_t st = (qsp_stream_t)stream;
  std::vector<BigGemm> h(batch);
  for (int b = 0; b < batch; ++b) {
    BigGemm g;
    memset(&g, 0, sizeof(g));
    g.A = (const cplx*)A + (long)b * m * k, g.B = (const cplx*)Bm + (long)b * k * n, g.C = (cplx*)C + (long)b * m * n;
    g.m = m, g.n = n, g.k = k;
    g.sAr = transA ? 1 : k, g.sAk = transA ? m : 1;
    g.sBk = transB ? 1 : n, g.sBc = transB ? k : 1;
    g.sCr = n, g.sCc = 1;
    g.conjA = conjA, g.conjB = conjB;
    h[b] = g;
  }
#ifdef QSP_EMU
  memcpy(work, h.data(), (size_t)batch * sizeof(BigGemm));
#else
  cudaMemcpyAsync(work, h.data(), (size_t)batch * sizeof(BigGemm), cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
#endif
  int rc = launch_big_gemm((const BigGemm*)work, batch, m, n, st);
  return rc ? rc : check_launch();
}

// tuning hook: column splits of the block-Jacobi update / Gram launches (0 = automatic)
extern "C" void qsp_big_tuning(int update_splits, int gram_splits) {
  g_big_splits = update_splits;
  g_big_gsplits = gram_splits;
}
