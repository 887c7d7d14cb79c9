// C ABI of libqsp_b200.so (include/qsp_b200.h): launch configuration, workspace
// partitioning and the batched layer loop of sequential_unitary_circuit
// (qsp/sequential/sequential.py:42-85).  Host-side only; all arithmetic is in the kernels.
#include "../../include/qsp_b200.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include "qsp_sweeps.cuh"
#include "qsp_big_sweeps.cuh"

// NVTX ranges per layer / sweep of the layer loop (header-only NVTX 3: a no-op unless a profiler is attached)
#ifndef QSP_EMU
#include <nvtx3/nvToolsExt.h>
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#else
struct NvtxRange {
  explicit NvtxRange(const char*) {}
};
#endif
#include "qsp_unitaries.cuh"
#include "qsp_circuit.cuh"

#ifdef QSP_EMU
thread_local emu::Block* emu::g_blk = nullptr;
thread_local dim3 emu::g_threadIdx, emu::g_blockIdx, emu::g_blockDim, emu::g_gridDim;
#endif

// Per-thread state only (the ABI is re-entrant per host thread and per stream): the last CUDA error string and
// the optional measurement buffer belong to the calling thread; the launch counter is an atomic tally.
#include <atomic>
#include <map>
#include <mutex>
static std::atomic<long long> g_launches{0};
static thread_local long long* g_prof = nullptr;  // optional device buffer [grid][8] for k_right_compress phase cycles
static thread_local char g_err[256] = "";

// ------------------------------------------------------------------------------------
// device facts
// ------------------------------------------------------------------------------------
struct DevInfo {
  int sms;
  size_t max_smem;  // per block, opt-in
};
static DevInfo dev_info() {
  DevInfo d;
#ifdef QSP_EMU
  d.sms = 2;
  d.max_smem = 227 * 1024;
#else
  int dev = 0, v = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  d.sms = v > 0 ? v : 148;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  d.max_smem = v > 0 ? (size_t)v : (size_t)227 * 1024;
#endif
  return d;
}

#ifdef QSP_EMU
#define QSP_PREP_KERNEL(kernel, smem) (0)
static int blocks_per_sm(const void*, int, size_t) { return 1; }
static int check_launch() { return QSP_OK; }
static void dev_fill_nan(qsp_c128* p, long n, qsp_stream_t) {
  for (long i = 0; i < n; ++i) p[i].re = p[i].im = NAN;
}
static void dev_copy(void* dst, const void* src, size_t bytes, qsp_stream_t) { memcpy(dst, src, bytes); }
static void dev_zero(void* dst, size_t bytes, qsp_stream_t) { memset(dst, 0, bytes); }
#else
// The dynamic shared-memory limit of a kernel is one attribute per device, shared by every host thread and stream: it is
// only ever RAISED (under a lock), so that a launch planned by one thread can never find the limit lowered by another
// thread that is running the same kernel with a smaller footprint (the two-stream chunk overlap runs two host threads).
static int prep_kernel(const void* kernel, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> limit;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = limit[std::make_pair(kernel, dev)];
  if (smem <= cur) return 0;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) cur = smem;
  return (int)e;
}
#define QSP_PREP_KERNEL(kernel, smem) prep_kernel((const void*)kernel, (size_t)(smem))
static int blocks_per_sm(const void* kernel, int nt, size_t smem) {
  int n = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, nt, smem);
  return n > 0 ? n : 1;
}
static int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s", cudaGetErrorString(e));
    return QSP_ERR_CUDA;
  }
  return QSP_OK;
}
__global__ void k_fill_nan(cplx* p, long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = cmake(nan(""), nan(""));
}
static void dev_fill_nan(qsp_c128* p, long n, qsp_stream_t s) {
  if (n <= 0) return;
  k_fill_nan<<<(unsigned)((n + 255) / 256), 256, 0, s>>>((cplx*)p, n);
  ++g_launches;
}
static void dev_copy(void* dst, const void* src, size_t bytes, qsp_stream_t s) {
  cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s);
}
static void dev_zero(void* dst, size_t bytes, qsp_stream_t s) { cudaMemsetAsync(dst, 0, bytes, s); }
#endif

#ifdef QSP_EMU
static int dev_read_int(const int* p, qsp_stream_t) { return *p; }
#else
static int dev_read_int(const int* p, qsp_stream_t s) {  // synchronises the stream
  int v = 0;
  cudaMemcpyAsync(&v, p, sizeof(int), cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  return v;
}
#endif

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
static int grid_for(int B, int per_sm, const DevInfo& d) {
  long g = (long)d.sms * per_sm;
#ifdef QSP_EMU
  g = 2;
#endif
  return (int)(B < g ? B : g);
}

// ------------------------------------------------------------------------------------
// launch plans
// ------------------------------------------------------------------------------------
struct LCanonPlan {
  int nt;
  size_t smem;
  bool big;  // bonds beyond one SM's shared memory: batched blocked-Householder path (big_left_canonize)
};
#ifdef QSP_DEV_BUILD  // libqsp_b200_dev.so / the test emulation only (include/qsp_b200_dev.h)
static int g_force_big = 0;  // test hook (qsp_debug_flags bit 0): take the large-bond path at any size
static int g_nt_rc = 0, g_nt_lc = 0;  // tuning hooks: threads per CTA of k_right_compress / k_left_canon (0 = planned)
static int g_enc_cap = 0;             // test hook: padded bond of the encoded-state buffers (0 = QSP_ENC_CAP)
static int g_nosort = 0;              // A/B hook (qsp_debug_flags bit 2): no norm ordering of the Jacobi rows
static int g_tmem = 0;                // experiment (qsp_debug_flags bit 1): qsp_jacobi_bench keeps 128 x 128 rows in tensor memory
#else
static const int g_force_big = 0, g_nt_rc = 0, g_nt_lc = 0, g_enc_cap = 0, g_nosort = 0;
#endif
static bool plan_lcanon(int D, const DevInfo& d, LCanonPlan* p) {
  p->big = false;
  p->nt = 256;
  p->smem = 0;
  if (2 * D <= 32 * QSP_QR_MAXR && !g_force_big) {  // a column must fit one warp's registers
    p->nt = D <= 8 ? 64 : (D <= 24 ? 128 : (D <= 48 ? 256 : 512));  // measured at chi = 8, 16, 32, 64
    if (g_nt_lc) p->nt = g_nt_lc;
    p->smem = lcanon_smem_bytes(D, p->nt);
    if (p->smem > d.max_smem && p->nt == 512) {
      p->nt = 256;
      p->smem = lcanon_smem_bytes(D, p->nt);
    }
    if (p->smem <= d.max_smem) return true;
  }
  p->big = true;  // batched blocked-Householder path (qsp_big_sweeps.cuh): any bond
  return true;
}
static int lcanon_grid(int B, const LCanonPlan& p, const DevInfo& d) {
  return grid_for(B, blocks_per_sm((const void*)k_left_canon, p.nt, p.smem), d);
}

struct RCompressPlan {
  int nt, tpp;
  int jac_elems;
  size_t smem;
};
static bool plan_rcompress(int Din, int Dout, const DevInfo& d, RCompressPlan* p) {
  const int dm = imax(Din / 2, Dout);
  // eight lanes per Jacobi pair (2 dm rows -> dm pairs): the power-of-two bonds then hit the compile-time
  // specialisations of the register-resident Jacobi (chi = 16 / 32 / 64 -> 128 / 256 / 512 threads)
  p->nt = dm >= 48 ? 512 : (dm >= 24 ? 256 : (dm >= 4 ? 128 : 64));  // measured: chi 8..16 -> 128, 24..32 -> 256, 48.. -> 512
  if (g_nt_rc) p->nt = g_nt_rc;
  int tpp = 32;
  while (tpp > 2 && p->nt / tpp < imin(Dout, Din)) tpp >>= 1;  // >= one team per pair when possible
  p->tpp = tpp;
  // largest Jacobi working set (rows padded to the team width, <= 32 columns of zeros):
  // mode R (b <= c): 2Dout x 2Dout; mode C (c < b): c x (b + c)
  auto r32 = [](long x) { return (x + 31) / 32 * 32; };
  long needR = 2L * Dout * r32(2L * Dout);
  long cmax = imin(2 * Dout, Din - 1);
  long needC = cmax > 0 ? cmax * (r32(Din) + r32(cmax)) : 0;
  long need = needR > needC ? needR : needC;
  size_t fixed = rcompress_smem_bytes(0, Dout, p->nt) - qsp_gemm_smem_bytes(p->nt, QSP_GEMM_TK_DEEP);
  long budget = ((long)d.max_smem - (long)fixed) / (long)sizeof(cplx);
  long gem = (long)(qsp_gemm_smem_bytes(p->nt, QSP_GEMM_TK_DEEP) / sizeof(cplx));
  if (budget < gem) return false;
  // at least two rows of the longest possible row must fit
  long maxlen = imax((int)r32(2 * Dout), (int)(r32(Din) + r32(cmax)));
  if (need < 2 * maxlen) need = 2 * maxlen;
  long je = need < budget ? need : budget;
  if (je < 2 * maxlen) return false;
  p->jac_elems = (int)je;
  p->smem = rcompress_smem_bytes(p->jac_elems, Dout, p->nt);
  return p->smem <= d.max_smem;
}

// ------------------------------------------------------------------------------------
// unit-level entry points
// ------------------------------------------------------------------------------------
extern "C" int qsp_version(void) { return 1; }
extern "C" int qsp_is_device_build(void) {
#ifdef QSP_EMU
  return 0;
#else
  return 1;
#endif
}
extern "C" const char* qsp_last_cuda_error(void) { return g_err; }
extern "C" void qsp_set_profile_buffer(void* p) { g_prof = (long long*)p; }
extern "C" long long qsp_launch_count(void) { return g_launches; }

// Dlive: largest bond that can occur in this call (<= D, the padded stride); the launch (threads,
// shared memory, CTAs per SM) is planned for Dlive, a larger live bond sets QSP_ST_BOND_OVERFLOW.
// ------------------------------------------------------------------------------------
// large-bond left canonisation (qsp_big_sweeps.cuh): blocked Householder QR, batched launches per site
// ------------------------------------------------------------------------------------
struct BigLCLayout {
  size_t off_S, off_g2, off_gW, off_gC, off_work, total;
  long stride;
};
static BigLCLayout big_lc_layout(int B, int Dm) {
  BigLCLayout l;
  size_t o = 0;
  l.off_S = o, o += align_up((size_t)B * sizeof(BigLCState));
  l.off_g2 = o, o += align_up((size_t)2 * B * sizeof(BigGemm));
  l.off_gW = o, o += align_up((size_t)B * sizeof(BigGemm));
  l.off_gC = o, o += align_up((size_t)B * sizeof(BigGemm));
  l.off_work = o;
  l.stride = big_lc_work_elems(Dm);
  l.total = o + align_up((size_t)B * l.stride * sizeof(cplx));
  return l;
}
static int launch_big_gemm(const BigGemm* gd, int ndesc, int m_max, int n_max, qsp_stream_t st);
static int big_left_canonize(qsp_c128* mps, int32_t* bonds, int B, int L, int D, int Dm, int32_t* status, void* work,
                             size_t work_bytes, void* stream) {
  BigLCLayout lay = big_lc_layout(B, Dm);
  if (!work || work_bytes < lay.total) return QSP_ERR_WORKSPACE;
  char* w = (char*)work;
  BigLCState* S = (BigLCState*)(w + lay.off_S);
  BigGemm* g2 = (BigGemm*)(w + lay.off_g2);
  BigGemm* gW = (BigGemm*)(w + lay.off_gW);
  BigGemm* gC = (BigGemm*)(w + lay.off_gC);
  BigLCArgs a;
  a.mps = (cplx*)mps, a.bonds = bonds, a.B = B, a.L = L, a.D = D, a.Dm = Dm, a.status = status;
  a.work = (cplx*)(w + lay.off_work), a.work_stride = lay.stride;
  qsp_stream_t st = (qsp_stream_t)stream;
  const int nb1 = (B + 127) / 128;
  const size_t psm = 64 * sizeof(double) + 2 * QSP_BIG_NBW * QSP_BIG_NBW * sizeof(cplx) + 64;
  QSP_LAUNCH(k_big_lc_init, nb1, 128, 0, st, a, S);
  ++g_launches;
  int rc;
  for (int i = 0; i < L; ++i) {
    // upper bounds: the new left bond of site i is at most 2^i
    const int a_max = i >= 30 ? Dm : imin(Dm, 1 << i);
    const int m_max = 2 * a_max, n_max = Dm, K_max = imin(m_max, n_max);
    const int parts = imax(1, imin(64, (int)(((long)m_max * n_max + 4095) / 4096)));
    QSP_LAUNCH(k_big_lc_site, nb1, 128, 0, st, a, S, i, g2);
    ++g_launches;
    if (i == 0) {
      QSP_LAUNCH(k_big_lc_elem, B * parts, 256, 0, st, a, S, i, 0, parts);
      ++g_launches;
    } else if ((rc = launch_big_gemm(g2, 2 * B, a_max, n_max, st))) {
      return rc;
    }
    if (i == L - 1) {
      QSP_LAUNCH(k_big_lc_elem, B * parts, 256, 0, st, a, S, i, 1, parts);
      ++g_launches;
      break;
    }
    for (int p0 = 0; p0 < K_max; p0 += QSP_BIG_NBW) {
      QSP_LAUNCH(k_big_qr_panel, B, 256, psm, st, a, (const BigLCState*)S, p0, gW, gC);
      ++g_launches;
      if (n_max - p0 - 1 > 0) {  // (a wide site, K < n, has trailing columns behind a short last panel)
        if ((rc = launch_big_gemm(gW, B, QSP_BIG_NBW, n_max - p0 - 1, st))) return rc;
        if ((rc = launch_big_gemm(gC, B, m_max - p0, n_max - p0 - 1, st))) return rc;
      }
    }
    QSP_LAUNCH(k_big_lc_elem, B * parts, 256, 0, st, a, S, i, 2, parts);
    ++g_launches;
    for (int p0 = (K_max - 1) / QSP_BIG_NBW * QSP_BIG_NBW; p0 >= 0; p0 -= QSP_BIG_NBW) {
      QSP_LAUNCH(k_big_lc_qdesc, nb1, 128, 0, st, a, (const BigLCState*)S, p0, gW, gC);
      ++g_launches;
      if ((rc = launch_big_gemm(gW, B, QSP_BIG_NBW, K_max - p0, st))) return rc;
      if ((rc = launch_big_gemm(gC, B, m_max - p0, K_max - p0, st))) return rc;
    }
    QSP_LAUNCH(k_big_lc_elem, B * parts, 256, 0, st, a, S, i, 3, parts);
    ++g_launches;
  }
  return check_launch();
}

// Dlive: largest bond that can occur in this call (<= D, the padded stride); the launch (threads,
// shared memory, CTAs per SM) is planned for Dlive, a larger live bond sets QSP_ST_BOND_OVERFLOW.
extern "C" size_t qsp_left_canonize_workspace_bytes(int B, int L, int D) {
  (void)L;
  DevInfo d = dev_info();
  LCanonPlan p;
  if (B <= 0 || D <= 0 || !plan_lcanon(D, d, &p) || !p.big) return 0;
  return big_lc_layout(B, D).total;
}
static int left_canonize_live(qsp_c128* mps, int32_t* bonds, int B, int L, int D, int Dlive, int32_t* status,
                              void* work, size_t work_bytes, void* stream) {
  if (!mps || !bonds || B <= 0 || L <= 0 || D <= 0) return QSP_ERR_ARG;
  if (Dlive > D || Dlive <= 0) Dlive = D;
  DevInfo d = dev_info();
  LCanonPlan p;
  if (!plan_lcanon(Dlive, d, &p)) return QSP_ERR_UNSUPPORTED;
  if (p.big) return big_left_canonize(mps, bonds, B, L, D, Dlive, status, work, work_bytes, stream);
  QSP_PREP_KERNEL(k_left_canon, p.smem);
  int grid = lcanon_grid(B, p, d);
  QSP_LAUNCH(k_left_canon, grid, p.nt, p.smem, (qsp_stream_t)stream, (cplx*)mps, bonds, B, L, D, Dlive, status, g_prof);
  ++g_launches;
  return check_launch();
}
extern "C" int qsp_left_canonize(qsp_c128* mps, int32_t* bonds, int B, int L, int D, int32_t* status, void* work,
                                 size_t work_bytes, void* stream) {
  return left_canonize_live(mps, bonds, B, L, D, D, status, work, work_bytes, stream);
}
#ifdef QSP_DEV_BUILD
extern "C" void qsp_debug_flags(int flags) {
  g_force_big = flags & 1;
  g_tmem = (flags >> 1) & 1;
  g_nosort = (flags >> 2) & 1;
  g_enc_cap = ((flags >> 4) & 0xf) ? 1 << ((flags >> 4) & 0xf) : 0;  // bits 4..7: log2 of the encoded-state buffer bond
  g_nt_rc = (flags >> 8) & 0xfff;   // bits 8..19: threads per CTA of k_right_compress
  g_nt_lc = (flags >> 20) & 0xfff;  // bits 20..31: threads per CTA of k_left_canon
}
#endif

// ------------------------------------------------------------------------------------
// large-bond right compression (qsp_big_sweeps.cuh): batched launches per site
// ------------------------------------------------------------------------------------
#define QSP_BIG_MAX_SWEEPS 60
// matrices beyond 128 x 128 no longer fit the register + shared memory of one SM
static bool rcompress_is_big(int DinL, int DoutL) { return g_force_big || imax(DinL / 2, DoutL) > 64; }
struct BigRCLayout {
  size_t off_S, off_jd, off_g1, off_g2, off_gU, off_flags, off_any, off_sweeps, off_jw, off_work, total;
  long stride;
};
static size_t big_jacobi_scratch_bytes(int B, int nrows_max, int len_max);
static BigRCLayout big_rc_layout(int B, int DinL, int DoutL) {
  BigRCLayout l;
  size_t o = 0;
  l.off_S = o, o += align_up((size_t)B * sizeof(BigRCState));
  l.off_jd = o, o += align_up((size_t)B * sizeof(BigJac));
  l.off_g1 = o, o += align_up((size_t)B * sizeof(BigGemm));
  l.off_g2 = o, o += align_up((size_t)2 * B * sizeof(BigGemm));
  l.off_gU = o, o += align_up((size_t)B * sizeof(BigGemm));
  l.off_flags = o, o += align_up((size_t)B * sizeof(int));
  l.off_any = o, o += 256;
  l.off_sweeps = o, o += align_up((size_t)B * sizeof(int));
  l.off_jw = o, o += big_jacobi_scratch_bytes(B, 2 * DoutL, DinL + 2 * DoutL + 64);
  l.off_work = o;
  l.stride = rcompress_work_elems(DinL, DoutL);
  l.total = o + align_up((size_t)B * l.stride * sizeof(cplx));
  return l;
}
static int launch_big_gemm(const BigGemm* gd, int ndesc, int m_max, int n_max, qsp_stream_t st) {
  if (ndesc <= 0 || m_max <= 0 || n_max <= 0) return QSP_OK;
  const int tm = (m_max + QSP_GB_TM - 1) / QSP_GB_TM, tn = (n_max + QSP_GB_TN - 1) / QSP_GB_TN;
  QSP_PREP_KERNEL(k_big_gemm, big_gemm_smem_bytes());
  QSP_LAUNCH(k_big_gemm, ndesc * tm * tn, QSP_BIG_NT, big_gemm_smem_bytes(), st, gd, tm, tn);
  ++g_launches;
  return QSP_OK;
}
// block Jacobi SVD of the B matrices described by jd (device); nrows_max / len_max bound every instance's extents.
// jw: scratch of big_jacobi_scratch_bytes(B, nrows_max, len_max) (partial Gram matrices, pair unitaries, skip flags)
#ifdef QSP_DEV_BUILD
static int g_big_splits = 0, g_big_gsplits = 0;  // tuning hooks (qsp_big_tuning)
#else
static const int g_big_splits = 0, g_big_gsplits = 0;
#endif
// Column splits of the tensor-core launches of a round: enough CTAs to keep 148 SMs two deep (about 1024 update CTAs, 512
// Gram CTAs -- the Gram launch pays a cross-warp reduction per CTA), at least two chunks per CTA.  Measured at 32 x (512 x 512):
// (update, gram) = (2, 1) 82.9 ms, (4, 2) 85.8 ms, (1, 1) 87.8 ms, (8, 2) 93.4 ms.
static void big_jacobi_splits(int B, int nrows_max, int len_max, int* usplits, int* gsplits) {
  const int nblk = ((nrows_max + QSP_BIG_RB - 1) / QSP_BIG_RB + 1) & ~1;
  const long tasks = (long)B * (nblk / 2);
  const int cap = imax(1, len_max / (2 * QSP_BIG_CK));
  *usplits = g_big_splits ? g_big_splits : imax(1, imin(imin(8, cap), (int)((1024 + tasks - 1) / tasks)));
  *gsplits = g_big_gsplits ? g_big_gsplits : imax(1, imin(imin(4, cap), (int)((512 + tasks - 1) / tasks)));
}
static size_t big_jacobi_scratch_bytes(int B, int nrows_max, int len_max) {
  const int nblk = ((nrows_max + QSP_BIG_RB - 1) / QSP_BIG_RB + 1) & ~1;
  const size_t tasks = (size_t)B * (nblk / 2), tile = (size_t)QSP_BIG_PR * QSP_BIG_PR * sizeof(cplx);
  return align_up(tasks * 4 * tile) + align_up(tasks * tile) + align_up(tasks * sizeof(int));  // (up to four Gram partials)
}
static int big_jacobi_run(BigJac* jd, int B, int nrows_max, int len_max, int* flags, int* any, int* sweeps, void* jw, qsp_stream_t st) {
  if (nrows_max < 2) return QSP_OK;
  const int nblk = ((nrows_max + QSP_BIG_RB - 1) / QSP_BIG_RB + 1) & ~1;
  const int pairs = nblk / 2, rounds = nblk - 1;
  int splits, gsplits;
  big_jacobi_splits(B, nrows_max, len_max, &splits, &gsplits);
  gsplits = imin(gsplits, 4);
  const size_t tasks = (size_t)B * pairs, tile = (size_t)QSP_BIG_PR * QSP_BIG_PR * sizeof(cplx);
  cplx* Gp = (cplx*)jw;
  cplx* Jg = (cplx*)((char*)jw + align_up(tasks * 4 * tile));
  int* skip = (int*)((char*)Jg + align_up(tasks * tile));
  QSP_PREP_KERNEL(k_big_jac_gram, big_gram_smem_bytes());
  QSP_PREP_KERNEL(k_big_jac_solve, big_solve_smem_bytes());
  QSP_PREP_KERNEL(k_big_jac_update, big_update_smem_bytes());
  dev_zero(flags, (size_t)B * sizeof(int), st);
  for (int sw = 0; sw < QSP_BIG_MAX_SWEEPS; ++sw) {
    dev_zero(any, sizeof(int), st);
    for (int r = 0; r < rounds; ++r) {
      QSP_LAUNCH(k_big_jac_gram, (int)tasks * gsplits, QSP_BIG_NT, big_gram_smem_bytes(), st, (const BigJac*)jd, r, pairs, gsplits, Gp);
      QSP_LAUNCH(k_big_jac_solve, (int)tasks, QSP_BIG_NT, big_solve_smem_bytes(), st, (const BigJac*)jd, flags, r, pairs, gsplits,
                 (const cplx*)Gp, Jg, skip);
      QSP_LAUNCH(k_big_jac_update, (int)tasks * splits, QSP_BIG_NT, big_update_smem_bytes(), st, (const BigJac*)jd, r, pairs, splits,
                 (const cplx*)Jg, (const int*)skip);
      g_launches += 3;
    }
    QSP_LAUNCH(k_big_sweep_end, (B + 127) / 128, 128, 0, st, jd, flags, any, sweeps, B);
    ++g_launches;
    if (dev_read_int(any, st) == 0) break;  // every instance converged (host sync once per sweep)
  }
  return check_launch();
}

static int big_right_compress(const RCompressArgs& a0, void* work, size_t work_bytes, void* stream) {
  RCompressArgs a = a0;
  const int B = a.B, L = a.L;
  BigRCLayout lay = big_rc_layout(B, a.DinL, a.DoutL);
  if (work_bytes < lay.total) return QSP_ERR_WORKSPACE;
  char* w = (char*)work;
  BigRCState* S = (BigRCState*)(w + lay.off_S);
  BigJac* jd = (BigJac*)(w + lay.off_jd);
  BigGemm* g1 = (BigGemm*)(w + lay.off_g1);
  BigGemm* g2 = (BigGemm*)(w + lay.off_g2);
  BigGemm* gU = (BigGemm*)(w + lay.off_gU);
  int* flags = (int*)(w + lay.off_flags);
  int* any = (int*)(w + lay.off_any);
  int* sweeps = (int*)(w + lay.off_sweeps);
  a.work = (cplx*)(w + lay.off_work);
  a.work_stride = lay.stride;
  a.queue = nullptr;
  qsp_stream_t st = (qsp_stream_t)stream;
  const bool fused = a.G != nullptr;
  const int kmax = (a.max_bond > 0) ? imin(a.max_bond, a.DoutL) : a.DoutL;
  const int nb1 = (B + 127) / 128;
  QSP_LAUNCH(k_big_rc_init, nb1, 128, 0, st, a, S);
  ++g_launches;
  auto pow2cap = [](int e, int cap) { return e >= 30 ? cap : imin(cap, 1 << e); };
  int rc;
  for (int i = L - 1; i >= 0; --i) {
    // upper bounds of this site's extents: a left-canonical input has bond j <= 2^j, the truncated bond j <= 2^(L-j)
    const int rp_max = pow2cap(L - 1 - i, kmax);                       // truncated bond i+1
    const int b_max = pow2cap(i, a.DinL);                              // left extent of the swept state
    const int c_max = 2 * rp_max;
    const int lj_max = i > 0 ? pow2cap(i - 1, a.Din) : 1;
    const int parts = imax(1, imin(64, (int)(((long)b_max * c_max + 4095) / 4096)));
    QSP_LAUNCH(k_big_rc_site, nb1, 128, 0, st, a, S, i, g1, g2);
    QSP_LAUNCH(k_big_rc_elem, B * parts, 256, 0, st, a, (const BigRCState*)S, i, 0, parts);
    g_launches += 2;
    if ((rc = launch_big_gemm(g1, B, 2 * imin(b_max, a.Din), rp_max, st))) return rc;
    if (fused) {
      QSP_LAUNCH(k_big_rc_elem, B * parts, 256, 0, st, a, (const BigRCState*)S, i, 1, parts);
      ++g_launches;
      if (i > 0) {
        if ((rc = launch_big_gemm(g2, 2 * B, 2 * lj_max, rp_max, st))) return rc;
        QSP_LAUNCH(k_big_rc_elem, B * parts, 256, 0, st, a, (const BigRCState*)S, i, 2, parts);
        ++g_launches;
      }
    }
    if (i == 0) {
      QSP_LAUNCH(k_big_rc_final, B, 256, 64 * sizeof(double) + 64, st, a, S);
      ++g_launches;
      break;
    }
    const int nrows_max = imin(b_max, c_max);
    QSP_LAUNCH(k_big_rc_jprep, B * parts, 256, 0, st, a, (const BigRCState*)S, jd, parts);
    ++g_launches;
    const int len_max = big_r32(imax(b_max, c_max)) + (b_max > c_max ? big_r32(c_max) : 0);
    if ((rc = big_jacobi_run(jd, B, nrows_max, len_max, flags, any, sweeps, w + lay.off_jw, st))) return rc;
    const size_t rsm = big_rank_smem_bytes(nrows_max);
    QSP_PREP_KERNEL(k_big_rc_rank, rsm);
    QSP_LAUNCH(k_big_rc_rank, B, 256, rsm, st, a, S, (const BigJac*)jd, i, gU);
    ++g_launches;
    if ((rc = launch_big_gemm(gU, B, b_max, imin(kmax, nrows_max), st))) return rc;
  }
  return check_launch();
}

static int rcompress_grid(int B, const RCompressPlan& p, const DevInfo& d) {
  return grid_for(B, blocks_per_sm((const void*)k_right_compress, p.nt, p.smem), d);
}

extern "C" size_t qsp_right_compress_workspace_bytes(int B, int L, int D_in, int D_out) {
  (void)L;
  DevInfo d = dev_info();
  RCompressPlan p;
  if (B <= 0 || D_in <= 0 || D_out <= 0) return 0;
  if (rcompress_is_big(D_in, D_out)) return big_rc_layout(B, D_in, D_out).total;
  if (!plan_rcompress(D_in, D_out, d, &p)) return 0;
#ifndef QSP_EMU
  QSP_PREP_KERNEL(k_right_compress, p.smem);
#endif
  int grid = rcompress_grid(B, p, d);
  return align_up(256 + (size_t)grid * rcompress_work_elems(D_in, D_out) * sizeof(cplx));
}

static int right_compress_live(const qsp_c128* in, const int32_t* bonds_in, int D_in, qsp_c128* out,
                               int32_t* bonds_out, int D_out, int DinL, int DoutL, int B, int L, int max_bond,
                               double cutoff, int normalize, double* svals, int nsv, qsp_c128* amp0, int32_t* status,
                               void* work, size_t work_bytes, void* stream, const qsp_c128* G = nullptr,
                               const int32_t* kind = nullptr) {
  if (!in || !bonds_in || !out || !bonds_out || !work || B <= 0 || L <= 0 || D_in <= 0 || D_out <= 0)
    return QSP_ERR_ARG;
  const bool fused = G != nullptr;
  if (fused && (!kind || D_in > D_out || in == out || bonds_in == bonds_out)) return QSP_ERR_ARG;
  if (fused) {
    if (DinL <= 0) DinL = 2 * D_in;  // live bonds of U^dagger psi
  } else if (DinL <= 0 || DinL > D_in) {
    DinL = D_in;
  }
  if (DoutL <= 0 || DoutL > D_out) DoutL = D_out;
  if (rcompress_is_big(DinL, DoutL)) {
    RCompressArgs a;
    memset(&a, 0, sizeof(a));
    a.in = (const cplx*)in, a.bonds_in = bonds_in, a.Din = D_in;
    a.out = (cplx*)out, a.bonds_out = bonds_out, a.Dout = D_out;
    a.DinL = DinL, a.DoutL = DoutL, a.B = B, a.L = L;
    a.max_bond = max_bond, a.cutoff = cutoff, a.normalize = normalize;
    a.svals = svals, a.nsv = nsv, a.amp0 = (cplx*)amp0, a.status = status;
    a.G = (const cplx*)G, a.kind = kind;
    return big_right_compress(a, work, work_bytes, stream);
  }
  DevInfo d = dev_info();
  RCompressPlan p;
  if (!plan_rcompress(DinL, DoutL, d, &p)) return QSP_ERR_UNSUPPORTED;
  QSP_PREP_KERNEL(k_right_compress, p.smem);
  int grid = rcompress_grid(B, p, d);
  long stride = rcompress_work_elems(DinL, DoutL);
  if (work_bytes < 256 + (size_t)grid * stride * sizeof(cplx)) return QSP_ERR_WORKSPACE;
  dev_zero(work, 256, (qsp_stream_t)stream);  // instance queue head
  RCompressArgs a;
  a.in = (const cplx*)in;
  a.bonds_in = bonds_in;
  a.Din = D_in;
  a.out = (cplx*)out;
  a.bonds_out = bonds_out;
  a.Dout = D_out;
  a.DinL = DinL;
  a.DoutL = DoutL;
  a.B = B;
  a.L = L;
  a.max_bond = max_bond;
  a.cutoff = cutoff;
  a.normalize = normalize;
  a.svals = svals;
  a.nsv = nsv;
  a.amp0 = (cplx*)amp0;
  a.status = status;
  a.work = (cplx*)((char*)work + 256);
  a.work_stride = stride;
  a.queue = (int*)work;
  a.G = (const cplx*)G;
  a.kind = kind;
  a.cap_rows_elems = p.jac_elems;
  a.tpp = p.tpp;
  a.max_sweeps = 60;
  a.prof = g_prof;
  a.presort = g_nosort ? 0 : 1;
  QSP_LAUNCH(k_right_compress, grid, p.nt, p.smem, (qsp_stream_t)stream, a);
  ++g_launches;
  return check_launch();
}
extern "C" int qsp_right_compress(const qsp_c128* in, const int32_t* bonds_in, int D_in, qsp_c128* out,
                                  int32_t* bonds_out, int D_out, int B, int L, int max_bond, double cutoff,
                                  int normalize, double* svals, int nsv, qsp_c128* amp0, int32_t* status, void* work,
                                  size_t work_bytes, void* stream) {
  return right_compress_live(in, bonds_in, D_in, out, bonds_out, D_out, D_in, D_out, B, L, max_bond, cutoff, normalize,
                             svals, nsv, amp0, status, work, work_bytes, stream);
}

extern "C" size_t qsp_apply_inverse_and_compress_workspace_bytes(int B, int L, int D) {
  return qsp_right_compress_workspace_bytes(B, L, 2 * D, D);
}
extern "C" int qsp_apply_inverse_and_compress(const qsp_c128* in, const int32_t* bonds_in, const qsp_c128* G,
                                              const int32_t* kind, qsp_c128* out, int32_t* bonds_out, int D, int B,
                                              int L, int max_bond, double cutoff, int normalize, qsp_c128* amp0,
                                              int32_t* status, void* work, size_t work_bytes, void* stream) {
  if (!G || !kind) return QSP_ERR_ARG;
  return right_compress_live(in, bonds_in, D, out, bonds_out, D, 2 * D, D, B, L, max_bond, cutoff, normalize, nullptr, 0,
                             amp0, status, work, work_bytes, stream, G, kind);
}

extern "C" int qsp_generate_unitaries(qsp_c128* d2, const int32_t* d2_bonds, int B, int L, qsp_c128* G,
                                      int32_t* kind, int32_t* status, void* stream) {
  if (!d2 || !d2_bonds || !G || !kind || B <= 0 || L <= 0) return QSP_ERR_ARG;
  const int nt = 64;
  QSP_LAUNCH(k_unitaries, (B + nt - 1) / nt, nt, 0, (qsp_stream_t)stream, (cplx*)d2, d2_bonds, B, L, (cplx*)G, kind,
             status);
  ++g_launches;
  return check_launch();
}

extern "C" int qsp_apply_inverse_layer(const qsp_c128* in, const int32_t* bonds_in, int D_in, const qsp_c128* G,
                                       const int32_t* kind, qsp_c128* out, int32_t* bonds_out, int D_out, int B,
                                       int L, int32_t* status, void* stream) {
  if (!in || !bonds_in || !G || !kind || !out || !bonds_out || B <= 0 || L <= 0) return QSP_ERR_ARG;
  DevInfo d = dev_info();
  int grid = grid_for(B, 4, d);
  QSP_LAUNCH(k_expand_inverse, grid, 256, 0, (qsp_stream_t)stream, (const cplx*)in, bonds_in, D_in, (const cplx*)G,
             kind, (cplx*)out, bonds_out, D_out, B, L, status);
  ++g_launches;
  return check_launch();
}

extern "C" int qsp_mps_axpby(const qsp_c128* x, const int32_t* bonds_x, int D_x, const qsp_c128* y, const int32_t* bonds_y,
                             int D_y, const qsp_c128* alpha, const qsp_c128* beta, qsp_c128* out, int32_t* bonds_out,
                             int D_out, int B, int L, int32_t* status, void* stream) {
  if (!x || !bonds_x || !y || !bonds_y || !alpha || !beta || !out || !bonds_out || B <= 0 || L <= 0) return QSP_ERR_ARG;
  if (out == x || out == y) return QSP_ERR_ARG;
  DevInfo d = dev_info();
  QSP_LAUNCH(k_mps_axpby, grid_for(B, 4, d), 256, 0, (qsp_stream_t)stream, (const cplx*)x, bonds_x, D_x, (const cplx*)y,
             bonds_y, D_y, (const cplx*)alpha, (const cplx*)beta, (cplx*)out, bonds_out, D_out, B, L, status);
  ++g_launches;
  return check_launch();
}
extern "C" int qsp_mps_scale(qsp_c128* x, const int32_t* bonds_x, int D_x, const qsp_c128* alpha, int B, int L,
                             void* stream) {
  if (!x || !bonds_x || !alpha || B <= 0 || L <= 0) return QSP_ERR_ARG;
  DevInfo d = dev_info();
  QSP_LAUNCH(k_mps_scale, grid_for(B, 4, d), 128, 0, (qsp_stream_t)stream, (cplx*)x, bonds_x, D_x, (const cplx*)alpha, B, L);
  ++g_launches;
  return check_launch();
}

extern "C" int qsp_apply_forward_layer(const qsp_c128* in, const int32_t* bonds_in, int D_in, const qsp_c128* G,
                                       const int32_t* kind, qsp_c128* out, int32_t* bonds_out, int D_out, int B,
                                       int L, int32_t* status, void* stream) {
  if (!in || !bonds_in || !G || !kind || !out || !bonds_out || B <= 0 || L <= 0) return QSP_ERR_ARG;
  DevInfo d = dev_info();
  int grid = grid_for(B, 4, d);
  QSP_LAUNCH(k_expand_forward, grid, 256, 0, (qsp_stream_t)stream, (const cplx*)in, bonds_in, D_in, (const cplx*)G,
             kind, (cplx*)out, bonds_out, D_out, B, L, status);
  ++g_launches;
  return check_launch();
}

static const int kOverlapNT = 256;
static bool overlap_is_big(int D_a, int D_b) { return g_force_big || imax(D_a, D_b) > 64; }
static size_t big_overlap_bytes(int B, int D_a, int D_b) {
  return align_up((size_t)3 * B * sizeof(BigGemm)) + align_up((size_t)B * big_ov_work_elems(D_a, D_b) * sizeof(cplx));
}
extern "C" size_t qsp_overlap_workspace_bytes(int B, int L, int D_a, int D_b) {
  (void)L;
  if (B > 0 && overlap_is_big(D_a, D_b)) return big_overlap_bytes(B, D_a, D_b);
  DevInfo d = dev_info();
  int grid = grid_for(B, 2, d);
  return align_up((size_t)grid * overlap_work_elems(D_a, D_b) * sizeof(cplx));
}
extern "C" int qsp_overlap(const qsp_c128* a, const int32_t* bonds_a, int D_a, const qsp_c128* b,
                           const int32_t* bonds_b, int D_b, int B, int L, qsp_c128* result, void* work,
                           size_t work_bytes, void* stream) {
  if (!a || !bonds_a || !b || !bonds_b || !result || !work || B <= 0 || L <= 0) return QSP_ERR_ARG;
  if (overlap_is_big(D_a, D_b)) {  // environments as batched tensor-core GEMMs (qsp_big_sweeps.cuh)
    if (work_bytes < big_overlap_bytes(B, D_a, D_b)) return QSP_ERR_WORKSPACE;
    BigGemm* gT = (BigGemm*)work;
    BigGemm* gE = gT + 2 * (size_t)B;
    BigOvArgs P;
    P.a = (const cplx*)a, P.bonds_a = bonds_a, P.Da = D_a, P.b = (const cplx*)b, P.bonds_b = bonds_b, P.Db = D_b;
    P.B = B, P.L = L, P.result = (cplx*)result;
    P.work = (cplx*)((char*)work + align_up((size_t)3 * B * sizeof(BigGemm)));
    P.work_stride = big_ov_work_elems(D_a, D_b);
    qsp_stream_t st = (qsp_stream_t)stream;
    int rc;
    for (int i = 0; i < L; ++i) {
      // upper bounds of the extents: bond j <= min(D, 2^j, 2^(L-j)) cannot be assumed for an arbitrary input: use D
      QSP_LAUNCH(k_big_ov_site, (B + 127) / 128, 128, 0, st, P, i, gT, gE);
      ++g_launches;
      if ((rc = launch_big_gemm(gT, 2 * B, D_a, D_b, st))) return rc;
      if ((rc = launch_big_gemm(gE, B, D_a, D_b, st))) return rc;
    }
    QSP_LAUNCH(k_big_ov_done, (B + 127) / 128, 128, 0, st, P);
    ++g_launches;
    return check_launch();
  }
  DevInfo d = dev_info();
  int grid = grid_for(B, 2, d);
  long stride = overlap_work_elems(D_a, D_b);
  if (work_bytes < (size_t)grid * stride * sizeof(cplx)) return QSP_ERR_WORKSPACE;
  size_t smem = qsp_gemm_smem_bytes(kOverlapNT, QSP_GEMM_TK_DEEP) + 64;  // both products on the FP64 tensor cores (cta_zgemm_deep)
  QSP_LAUNCH(k_overlap, grid, kOverlapNT, smem, (qsp_stream_t)stream, (const cplx*)a, bonds_a, D_a, (const cplx*)b,
             bonds_b, D_b, B, L, (cplx*)result, (cplx*)work, stride);
  ++g_launches;
  return check_launch();
}

// ------------------------------------------------------------------------------------
// statevector of the emitted circuit (qiskit_circuit.py:34-70)
// ------------------------------------------------------------------------------------
extern "C" int qsp_statevector_apply(qsp_c128* state, int B, int L, const qsp_c128* gates, int per_instance, const int32_t* kind,
                                     int q_hi, int q_lo, void* stream) {
  if (!state || !gates || B <= 0 || L <= 0 || L > 30 || q_hi < 0 || q_hi >= L || q_lo >= L || q_lo == q_hi) return QSP_ERR_ARG;
  if (q_lo < 0 && kind) return QSP_ERR_ARG;
  DevInfo d = dev_info();
  const long tot = (long)B << (L >= 2 ? L - 2 : 0);
  const int grid = (int)imax(1, (int)((tot + 255) / 256 < (long)d.sms * 16 ? (tot + 255) / 256 : (long)d.sms * 16));
  QSP_LAUNCH(k_sv_gate, grid, 256, 0, (qsp_stream_t)stream, (cplx*)state, B, L, (const cplx*)gates, per_instance ? 16L : 0L, kind, 1L,
             q_hi, q_lo);
  ++g_launches;
  return check_launch();
}
// |psi_b> = U_0 ... U_{n-1} |0..0> for the gate layers G[n][B][L][4][4], kind[n][B][L] of a batch: layers in reverse
// order, sites ascending, a two-qubit gate of site `it` on [it+1, it], a one-qubit gate on [it] (qiskit_circuit.py:38-48)
extern "C" int qsp_circuit_statevector(const qsp_c128* G, const int32_t* kind, int n, int B, int L, qsp_c128* state, void* stream) {
  if (!G || !kind || !state || n <= 0 || B <= 0 || L <= 0 || L > 30) return QSP_ERR_ARG;
  DevInfo d = dev_info();
  const long tot = (long)B << L;
  QSP_LAUNCH(k_sv_init, (int)imin((int)((tot + 255) / 256 < 65535 ? (tot + 255) / 256 : 65535), d.sms * 16), 256, 0, (qsp_stream_t)stream,
             (cplx*)state, B, L);
  ++g_launches;
  // one launch per (layer, site): instance b applies G[layer][b][site] with its own kind (segments differ per instance)
  for (int it_layer = n - 1; it_layer >= 0; --it_layer)
    for (int it = 0; it < L; ++it) {
      const qsp_c128* g = G + ((size_t)it_layer * B * L + it) * 16;
      const int32_t* kd = kind + (size_t)it_layer * B * L + it;
      const long tg = (long)B << (L >= 2 ? L - 2 : 0);
      const int grid = (int)imax(1, (int)((tg + 255) / 256 < (long)d.sms * 16 ? (tg + 255) / 256 : (long)d.sms * 16));
      QSP_LAUNCH(k_sv_gate, grid, 256, 0, (qsp_stream_t)stream, (cplx*)state, B, L, (const cplx*)g, (long)L * 16, kd, (long)L, it,
                 it + 1 < L ? it + 1 : -1);
      ++g_launches;
    }
  return check_launch();
}

// ------------------------------------------------------------------------------------
// input pipeline: MPSPreparation.__init__ + sequential.py:38-40
// ------------------------------------------------------------------------------------
static size_t mps_bytes(int B, int L, int D) { return align_up((size_t)B * L * D * 2 * D * sizeof(cplx)); }
static size_t bonds_bytes(int B, int L) { return align_up((size_t)B * (L + 1) * sizeof(int32_t)); }

// scratch shared by the left sweep (large bonds only) and the right sweep of a compress at bond D
static size_t compress_scratch_bytes(int B, int L, int D_in, int D_out) {
  DevInfo d = dev_info();
  LCanonPlan lp;
  if (!plan_lcanon(D_in, d, &lp)) return 0;
  size_t w = qsp_right_compress_workspace_bytes(B, L, D_in, D_out);
  if (w == 0) return 0;
  size_t l = qsp_left_canonize_workspace_bytes(B, L, D_in);
  return w > l ? w : l;
}
extern "C" size_t qsp_prepare_target_workspace_bytes(int B, int L, int D) {
  size_t w = compress_scratch_bytes(B, L, D, D);
  if (w == 0) return 0;
  return mps_bytes(B, L, D) + bonds_bytes(B, L) + w;
}
extern "C" int qsp_prepare_target(qsp_c128* mps, int32_t* bonds, int B, int L, int D, double cutoff, int32_t* status,
                                  void* work, size_t work_bytes, void* stream) {
  if (!mps || !bonds || !work) return QSP_ERR_ARG;
  size_t need = qsp_prepare_target_workspace_bytes(B, L, D);
  if (need == 0) return QSP_ERR_UNSUPPORTED;
  if (work_bytes < need) return QSP_ERR_WORKSPACE;
  char* w = (char*)work;
  qsp_c128* tmp = (qsp_c128*)w;
  w += mps_bytes(B, L, D);
  int32_t* tb = (int32_t*)w;
  w += bonds_bytes(B, L);
  size_t wb = compress_scratch_bytes(B, L, D, D);
  int rc = qsp_left_canonize(mps, bonds, B, L, D, status, w, wb, stream);
  if (rc) return rc;
  rc = qsp_right_compress(mps, bonds, D, tmp, tb, D, B, L, 0, cutoff, 1, nullptr, 0, nullptr, status, w, wb, stream);
  if (rc) return rc;
  dev_copy(mps, tmp, (size_t)B * L * D * 2 * D * sizeof(cplx), (qsp_stream_t)stream);
  dev_copy(bonds, tb, (size_t)B * (L + 1) * sizeof(int32_t), (qsp_stream_t)stream);
  return check_launch();
}

// ------------------------------------------------------------------------------------
// the layer loop
// ------------------------------------------------------------------------------------
// Encoded state U_0..U_k|0..0> (core.py:217-221).  Every forward layer doubles the bonds exactly
// (k_expand_forward) and keeps the state right canonical and normalised, so the reference's
// compress('right') after each layer (core.py:140, cutoff only) changes nothing but the bond
// sizes.  Mode A (2^layers <= QSP_ENC_CAP, i.e. up to 11 layers): buffers of 2^layers, no compression at
// all.  Mode B (more layers, any chain): like the reference the state is compressed (cutoff only) whenever
// the next doubling would not fit the De-padded buffers; the compressed bonds must fit De / 2 = 1024
// (QSP_ST_BOND_OVERFLOW otherwise; the reference's Heisenberg L=32 benchmark reaches 517 at 12 layers).  The exact rank of an L-site state is at most 2^(L/2), so short chains get by with
// De = 2^(L/2+1).
#define QSP_ENC_CAP 2048  // (134 MB per site and instance at this padding: large batches are chunked by the caller)
struct EncPlan {
  int De;        // padded bond of both encoded-state buffers
  int rank_cap;  // bound on the bonds after a compression
  bool compress; // mode B
  bool ok;
};
static EncPlan plan_enc(int L, int layers) {
  EncPlan e;
  e.ok = true;
  e.compress = false;
  const int half = imin(L / 2, 20);
  const int cap = g_enc_cap ? g_enc_cap : QSP_ENC_CAP;
  if (layers <= 30 && (1L << layers) <= cap) {
    e.De = 1 << layers;
    e.rank_cap = e.De;
  } else {
    e.compress = true;
    e.De = (half + 1 <= 6 && cap >= 64) ? 64 : cap;  // 2^(L/2+1) <= 64: the small kernels take it
    e.rank_cap = imin(1 << half, e.De / 2);
  }
  return e;
}
extern "C" size_t qsp_encoded_overlap_workspace_bytes(int B, int L, int D, int layers) {
  EncPlan e = plan_enc(L, layers);
  if (!e.ok) return 0;
  size_t s = qsp_overlap_workspace_bytes(B, L, e.De, D);
  if (e.compress) {
    size_t t = compress_scratch_bytes(B, L, e.De, e.De);
    if (!t) return 0;
    s = t > s ? t : s;
  }
  return 2 * (mps_bytes(B, L, e.De) + bonds_bytes(B, L)) + s;
}
__global__ void k_zero_state(cplx* mps, int32_t* bonds, int B, int L, int D);
// <target| U_0 ... U_{n-1} |0..0> for G[n][B][L][4][4], kind[n][B][L]  (sequential.py:67-70, tsp.py:130-132)
extern "C" int qsp_encoded_overlap(const qsp_c128* G, const int32_t* kind, int n, const qsp_c128* target,
                                   const int32_t* target_bonds, int B, int L, int D, double cutoff, qsp_c128* result,
                                   int32_t* status, void* work, size_t work_bytes, void* stream) {
  if (!G || !kind || !target || !target_bonds || !result || !work || n <= 0 || B <= 0 || L <= 0) return QSP_ERR_ARG;
  EncPlan e = plan_enc(L, n);
  if (!e.ok) return QSP_ERR_UNSUPPORTED;
  size_t need = qsp_encoded_overlap_workspace_bytes(B, L, D, n);
  if (!need) return QSP_ERR_UNSUPPORTED;
  if (work_bytes < need) return QSP_ERR_WORKSPACE;
  const int De = e.De;
  char* w = (char*)work;
  qsp_c128* ea = (qsp_c128*)w;
  w += mps_bytes(B, L, De);
  int32_t* eab = (int32_t*)w;
  w += bonds_bytes(B, L);
  qsp_c128* eb = (qsp_c128*)w;
  w += mps_bytes(B, L, De);
  int32_t* ebb = (int32_t*)w;
  w += bonds_bytes(B, L);
  const size_t sb = work_bytes - (size_t)(w - (char*)work);
  qsp_stream_t st = (qsp_stream_t)stream;
  long nl = (long)B * L;
  QSP_LAUNCH(k_zero_state, (unsigned)((nl + 127) / 128), 128, 0, st, (cplx*)ea, eab, B, L, De);
  ++g_launches;
  int rc, cur = 1;
  const int rank_cap = e.rank_cap;
  for (int k = n - 1; k >= 0; --k) {
    if (2 * cur > De) {  // mode B: bring the bonds back to the true rank before doubling them again
      if ((rc = left_canonize_live(ea, eab, B, L, De, cur, status, w, sb, stream))) return rc;
      if ((rc = right_compress_live(ea, eab, De, eb, ebb, De, cur, imin(cur, rank_cap), B, L, 0, cutoff, 0, nullptr, 0,
                                    nullptr, status, w, sb, stream)))
        return rc;
      qsp_c128* t = ea; ea = eb; eb = t;
      int32_t* tb = eab; eab = ebb; ebb = tb;
      cur = imin(cur, rank_cap);  // a larger true rank was truncated and flagged QSP_ST_BOND_OVERFLOW by the sweep
    }
    const qsp_c128* Gk = G + (size_t)k * B * L * 16;
    const int32_t* kk = kind + (size_t)k * B * L;
    if ((rc = qsp_apply_forward_layer(ea, eab, De, Gk, kk, eb, ebb, De, B, L, status, stream))) return rc;
    qsp_c128* t = ea; ea = eb; eb = t;
    int32_t* tb = eab; eab = ebb; ebb = tb;
    cur = 2 * cur;
  }
  return qsp_overlap(ea, eab, De, target, target_bonds, D, B, L, result, w, sb, stream);
}

struct SeqPlan {
  int De;
  size_t off_psi, off_psib, off_psi2, off_psi2b, off_d2, off_d2b;
  size_t off_enc, off_scratch;
  size_t enc_bytes, scratch_bytes, total;
  bool ok;
};
static SeqPlan plan_seq(int B, int L, int D, int layers, int max_bond, int flags) {
  (void)max_bond;
  SeqPlan p;
  memset(&p, 0, sizeof(p));
  p.ok = true;
  const bool enc = !(flags & QSP_FLAG_NO_ENCODED);
  const bool second = (flags & QSP_FLAG_SECOND_PASS) != 0;
  p.De = 1;
  size_t o = 0;
  p.off_psi = o, o += mps_bytes(B, L, D);
  p.off_psib = o, o += bonds_bytes(B, L);
  p.off_psi2 = o, o += mps_bytes(B, L, D);  // ping-pong partner of psi (the doubled U^dagger psi is never stored)
  p.off_psi2b = o, o += bonds_bytes(B, L);
  p.off_d2 = o, o += mps_bytes(B, L, 2);
  p.off_d2b = o, o += bonds_bytes(B, L);
  size_t s = 0, t;
  DevInfo d = dev_info();
  LCanonPlan lp;
  if (!plan_lcanon(D, d, &lp)) p.ok = false;
  s = qsp_left_canonize_workspace_bytes(B, L, D);
  t = qsp_right_compress_workspace_bytes(B, L, D, 2);
  if (!t) p.ok = false;
  s = t > s ? t : s;
  t = qsp_right_compress_workspace_bytes(B, L, 2 * D, D);
  if (!t) p.ok = false;
  s = t > s ? t : s;
  if (second) {
    t = qsp_right_compress_workspace_bytes(B, L, D, D);
    if (!t) p.ok = false;
    s = t > s ? t : s;
  }
  if (enc) {
    p.enc_bytes = qsp_encoded_overlap_workspace_bytes(B, L, D, layers);
    if (!p.enc_bytes) p.ok = false;
    p.off_enc = o, o += align_up(p.enc_bytes);
  }
  p.off_scratch = o;
  p.scratch_bytes = s;
  p.total = o + s;
  return p;
}

extern "C" size_t qsp_seq_prepare_workspace_bytes(int B, int L, int D, int layers, int max_bond, int flags) {
  if (B <= 0 || L <= 0 || D <= 0 || layers <= 0) return 0;
  SeqPlan p = plan_seq(B, L, D, layers, max_bond, flags);
  return p.ok ? p.total : 0;
}

// zero state |0...0> into enc (pad De): helper kernel
__global__ void k_zero_state(cplx* mps, int32_t* bonds, int B, int L, int D) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)B * L) return;
  int inst = (int)(idx / L), i = (int)(idx - (long)inst * L);
  cplx* a = mps + idx * ((long)D * 2 * D);
  a[0] = cmake(1.0, 0.0);      // (l=0,p=0,r=0)
  a[(long)D] = cmake(0.0, 0.0);  // (l=0,p=1,r=0)
  bonds[(long)inst * (L + 1) + i] = 1;
  if (i == L - 1) bonds[(long)inst * (L + 1) + L] = 1;
}

extern "C" int qsp_seq_prepare(const qsp_c128* target, const int32_t* target_bonds, int B, int L, int D, int layers,
                               int max_bond, double cutoff, int overlaps_gap, int flags, qsp_c128* overlaps,
                               qsp_c128* dis_overlaps, qsp_c128* G, int32_t* kind, double* d2_svals,
                               qsp_c128* final_mps, int32_t* final_bonds, int32_t* status, void* work,
                               size_t work_bytes, void* stream) {
  if (!target || !target_bonds || !dis_overlaps || !G || !kind || !status || !work) return QSP_ERR_ARG;
  if (B <= 0 || L <= 0 || D <= 0 || layers <= 0 || overlaps_gap <= 0) return QSP_ERR_ARG;
  if (max_bond > D) return QSP_ERR_ARG;
  SeqPlan p = plan_seq(B, L, D, layers, max_bond, flags);
  if (!p.ok) return QSP_ERR_UNSUPPORTED;
  if (work_bytes < p.total) return QSP_ERR_WORKSPACE;
  const bool enc_on = !(flags & QSP_FLAG_NO_ENCODED) && overlaps;
  const bool second = (flags & QSP_FLAG_SECOND_PASS) != 0 && max_bond > 0;
  qsp_stream_t st = (qsp_stream_t)stream;
  char* w = (char*)work;
  qsp_c128* psi = (qsp_c128*)(w + p.off_psi);
  int32_t* psib = (int32_t*)(w + p.off_psib);
  qsp_c128* psi2 = (qsp_c128*)(w + p.off_psi2);
  int32_t* psi2b = (int32_t*)(w + p.off_psi2b);
  qsp_c128* d2 = (qsp_c128*)(w + p.off_d2);
  int32_t* d2b = (int32_t*)(w + p.off_d2b);
  void* encw = w + p.off_enc;
  void* scratch = w + p.off_scratch;
  const size_t sb = p.scratch_bytes;
  int rc;

  dev_copy(psi, target, (size_t)B * L * D * 2 * D * sizeof(cplx), st);
  dev_copy(psib, target_bonds, (size_t)B * (L + 1) * sizeof(int32_t), st);
  if (overlaps) dev_fill_nan(overlaps, (long)layers * B, st);

  for (int it = 0; it < layers; ++it) {
    char nm[48];
    snprintf(nm, sizeof(nm), "qsp layer %d", it);
    NvtxRange layer_range(nm);
    qsp_c128* Git = G + (size_t)it * B * L * 16;
    int32_t* kit = kind + (size_t)it * B * L;
    double* sv = d2_svals ? d2_svals + (size_t)it * B * (L - 1) * 2 : nullptr;
    // generate_bond_d_unitary (core.py:12-22): compress('right', max_bond=2) = left sweep + truncating right sweep
    {
      NvtxRange r("left_canonize");
      if ((rc = qsp_left_canonize(psi, psib, B, L, D, status, scratch, sb, stream))) return rc;
    }
    {
      NvtxRange r("d2_truncation + unitaries");
      if ((rc = qsp_right_compress(psi, psib, D, d2, d2b, 2, B, L, 2, cutoff, 1, sv, 2, nullptr, status, scratch, sb,
                                   stream)))
        return rc;
      if ((rc = qsp_generate_unitaries(d2, d2b, B, L, Git, kit, status, stream))) return rc;
    }
    NvtxRange fused_range("inverse layer + re-truncation sweep");
    // sequential.py:51-57 in ONE sweep: apply_inverse_unitary_layer_on_wfn on the left-canonical psi (exact,
    // left canonical, never materialised) fused with [right_canonize; compress('right', max_bond)];
    // right_canonize(normalize); compress()  -- the fused gate-apply-and-split kernel
    qsp_c128* amp = dis_overlaps + (size_t)it * B;
    if ((rc = right_compress_live(psi, psib, D, psi2, psi2b, D, 2 * D, D, B, L, max_bond, cutoff, 1, nullptr, 0,
                                  second ? nullptr : amp, status, scratch, sb, stream, Git, kit)))
      return rc;
    {
      qsp_c128* t = psi; psi = psi2; psi2 = t;
      int32_t* tb = psib; psib = psi2b; psi2b = tb;
    }
    if (second) {
      if ((rc = qsp_left_canonize(psi, psib, B, L, D, status, scratch, sb, stream))) return rc;
      if ((rc = qsp_right_compress(psi, psib, D, psi2, psi2b, D, B, L, 0, cutoff, 1, nullptr, 0, amp, status, scratch,
                                   sb, stream)))
        return rc;
      qsp_c128* t = psi; psi = psi2; psi2 = t;
      int32_t* tb = psib; psib = psi2b; psi2b = tb;
    }
    // sequential.py:67-70: encoded state U_0 .. U_it |0..0> and its overlap with the target
    if (enc_on && (it % overlaps_gap == 0 || it + 1 == layers)) {
      if ((rc = qsp_encoded_overlap(G, kind, it + 1, target, target_bonds, B, L, D, cutoff, overlaps + (size_t)it * B,
                                    status, encw, p.enc_bytes, stream)))
        return rc;
    }
  }
  if (final_mps) dev_copy(final_mps, psi, (size_t)B * L * D * 2 * D * sizeof(cplx), st);
  if (final_bonds) dev_copy(final_bonds, psib, (size_t)B * (L + 1) * sizeof(int32_t), st);
  return check_launch();
}
