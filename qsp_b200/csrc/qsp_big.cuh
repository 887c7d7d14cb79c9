// Large-bond path (bonds beyond what one SM holds on chip; BASELINE.json configs[4]: chi = 256).
//
// The small-bond kernels keep one MPS instance on one SM for a whole sweep.  Here a site matrix (up to
// 2 chi x 2 chi complex128 = 4 MB at chi = 256) lives in HBM / L2 and EVERY step of a sweep is a batched
// launch over (instance x tile), so all 148 SMs work whatever the batch size:
//
//   k_big_gemm          batched strided complex GEMM on the FP64 tensor cores (DMMA m8n8k4), per-instance
//                       operand descriptors in device memory
//   k_big_jacobi_round  one round of a BLOCK one-sided Jacobi SVD: a CTA owns a pair of 16-row blocks of the
//                       matrix, streams the 32-row panel through shared memory with TMA bulk copies, forms the
//                       32 x 32 Gram matrix on the tensor cores, factors it (Cholesky), runs one sweep of scalar
//                       Jacobi rotations on the 32 x 32 factor (accumulating the 32 x 32 unitary J), and applies
//                       J to the panel as a second tensor-core GEMM.  Pairs follow the round-robin tournament;
//                       block pairs of one round touch disjoint rows.
//   k_big_*             elementwise / per-instance helpers of the two sweeps
//
// The inner factor keeps the accuracy of Hestenes' method: the Gram entries of two rows are exactly what the
// scalar method reduces, the Cholesky factor L of G = X X^H has the same row geometry as X (L L^H = G), and
// the rotations are computed from L (Veselic-Hari), so small rows are resolved relative to their own norms.
#pragma once
#include "qsp_common.cuh"
#include "qsp_jacobi.cuh"
#include "qsp_mma.cuh"

#define QSP_BIG_RB 16   // rows per Jacobi block
#define QSP_BIG_PR 32   // rows per block pair (panel)
#define QSP_BIG_CK 32   // columns per streamed chunk (rows of W are zero-padded to a multiple of it)
#define QSP_BIG_LDS 34  // shared row stride of a chunk (complex): 4 k-rows x 32 B hit distinct bank groups
#define QSP_BIG_LDY 69  // shared row stride of Y = [G or L | J] in the solve kernel (odd: rows and columns conflict free)
#define QSP_BIG_LDJ 36  // shared row stride of J in the update kernel
#define QSP_BIG_NS 4    // TMA ring stages (two chunks consumed between block barriers, two in flight)
#define QSP_BIG_NT 256  // threads per CTA of the round and GEMM kernels

// ----------------------------------------------------------------------------------
// batched GEMM
// ----------------------------------------------------------------------------------
struct BigGemm {
  const cplx* A;
  const cplx* B;
  cplx* C;
  int m, n, k;
  long sAr, sAk, sBk, sBc, sCr, sCc;
  int conjA, conjB;
  int mode;  // 0: C = A B, 1: C += A B, 2: C -= A B
  int pad;
};

#define QSP_GB_TM 64
#define QSP_GB_TN 64
#define QSP_GB_TK 8
#define QSP_GB_LD 12  // shared stride of a k-row group: 2 rows x 64 B of a quarter-warp hit distinct banks

__host__ __device__ inline size_t big_gemm_smem_bytes() {
  return (size_t)2 * (QSP_GB_TM + QSP_GB_TN) * QSP_GB_LD * sizeof(cplx) + 64;
}

// grid.x = ndesc * tiles_m * tiles_n (tile index fastest); CTAs whose tile lies outside the live extent exit
__global__ void __launch_bounds__(QSP_BIG_NT, 2) k_big_gemm(const BigGemm* __restrict__ gd, int tiles_m, int tiles_n) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = tiles_m * tiles_n;
  const int di = blockIdx.x / per, t = blockIdx.x - di * per;
  const int tm = t / tiles_n, tn = t - tm * tiles_n;
  const BigGemm g = gd[di];
  const int i0 = tm * QSP_GB_TM, j0 = tn * QSP_GB_TN;
  if (g.m <= 0 || g.n <= 0 || i0 >= g.m || j0 >= g.n) return;
  cplx* sm = (cplx*)QSP_SMEM_BASE();
  // two buffers of As[64][LD] (element (row, k)) and Bs[64][LD] (element (col, k))
  constexpr int kBuf = (QSP_GB_TM + QSP_GB_TN) * QSP_GB_LD;
  cplx* const As0 = sm;
  cplx* const Bs0 = sm + QSP_GB_TM * QSP_GB_LD;
  // loader mapping: 512 elements of A and of B per k-chunk, two of each per thread
  const bool kfastA = (g.sAk == 1), kfastB = (g.sBk == 1);
  int ar[2], ak[2], bc[2], bk[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int idx = tid + u * QSP_BIG_NT;
    if (kfastA) { ak[u] = idx % QSP_GB_TK; ar[u] = idx / QSP_GB_TK; } else { ar[u] = idx % QSP_GB_TM; ak[u] = idx / QSP_GB_TM; }
    if (kfastB) { bk[u] = idx % QSP_GB_TK; bc[u] = idx / QSP_GB_TK; } else { bc[u] = idx % QSP_GB_TN; bk[u] = idx / QSP_GB_TN; }
  }
  cplx ra[2], rb[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      cplx v = cmake(0.0, 0.0);
      if (i0 + ar[u] < g.m && k0 + ak[u] < g.k) {
        v = g.A[(long)(i0 + ar[u]) * g.sAr + (long)(k0 + ak[u]) * g.sAk];
        if (g.conjA) v.y = -v.y;
      }
      ra[u] = v;
      cplx w = cmake(0.0, 0.0);
      if (j0 + bc[u] < g.n && k0 + bk[u] < g.k) {
        w = g.B[(long)(k0 + bk[u]) * g.sBk + (long)(j0 + bc[u]) * g.sBc];
        if (g.conjB) w.y = -w.y;
      }
      rb[u] = w;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      As0[buf * kBuf + ar[u] * QSP_GB_LD + ak[u]] = ra[u];
      Bs0[buf * kBuf + bc[u] * QSP_GB_LD + bk[u]] = rb[u];
    }
  };
  // warp tile: 16 rows x 32 columns = 2 x 4 complex 8x8 tiles; warps 4 (m) x 2 (n)
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 32;
  ctile acc[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = ctile_zero();
  const int fr = lane >> 2, fk = lane & 3;
  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < g.k; k0 += QSP_GB_TK) {
    const bool more = k0 + QSP_GB_TK < g.k;
    if (more) gload(k0 + QSP_GB_TK);
#pragma unroll
    for (int ks = 0; ks < QSP_GB_TK; ks += 4) {
      cplx af[2], bf[4];
#pragma unroll
      for (int a = 0; a < 2; ++a) af[a] = As0[buf * kBuf + (wm + a * 8 + fr) * QSP_GB_LD + ks + fk];
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = Bs0[buf * kBuf + (wn + b * 8 + fr) * QSP_GB_LD + ks + fk];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) cdmma(acc[a][b], af[a].x, af[a].y, bf[b].x, bf[b].y);
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
  // epilogue: lane holds rows fr, columns 2 fk + {0, 1} of every tile
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int gi = i0 + wm + a * 8 + fr;
    if (gi >= g.m) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gj = j0 + wn + b * 8 + 2 * fk + e;
        if (gj >= g.n) continue;
        cplx v = cmake(e ? acc[a][b].r1 : acc[a][b].r0, e ? acc[a][b].i1 : acc[a][b].i0);
        cplx* dst = g.C + (long)gi * g.sCr + (long)gj * g.sCc;
        if (g.mode == 1) v = cadd(*dst, v);
        else if (g.mode == 2) v = csub(*dst, v);
        *dst = v;
      }
    }
  }
}

// ----------------------------------------------------------------------------------
// block Jacobi
// ----------------------------------------------------------------------------------
struct BigJac {
  cplx* W;     // nrows x ldw (row-major); columns [len, ldw) and the Gram padding are zero
  int nrows;   // rows to orthogonalise
  int ndotp;   // leading columns that enter the Gram matrix (multiple of QSP_BIG_CK)
  int len;     // columns to rotate (multiple of QSP_BIG_CK, >= ndotp)
  int ldw;
  int nblk;    // number of 16-row blocks, padded to an even number (0: instance inactive)
  int inner;   // Jacobi sweeps on the 32 x 32 factor per visit
  double tol2; // rotation threshold on |g|^2 / (alpha beta)
};

// row r (0..31) of the panel -> row of W, or -1
__device__ __forceinline__ int big_panel_row(int r, int ra0, int na, int rb0, int nb) {
  if (r < QSP_BIG_RB) return r < na ? ra0 + r : -1;
  return (r - QSP_BIG_RB) < nb ? rb0 + (r - QSP_BIG_RB) : -1;
}

// A round of the block Jacobi = three launches (a CTA owns one block pair = "task" in each of them):
//   k_big_jac_gram    G = X X^H of the 32-row panel on the tensor cores, panel streamed with TMA; the Gram columns
//                     can be split over `splits` CTAs (partial Gram matrices, summed in fixed order by the solver)
//   k_big_jac_solve   32 x 32 problem: convergence test, Cholesky, scalar Jacobi sweep -> unitary J (or "skip")
//   k_big_jac_update  X <- J X on the tensor cores, columns split over `splits` CTAs
// Splitting keeps the tensor-core launches many waves deep (no tail) and lets the latency-bound 32 x 32 solves of all
// pairs run side by side (six CTAs per SM) instead of serialising inside two fat CTAs per SM.
struct BigTask {
  int ra0, na, rb0, nbr;  // rows of the two blocks
  bool live;
};
__device__ __forceinline__ BigTask big_task(const BigJac& J, int pi, int round) {
  BigTask t;
  t.live = false;
  t.ra0 = t.na = t.rb0 = t.nbr = 0;
  const int nb = J.nblk;
  if (nb < 2 || pi >= nb / 2 || round >= nb - 1) return t;
  int ba, bb;
  if (pi == 0) {
    ba = nb - 1;
    bb = round;
  } else {
    ba = (round + pi) % (nb - 1);
    bb = (round + nb - 1 - pi) % (nb - 1);
  }
  if (ba > bb) {
    const int x = ba;
    ba = bb;
    bb = x;
  }
  t.ra0 = ba * QSP_BIG_RB, t.rb0 = bb * QSP_BIG_RB;
  t.na = imax(0, imin(QSP_BIG_RB, J.nrows - t.ra0)), t.nbr = imax(0, imin(QSP_BIG_RB, J.nrows - t.rb0));
  t.live = t.na + t.nbr >= 2;
  return t;
}
// chunks [c0, c1) of `nch` handled by split `sp` of `splits`
__device__ __forceinline__ void big_split_range(int nch, int sp, int splits, int* c0, int* c1) {
  const int per = (nch + splits - 1) / splits;
  *c0 = imin(nch, sp * per);
  *c1 = imin(nch, *c0 + per);
}

__host__ __device__ inline size_t big_ring_bytes() { return (size_t)QSP_BIG_NS * QSP_BIG_PR * QSP_BIG_LDS * sizeof(cplx); }
__host__ __device__ inline size_t big_gram_smem_bytes() { return big_ring_bytes() + 2 * QSP_BIG_NS * 32 + 64; }
__host__ __device__ inline size_t big_solve_smem_bytes() { return (size_t)QSP_BIG_PR * QSP_BIG_LDY * sizeof(cplx) + 64 * sizeof(double) + 64; }
__host__ __device__ inline size_t big_update_smem_bytes() {
  return big_ring_bytes() + (size_t)QSP_BIG_PR * QSP_BIG_LDJ * sizeof(cplx) + 2 * QSP_BIG_NS * 32 + 64;
}

// TMA ring of 32-row x 32-column chunks of one panel: warp 0 issues one bulk copy per live row, chunks are consumed in
// issue order (stage = n % NS, parity = (n / NS) & 1)
struct BigRing {
  cplx* ring;
  qsp_mbar* full;   // [NS] bytes landed
  qsp_mbar* empty;  // [NS] all warps have finished reading the stage
  const cplx* W;
  long ldw;
  int myrow, nlive, issued, consumed;
  __device__ __forceinline__ void init(char* smem_ring, qsp_mbar* b, const BigJac& J, const BigTask& t) {
    ring = (cplx*)smem_ring;
    full = b;
    empty = b + QSP_BIG_NS;
    W = J.W;
    ldw = J.ldw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    nlive = t.na + t.nbr;
    myrow = (warp == 0) ? big_panel_row(lane, t.ra0, t.na, t.rb0, t.nbr) : -1;
    issued = consumed = 0;
    // dead rows are zero for the whole kernel (live rows are overwritten by the bulk copies before they are read)
    if (nlive < QSP_BIG_PR)
      for (int idx = tid; idx < QSP_BIG_NS * QSP_BIG_PR * QSP_BIG_LDS; idx += blockDim.x) {
        const int r = (idx / QSP_BIG_LDS) % QSP_BIG_PR;
        if (big_panel_row(r, t.ra0, t.na, t.rb0, t.nbr) < 0) ring[idx] = cmake(0.0, 0.0);
      }
    fence_proxy_async();
    if (tid == 0) {
      for (int s = 0; s < QSP_BIG_NS; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], (int)(blockDim.x >> 5));
      }
      fence_mbar_init();
    }
    __syncthreads();
  }
  // warp 0 (all lanes call): waits until every warp has released the stage's previous chunk, then refills it
  __device__ __forceinline__ void issue(int chunk) {
    const int n = issued++, stage = n % QSP_BIG_NS;
    if ((threadIdx.x >> 5) == 0) {
      const int lane = threadIdx.x & 31;
      if (n >= QSP_BIG_NS) mbar_wait(&empty[stage], ((n / QSP_BIG_NS) - 1) & 1);
      if (lane == 0) mbar_expect_tx(&full[stage], (unsigned)(nlive * QSP_BIG_CK * sizeof(cplx)));
      __syncwarp();
      if (myrow >= 0)
        tma_load_1d(ring + ((long)stage * QSP_BIG_PR + lane) * QSP_BIG_LDS, W + (long)myrow * ldw + (long)chunk * QSP_BIG_CK,
                    (unsigned)(QSP_BIG_CK * sizeof(cplx)), &full[stage]);
    }
  }
  __device__ __forceinline__ const cplx* wait_next() {
    const int stage = consumed % QSP_BIG_NS;
    mbar_wait(&full[stage], (consumed / QSP_BIG_NS) & 1);
    return ring + (long)stage * QSP_BIG_PR * QSP_BIG_LDS;
  }
  // the calling warp has finished reading the chunk returned by the last wait_next()
  __device__ __forceinline__ void release() {
    const int stage = consumed % QSP_BIG_NS;
    ++consumed;
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[stage]);
  }
};

// Gp[task][split][32][32]: partial Gram matrices (full Hermitian storage, row-major)
__global__ void __launch_bounds__(QSP_BIG_NT) k_big_jac_gram(const BigJac* __restrict__ jd, int round, int pairs_max, int splits,
                                                             cplx* __restrict__ Gp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int task = blockIdx.x / splits, sp = blockIdx.x - task * splits;
  const int inst = task / pairs_max, pi = task - inst * pairs_max;
  const BigJac J = jd[inst];
  const BigTask t = big_task(J, pi, round);
  if (!t.live) return;
  int c0, c1;
  big_split_range(J.ndotp / QSP_BIG_CK, sp, splits, &c0, &c1);
  char* smc = QSP_SMEM_BASE();
  BigRing R;
  R.init(smc, (qsp_mbar*)(smc + big_ring_bytes()), J, t);
  ctile acc[10];
#pragma unroll
  for (int u = 0; u < 10; ++u) acc[u] = ctile_zero();
  const int fr = lane >> 2, fk = lane & 3;
  int next = c0;
  for (; next < c1 && next < c0 + 2; ++next) R.issue(next);
  for (int c = c0; c < c1; c += 2) {
    for (int u = 0; u < 2 && next < c1; ++u) R.issue(next++);
    for (int u = 0; u < 2 && c + u < c1; ++u) {
      const cplx* X = R.wait_next();
      // warp w takes k-step w of the chunk (columns 4w .. 4w+3): four row-tile fragments feed all ten tiles; the four
      // real DMMAs of a complex tile are issued tile-interleaved so that no DMMA waits for its predecessor
      cplx f[4];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti) f[ti] = X[(ti * 8 + fr) * QSP_BIG_LDS + warp * 4 + fk];
      int q = 0;
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = ti; tj < 4; ++tj, ++q) dmma_8x8x4(acc[q].r0, acc[q].r1, f[ti].x, f[tj].x);
      q = 0;
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = ti; tj < 4; ++tj, ++q) dmma_8x8x4(acc[q].i0, acc[q].i1, f[ti].y, f[tj].x);
      q = 0;
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = ti; tj < 4; ++tj, ++q) dmma_8x8x4(acc[q].r0, acc[q].r1, f[ti].y, f[tj].y);
      q = 0;
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = ti; tj < 4; ++tj, ++q) dmma_8x8x4(acc[q].i0, acc[q].i1, -f[ti].x, f[tj].y);
      R.release();  // (the fragments are in registers: the stage may be refilled)
    }
  }
  __syncthreads();
  // ---- reduce the eight partial Gram matrices (deterministic tree through the idle ring) ---
  double* part = (double*)R.ring;  // [warp][tile][4][32 lanes]
  for (int half = 4; half >= 1; half >>= 1) {
    if (warp >= half && warp < 2 * half) {
      double* p = part + (long)(warp - half) * 10 * 4 * 32;
#pragma unroll
      for (int u = 0; u < 10; ++u) {
        p[(u * 4 + 0) * 32 + lane] = acc[u].r0;
        p[(u * 4 + 1) * 32 + lane] = acc[u].r1;
        p[(u * 4 + 2) * 32 + lane] = acc[u].i0;
        p[(u * 4 + 3) * 32 + lane] = acc[u].i1;
      }
    }
    __syncthreads();
    if (warp < half) {
      const double* p = part + (long)warp * 10 * 4 * 32;
#pragma unroll
      for (int u = 0; u < 10; ++u) {
        acc[u].r0 += p[(u * 4 + 0) * 32 + lane];
        acc[u].r1 += p[(u * 4 + 1) * 32 + lane];
        acc[u].i0 += p[(u * 4 + 2) * 32 + lane];
        acc[u].i1 += p[(u * 4 + 3) * 32 + lane];
      }
    }
    __syncthreads();
  }
  if (warp == 0) {
    cplx* G = Gp + ((long)task * splits + sp) * QSP_BIG_PR * QSP_BIG_PR;
    int q = 0;
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = ti; tj < 4; ++tj, ++q) {
        const int r = ti * 8 + fr, c = tj * 8 + 2 * fk;
        const cplx v0 = cmake(acc[q].r0, acc[q].i0), v1 = cmake(acc[q].r1, acc[q].i1);
        G[r * QSP_BIG_PR + c] = v0;
        G[r * QSP_BIG_PR + c + 1] = v1;
        if (tj > ti) {
          G[c * QSP_BIG_PR + r] = cconj(v0);
          G[(c + 1) * QSP_BIG_PR + r] = cconj(v1);
        }
      }
  }
}

// flags[inst] |= 1 when a rotation above QSP_JAC_QUAD happened (the sweep has to be repeated).
// Jg[task][32][32] receives the unitary of the pair, skip[task] = 1 when the pair needed no rotation.
__global__ void __launch_bounds__(QSP_BIG_NT, 4) k_big_jac_solve(const BigJac* __restrict__ jd, int* __restrict__ flags, int round, int pairs_max,
                                                              int splits, const cplx* __restrict__ Gp, cplx* __restrict__ Jg,
                                                              int* __restrict__ skip) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int task = blockIdx.x;
  const int inst = task / pairs_max, pi = task - inst * pairs_max;
  const BigJac J = jd[inst];
  const BigTask t = big_task(J, pi, round);
  if (!t.live) {
    if (tid == 0) skip[task] = 1;
    return;
  }
  char* smc = QSP_SMEM_BASE();
  cplx* Y = (cplx*)smc;
  double* sd = (double*)(Y + QSP_BIG_PR * QSP_BIG_LDY);  // [0..31] diagonal of G, [32] max ratio, [33] rotation flag
  {  // G = sum of the column-split partials, in fixed order
    int c0, c1;
    for (int idx = tid; idx < QSP_BIG_PR * QSP_BIG_PR; idx += QSP_BIG_NT) {
      cplx v = cmake(0.0, 0.0);
      for (int sp = 0; sp < splits; ++sp) {
        big_split_range(J.ndotp / QSP_BIG_CK, sp, splits, &c0, &c1);
        if (c1 > c0) v = cadd(v, Gp[((long)task * splits + sp) * QSP_BIG_PR * QSP_BIG_PR + idx]);
      }
      Y[(idx / QSP_BIG_PR) * QSP_BIG_LDY + (idx % QSP_BIG_PR)] = v;
    }
  }
  __syncthreads();
  // ---- convergence of this block pair: max |g_pq|^2 / (g_pp g_qq), and the spread of the row norms --------------------
  if (tid < QSP_BIG_PR) sd[tid] = Y[tid * QSP_BIG_LDY + tid].x;
  __syncthreads();
  {
    double mx = 0.0;
    for (int idx = tid; idx < QSP_BIG_PR * QSP_BIG_PR; idx += QSP_BIG_NT) {
      const int r = idx / QSP_BIG_PR, c = idx - r * QSP_BIG_PR;
      const double ab = sd[r] * sd[c];
      if (r != c && ab > 1e-280) mx = fmax(mx, cabs2(Y[r * QSP_BIG_LDY + c]) / ab);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, shfl_xor_d(mx, o));
    if (lane == 0) sd[40 + warp] = mx;
    if (warp == 0) {
      double dmax = sd[lane], dmin = sd[lane] > 0.0 ? sd[lane] : 0.0;
      for (int o = 16; o > 0; o >>= 1) {
        dmax = fmax(dmax, shfl_xor_d(dmax, o));
        dmin = fmin(dmin, shfl_xor_d(dmin, o));
      }
      if (lane == 0) sd[34] = (dmin >= 1e-8 * dmax) ? 1.0 : 0.0;  // well scaled: two-sided route below
    }
  }
  __syncthreads();
  if (tid == 0) {
    double mx = sd[40];
    for (int w = 1; w < QSP_BIG_NT / 32; ++w) mx = fmax(mx, sd[40 + w]);
    sd[32] = mx;
  }
  __syncthreads();
  const double maxratio = sd[32];
  if (!(maxratio > J.tol2)) {  // nothing to rotate in this pair
    if (tid == 0) skip[task] = 1;
    return;
  }
  if (tid == 0) {
    skip[task] = 0;
    if (maxratio > QSP_JAC_QUAD * QSP_JAC_QUAD) flags[inst] = 1;
  }

  // ---- well-scaled panel (row norms within 1e4 of each other): rotate G itself two-sidedly, G <- R G R^H.  g_pp, g_qq, g_pq
  // ARE the Gram entries Hestenes' method reduces, so this is the same rotation sequence as the sweep on the rows of the
  // Cholesky factor below, without the factorisation, the dot products and the shuffle reductions.  J is a product of exact
  // plane rotations, so X <- J X stays an orthogonal transformation whatever rounding G accumulates, and convergence is
  // always judged on a freshly computed Gram matrix; only the RELATIVE accuracy of entries far below the largest one is
  // lost by updating G explicitly, which is why strongly graded (or rank deficient) panels take the Cholesky route. -------
  if (sd[34] != 0.0) {
    for (int idx = tid; idx < QSP_BIG_PR * QSP_BIG_PR; idx += QSP_BIG_NT) {
      const int i = idx / QSP_BIG_PR, k = idx - i * QSP_BIG_PR;
      Y[i * QSP_BIG_LDY + QSP_BIG_PR + k] = cmake(i == k ? 1.0 : 0.0, 0.0);
    }
    const int team = tid >> 4, tl = tid & 15;  // 16 pairs x 16 lanes
    volatile double* rotflag = sd + 33;
    const bool bip = round > 0;
    const int nr = bip ? QSP_BIG_RB : QSP_BIG_PR - 1;
    for (int sw = 0; sw < J.inner; ++sw) {
      if (tid == 0) *rotflag = 0.0;
      __syncthreads();
      for (int r = 0; r < nr; ++r) {
        int a, b;
        if (bip) {
          a = team;
          b = QSP_BIG_RB + ((team + r) & (QSP_BIG_RB - 1));
        } else if (team == 0) {
          a = QSP_BIG_PR - 1;
          b = r;
        } else {
          a = (r + team) % (QSP_BIG_PR - 1);
          b = (r + QSP_BIG_PR - 1 - team) % (QSP_BIG_PR - 1);
        }
        const int p = a < b ? a : b, q = a < b ? b : a;
        const double al = Y[p * QSP_BIG_LDY + p].x, be = Y[q * QSP_BIG_LDY + q].x;
        const cplx g = Y[p * QSP_BIG_LDY + q];
        const double g2 = cabs2(g), ab = al * be;
        const bool rot = ab > 1e-280 && g2 > J.tol2 * ab;
        double cs = 1.0, sx = 0.0, sy = 0.0;
        if (rot && tl == 0) jacobi_rotation(al, be, g.x, g.y, g2, cs, sx, sy);
        cs = __shfl_sync(QSP_FULL_MASK, cs, 0, 16);
        sx = __shfl_sync(QSP_FULL_MASK, sx, 0, 16);
        sy = __shfl_sync(QSP_FULL_MASK, sy, 0, 16);
        __syncthreads();  // every pair has read its three entries
        if (rot) {        // rows p, q of [G | J]:  x' = cs x - se y,  y' = conj(se) x + cs y
          cplx* xp = Y + p * QSP_BIG_LDY;
          cplx* yq = Y + q * QSP_BIG_LDY;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const cplx x = xp[tl + 16 * u], y = yq[tl + 16 * u];
            xp[tl + 16 * u] = cmake(cs * x.x - (sx * y.x - sy * y.y), cs * x.y - (sx * y.y + sy * y.x));
            yq[tl + 16 * u] = cmake(cs * y.x + (sx * x.x + sy * x.y), cs * y.y + (sx * x.y - sy * x.x));
          }
          if (tl == 0) *rotflag = 1.0;
        }
        __syncthreads();
        if (rot) {  // columns p, q of G (G R^H):  col_p' = cs col_p - conj(se) col_q,  col_q' = se col_p + cs col_q
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int i = tl + 16 * u;
            const cplx x = Y[i * QSP_BIG_LDY + p], y = Y[i * QSP_BIG_LDY + q];
            Y[i * QSP_BIG_LDY + p] = cmake(cs * x.x - (sx * y.x + sy * y.y), cs * x.y - (sx * y.y - sy * y.x));
            Y[i * QSP_BIG_LDY + q] = cmake(cs * y.x + (sx * x.x - sy * x.y), cs * y.y + (sx * x.y + sy * x.x));
          }
        }
        __syncthreads();
      }
      if (*rotflag == 0.0) break;
      __syncthreads();
    }
    __syncthreads();
    cplx* Jo2 = Jg + (long)task * QSP_BIG_PR * QSP_BIG_PR;
    for (int idx = tid; idx < QSP_BIG_PR * QSP_BIG_PR; idx += QSP_BIG_NT)
      Jo2[idx] = Y[(idx / QSP_BIG_PR) * QSP_BIG_LDY + QSP_BIG_PR + (idx % QSP_BIG_PR)];
    return;
  }

  // ---- Cholesky G = L L^H: L is built in the right half of Y (one barrier per column: the trailing update reads the
  // UNSCALED column j of G and divides by the pivot on the fly); dependent rows give zero columns -------------------
  cplx* Lh = Y + QSP_BIG_PR;  // L[i][j] at Lh[i * LDY + j]
  for (int j = 0; j < QSP_BIG_PR; ++j) {
    const double piv = Y[j * QSP_BIG_LDY + j].x;
    const bool ok = piv > 4e-15 * sd[j] && piv > 0.0;
    double inv = 0.0;  // 1 / sqrt(pivot): one lane per warp computes it (the FP64 pipe is shared with three other CTAs)
    if (lane == 0 && ok) inv = qsp_rsqrt(piv);
    inv = __shfl_sync(QSP_FULL_MASK, inv, 0);
    const double ipiv = inv * inv;
    if (tid < QSP_BIG_PR) {  // column j of L (rows >= j), zeros above the diagonal
      const int i = tid;
      cplx v = cmake(0.0, 0.0);
      if (i == j) v = cmake(ok ? piv * inv : 0.0, 0.0);
      else if (i > j) v = cscale(Y[i * QSP_BIG_LDY + j], inv);
      Lh[i * QSP_BIG_LDY + j] = v;
    }
    if (ok) {
      // trailing update of the lower triangle: thread (i, k-lane) pairs, rows i = j+1.., eight k per row and pass
      const int i = j + 1 + (tid >> 3);
      if (i < QSP_BIG_PR) {
        const cplx gi = Y[i * QSP_BIG_LDY + j];
        for (int k = j + 1 + (tid & 7); k <= i; k += 8) {
          const cplx gk = Y[k * QSP_BIG_LDY + j];
          cplx v = Y[i * QSP_BIG_LDY + k];
          // v -= gi * conj(gk) / piv
          v.x -= (gi.x * gk.x + gi.y * gk.y) * ipiv;
          v.y -= (gi.y * gk.x - gi.x * gk.y) * ipiv;
          Y[i * QSP_BIG_LDY + k] = v;
        }
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < QSP_BIG_PR * QSP_BIG_PR; idx += QSP_BIG_NT) {  // Y <- [L | I]
    const int i = idx / QSP_BIG_PR, k = idx - i * QSP_BIG_PR;
    Y[i * QSP_BIG_LDY + k] = Lh[i * QSP_BIG_LDY + k];
  }
  __syncthreads();
  for (int idx = tid; idx < QSP_BIG_PR * QSP_BIG_PR; idx += QSP_BIG_NT) {
    const int i = idx / QSP_BIG_PR, k = idx - i * QSP_BIG_PR;
    Lh[i * QSP_BIG_LDY + k] = cmake(i == k ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();

  // ---- scalar Jacobi on the rows of L, rotations accumulated in J (right half of Y).  The pairs inside a 16-row block
  // are visited in round 0 of a sweep only (every block is paired in every round); later rounds rotate the 16 x 16 pairs
  // between the two blocks. ---------------------------------------------------------------------------------------
  {
    const int team = tid >> 4, tl = tid & 15;  // 16 pairs x 16 lanes
    volatile double* rotflag = sd + 33;
    const bool bip = round > 0;
    const int nr = bip ? QSP_BIG_RB : QSP_BIG_PR - 1;
    for (int sw = 0; sw < J.inner; ++sw) {
      if (tid == 0) *rotflag = 0.0;
      __syncthreads();
      for (int r = 0; r < nr; ++r) {
        int a, b;
        if (bip) {
          a = team;
          b = QSP_BIG_RB + ((team + r) & (QSP_BIG_RB - 1));
        } else if (team == 0) {
          a = QSP_BIG_PR - 1;
          b = r;
        } else {
          a = (r + team) % (QSP_BIG_PR - 1);
          b = (r + QSP_BIG_PR - 1 - team) % (QSP_BIG_PR - 1);
        }
        const int p = a < b ? a : b, q = a < b ? b : a;
        cplx* xp = Y + p * QSP_BIG_LDY;
        cplx* yq = Y + q * QSP_BIG_LDY;
        cplx x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          x[u] = xp[tl + 16 * u];
          y[u] = yq[tl + 16 * u];
        }
        double al = 0.0, be = 0.0, gx = 0.0, gy = 0.0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          al = fma(x[u].x, x[u].x, fma(x[u].y, x[u].y, al));
          be = fma(y[u].x, y[u].x, fma(y[u].y, y[u].y, be));
          gx = fma(x[u].x, y[u].x, fma(x[u].y, y[u].y, gx));
          gy = fma(x[u].y, y[u].x, fma(-x[u].x, y[u].y, gy));
        }
        for (int o = 8; o > 0; o >>= 1) {
          al += shfl_xor_d(al, o);
          be += shfl_xor_d(be, o);
          gx += shfl_xor_d(gx, o);
          gy += shfl_xor_d(gy, o);
        }
        const double g2 = gx * gx + gy * gy, ab = al * be;
        const bool rot = ab > 1e-280 && g2 > J.tol2 * ab;
        double cs = 1.0, sx = 0.0, sy = 0.0;  // the team leader computes the rotation, the team gets it by shuffle
        if (rot && tl == 0) jacobi_rotation(al, be, gx, gy, g2, cs, sx, sy);
        cs = __shfl_sync(QSP_FULL_MASK, cs, 0, 16);
        sx = __shfl_sync(QSP_FULL_MASK, sx, 0, 16);
        sy = __shfl_sync(QSP_FULL_MASK, sy, 0, 16);
        if (rot) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            cplx xn, yn;
            xn.x = cs * x[u].x - (sx * y[u].x - sy * y[u].y);
            xn.y = cs * x[u].y - (sx * y[u].y + sy * y[u].x);
            yn.x = cs * y[u].x + (sx * x[u].x + sy * x[u].y);
            yn.y = cs * y[u].y + (sx * x[u].y - sy * x[u].x);
            xp[tl + 16 * u] = xn;
            yq[tl + 16 * u] = yn;
          }
          if (tl == 0) *rotflag = 1.0;
        }
        __syncthreads();
      }
      if (*rotflag == 0.0) break;  // (uniform: read after the barrier above)
      __syncthreads();
    }
  }
  __syncthreads();
  cplx* Jo = Jg + (long)task * QSP_BIG_PR * QSP_BIG_PR;
  for (int idx = tid; idx < QSP_BIG_PR * QSP_BIG_PR; idx += QSP_BIG_NT)
    Jo[idx] = Y[(idx / QSP_BIG_PR) * QSP_BIG_LDY + QSP_BIG_PR + (idx % QSP_BIG_PR)];
}

// X <- J X on the columns of split `sp`, J fragments register resident
__global__ void __launch_bounds__(QSP_BIG_NT) k_big_jac_update(const BigJac* __restrict__ jd, int round, int pairs_max, int splits,
                                                               const cplx* __restrict__ Jg, const int* __restrict__ skip) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int task = blockIdx.x / splits, sp = blockIdx.x - task * splits;
  if (skip[task]) return;
  const int inst = task / pairs_max, pi = task - inst * pairs_max;
  const BigJac J = jd[inst];
  const BigTask t = big_task(J, pi, round);
  if (!t.live) return;
  int c0, c1;
  big_split_range(J.len / QSP_BIG_CK, sp, splits, &c0, &c1);
  if (c1 <= c0) return;
  char* smc = QSP_SMEM_BASE();
  cplx* Js = (cplx*)(smc + big_ring_bytes());
  BigRing R;
  R.init(smc, (qsp_mbar*)(Js + QSP_BIG_PR * QSP_BIG_LDJ), J, t);
  int next = c0;
  for (; next < c1 && next < c0 + 2; ++next) R.issue(next);
  const cplx* Ji = Jg + (long)task * QSP_BIG_PR * QSP_BIG_PR;
  for (int idx = tid; idx < QSP_BIG_PR * QSP_BIG_PR; idx += QSP_BIG_NT)
    Js[(idx / QSP_BIG_PR) * QSP_BIG_LDJ + (idx % QSP_BIG_PR)] = Ji[idx];
  __syncthreads();
  const int fr = lane >> 2, fk = lane & 3;
  const int rt0 = (warp >> 2) * 2, ct = warp & 3;
  double jr[2][8], ji[2][8];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const cplx v = Js[((rt0 + u) * 8 + fr) * QSP_BIG_LDJ + ks * 4 + fk];
      jr[u][ks] = v.x;
      ji[u][ks] = v.y;
    }
  int orow[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) orow[u] = big_panel_row((rt0 + u) * 8 + fr, t.ra0, t.na, t.rb0, t.nbr);
  int cdone = c0;
  for (int c = c0; c < c1; c += 2) {
    for (int u = 0; u < 2 && next < c1; ++u) R.issue(next++);
    for (int u = 0; u < 2 && c + u < c1; ++u, ++cdone) {
      const cplx* X = R.wait_next();
      ctile o[2] = {ctile_zero(), ctile_zero()};
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const cplx bq = X[(ks * 4 + fk) * QSP_BIG_LDS + ct * 8 + fr];
        // (re, im) of both row tiles interleaved: four independent accumulators per k-step
        dmma_8x8x4(o[0].r0, o[0].r1, jr[0][ks], bq.x);
        dmma_8x8x4(o[1].r0, o[1].r1, jr[1][ks], bq.x);
        dmma_8x8x4(o[0].i0, o[0].i1, jr[0][ks], bq.y);
        dmma_8x8x4(o[1].i0, o[1].i1, jr[1][ks], bq.y);
        dmma_8x8x4(o[0].r0, o[0].r1, -ji[0][ks], bq.y);
        dmma_8x8x4(o[1].r0, o[1].r1, -ji[1][ks], bq.y);
        dmma_8x8x4(o[0].i0, o[0].i1, ji[0][ks], bq.x);
        dmma_8x8x4(o[1].i0, o[1].i1, ji[1][ks], bq.x);
      }
      R.release();  // (all fragments of the chunk are loaded: the stage may be refilled)
#pragma unroll
      for (int u2 = 0; u2 < 2; ++u2)
        if (orow[u2] >= 0) {
          cplx* dst = J.W + (long)orow[u2] * J.ldw + (long)cdone * QSP_BIG_CK + ct * 8 + 2 * fk;
          dst[0] = cmake(o[u2].r0, o[u2].i0);
          dst[1] = cmake(o[u2].r1, o[u2].i1);
        }
    }
  }
}

// after a sweep: converged instances (no rotation above QSP_JAC_QUAD) are retired, *any counts the others
__global__ void k_big_sweep_end(BigJac* __restrict__ jd, int* __restrict__ flags, int* __restrict__ any, int* __restrict__ sweeps,
                                int B) {
  const int inst = blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= B) return;
  if (jd[inst].nblk > 0) {
    if (sweeps) sweeps[inst] += 1;
    if (flags[inst] == 0)
      jd[inst].nblk = 0;
    else
      atomicAdd(any, 1);
    flags[inst] = 0;
  }
}

// ----------------------------------------------------------------------------------
// right compression sweep, large bonds: per-instance state + per-site helper kernels
// (same mathematics and the same argument block as k_right_compress in qsp_sweeps.cuh)
// ----------------------------------------------------------------------------------
struct RCompressArgs;  // qsp_sweeps.cuh

struct BigRCState {
  int bad, rp, fcur, st;
  // dims of the current site
  int b, c, ro, start, kd, modeC, nrows, ndot, ndotp, len, ldw, joff, k;
  int pad;
};

// scratch slices of one instance (elements): same partition as the small-bond kernel
struct BigRCWork {
  cplx *Mg, *Wg, *Vg, *Cg, *fg;
};
