// One-sided (Hestenes) Jacobi SVD in complex128: the rows of a matrix are rotated in
// pairs until they are mutually orthogonal; the row norms are then the singular values
// and the normalised rows the right singular vectors.
//
// Rows live in shared memory while they are rotated.  A pair of rows is owned by a
// "team" of `tpp` lanes (tpp | 32): the three Gram entries are reduced with warp
// shuffles, the 2x2 unitary is computed redundantly by every lane and applied to the
// team's columns.  Pairs follow the round-robin (circle) tournament so that all teams of
// a round touch disjoint rows.  Matrices with more rows than fit in shared memory are
// processed block-pair by block-pair out of an L2-resident scratch copy.
//
// `ndot` <= `len`: only the first ndot columns enter the Gram entries, the rotation is
// applied to all len columns (used to accumulate the rotations in an appended identity).
#pragma once
#include "qsp_common.cuh"

// A sweep whose largest rotated |g|/sqrt(alpha beta) is <= QSP_JAC_QUAD leaves off-diagonals of
// order QSP_JAC_QUAD^2 ~ 1e-15 (quadratic convergence), i.e. below the rotation threshold 4e-16 sqrt(n)
// that defines convergence anyway: no verification sweep is needed.
#define QSP_JAC_QUAD 3e-8

// One full tournament sweep over `nrows` rows of W (shared memory, row stride ldw).
// Returns the block-wide number of rotations above QSP_JAC_QUAD (0 = converged).  All threads must call.
// n0 > 0: bipartite sweep -- only the pairs (i, j) with i < n0 <= j (rows of two different blocks) are
// visited, max(n0, nrows - n0) rounds of disjoint pairs; n0 = 0: all pairs (circle tournament).
__device__ inline int jacobi_sweep_smem(cplx* W, int nrows, int ndot, int len, int ldw, int tpp, double tol2,
                                        double* red, int n0 = 0) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nteams = nt / tpp, team = tid / tpp, tl = tid % tpp;
  const int np = (nrows + 1) & ~1;  // padded to even
  const int n1 = nrows - n0, nmax = n0 > n1 ? n0 : n1;
  const int npairs = n0 > 0 ? nmax : np / 2, nrounds = n0 > 0 ? nmax : np - 1;
  const int passes = (npairs + nteams - 1) / nteams;
  int nrot = 0;
  for (int r = 0; r < nrounds; ++r) {
    for (int ps = 0; ps < passes; ++ps) {
      const int pi = team + ps * nteams;
      int a = 0, b = 0;
      bool valid = pi < npairs;
      if (valid && n0 > 0) {
        const int jb = (pi + r) % nmax;
        a = pi;
        b = n0 + jb;
        valid = (pi < n0) && (jb < n1);
      } else if (valid) {
        if (pi == 0) {
          a = np - 1;
          b = r;
        } else {
          a = (r + pi) % (np - 1);
          b = (r + np - 1 - pi) % (np - 1);
        }
        valid = (a < nrows) && (b < nrows);
      }
      const int p = a < b ? a : b, q = a < b ? b : a;
      const cplx* xp = W + (long)p * ldw;
      const cplx* yq = W + (long)q * ldw;
      double al = 0.0, be = 0.0, gx = 0.0, gy = 0.0;
      if (valid) {
        for (int c = tl; c < ndot; c += tpp) {
          cplx x = xp[c], y = yq[c];
          al = fma(x.x, x.x, al);
          al = fma(x.y, x.y, al);
          be = fma(y.x, y.x, be);
          be = fma(y.y, y.y, be);
          // g += x * conj(y)
          gx = fma(x.x, y.x, gx);
          gx = fma(x.y, y.y, gx);
          gy = fma(x.y, y.x, gy);
          gy = fma(-x.x, y.y, gy);
        }
      }
      for (int o = tpp >> 1; o > 0; o >>= 1) {
        al += shfl_xor_d(al, o);
        be += shfl_xor_d(be, o);
        gx += shfl_xor_d(gx, o);
        gy += shfl_xor_d(gy, o);
      }
      const double g2 = gx * gx + gy * gy;
      const double ab = al * be;
      if (valid && ab > 1e-280 && g2 > tol2 * ab) {
        const double ag = sqrt(g2);
        const double zeta = (be - al) / (2.0 * ag);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + t * t);
        const double sn = cs * t;
        const double ex = gx / ag, ey = gy / ag;  // e^{i phi}
        const cplx se = cmake(sn * ex, sn * ey);  // s e^{i phi}
        cplx* xw = W + (long)p * ldw;
        cplx* yw = W + (long)q * ldw;
        for (int c = tl; c < len; c += tpp) {
          cplx x = xw[c], y = yw[c];
          // x' = c x - s e y ; y' = s conj(e) x + c y
          cplx xn, yn;
          xn.x = cs * x.x - (se.x * y.x - se.y * y.y);
          xn.y = cs * x.y - (se.x * y.y + se.y * y.x);
          yn.x = cs * y.x + (se.x * x.x + se.y * x.y);
          yn.y = cs * y.y + (se.x * x.y - se.y * x.x);
          xw[c] = xn;
          yw[c] = yn;
        }
        if (tl == 0 && g2 > (QSP_JAC_QUAD * QSP_JAC_QUAD) * ab) ++nrot;  // rotations below it leave ~1e-15: no verification sweep
      }
    }
    __syncthreads();
  }
  return (int)(block_sum((double)nrot, red) + 0.5);
}

__device__ inline void copy_rows(cplx* dst, long ldd, const cplx* src, long lds, int nrows, int len) {
  const int total = nrows * len;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    int r = idx / len, c = idx - r * len;
    dst[(long)r * ldd + c] = src[(long)r * lds + c];
  }
}

// Orthogonalise the rows of Wg (nrows x len, row stride len, global scratch) in place.
// Wsm: shared buffer that holds cap_rows rows of len columns.  Returns true on
// convergence.  All threads must call.
__device__ inline bool jacobi_rows(cplx* Wg, int nrows, int ndot, int len, cplx* Wsm, int cap_rows, int tpp,
                                   double tol2, int max_sweeps, double* red, int* sweeps_out) {
  bool conv = false;
  int sweeps = 0;
  if (nrows <= 1) {
    if (sweeps_out) *sweeps_out = 0;
    return true;
  }
  // team width per stage: as wide as the number of concurrent pairs allows (the caller's hint is
  // sized for the whole matrix, a block pair has far fewer pairs than the CTA has lanes)
  auto team_width = [&](int rows) {
    const int pairs = (rows + 1) >> 1;
    int t = 32;
    while (t > 1 && ((int)blockDim.x / t < pairs || t > ndot)) t >>= 1;
    return t;
  };
  (void)tpp;
  if (nrows <= cap_rows) {
    tpp = team_width(nrows);
    copy_rows(Wsm, len, Wg, len, nrows, len);
    __syncthreads();
    while (sweeps < max_sweeps) {
      int nrot = jacobi_sweep_smem(Wsm, nrows, ndot, len, len, tpp, tol2, red);
      ++sweeps;
      if (nrot == 0) {
        conv = true;
        break;
      }
    }
    __syncthreads();
    copy_rows(Wg, len, Wsm, len, nrows, len);
    __syncthreads();
  } else {
    int rb = cap_rows / 2;
    const int nblk = (nrows + rb - 1) / rb;
    rb = (nrows + nblk - 1) / nblk;  // balance the blocks
    tpp = team_width(2 * rb);
    while (sweeps < max_sweeps) {
      int nrot = 0;
      // pairs inside a block once per sweep ...
      for (int bi = 0; bi < nblk; ++bi) {
        const int r0 = bi * rb, n0 = imin(rb, nrows - r0);
        if (n0 < 2) continue;
        copy_rows(Wsm, len, Wg + (long)r0 * len, len, n0, len);
        __syncthreads();
        nrot += jacobi_sweep_smem(Wsm, n0, ndot, len, len, tpp, tol2, red);
        __syncthreads();
        copy_rows(Wg + (long)r0 * len, len, Wsm, len, n0, len);
        __syncthreads();
      }
      // ... and the pairs between two blocks once per block pair
      for (int bi = 0; bi < nblk; ++bi) {
        for (int bj = bi + 1; bj < nblk; ++bj) {
          const int r0 = bi * rb, n0 = imin(rb, nrows - r0);
          const int r1 = bj * rb, n1 = imin(rb, nrows - r1);
          copy_rows(Wsm, len, Wg + (long)r0 * len, len, n0, len);
          copy_rows(Wsm + (long)n0 * len, len, Wg + (long)r1 * len, len, n1, len);
          __syncthreads();
          nrot += jacobi_sweep_smem(Wsm, n0 + n1, ndot, len, len, tpp, tol2, red, n0);
          __syncthreads();
          copy_rows(Wg + (long)r0 * len, len, Wsm, len, n0, len);
          copy_rows(Wg + (long)r1 * len, len, Wsm + (long)n0 * len, len, n1, len);
          __syncthreads();
        }
      }
      ++sweeps;
      if (nrot == 0) {
        conv = true;
        break;
      }
    }
  }
  if (sweeps_out) *sweeps_out = sweeps;
  return conv;
}

// ----------------------------------------------------------------------------------
// Register-resident odd-even Jacobi (the fast path).
//
// Rows are laid out on a line; rounds alternate between the pairs (2t, 2t+1) and
// (2t+1, 2t+2) and the two rows of a pair swap places after every visit, so n rounds
// visit each of the n(n-1)/2 pairs exactly once (odd-even transposition ordering).
// Team t keeps the row at the odd position 2t+1 in REGISTERS for the whole SVD; the rows at
// even positions live in shared memory.  A visit therefore costs two shared reads and one
// shared write of a row instead of four reads + two writes, and a 128 x 128 complex128
// matrix (256 KB) stays on chip: 128 KB in registers, 128 KB in shared memory.
// Lane tl of a team owns columns tl, tl+tpp, ... ; rows are zero-padded to a multiple of
// tpp columns (ldw = jlen*tpp) and the Gram part to jdot*tpp columns, so the column loops
// are branch-free and warp-uniform.
// ----------------------------------------------------------------------------------
// Ablation switches for timing experiments (scripts/jacobi_bench.py, results invalid when set):
// bit0 no dot, bit1 no shuffles, bit2 no scalar chain, bit3 no rotation update, bit4 no barrier.
// Compiled in only with -DQSP_JACOBI_ABLATION; the product build folds them away.
#ifdef QSP_JACOBI_ABLATION
__device__ int g_jvar = 0;
#define QSP_JV() (g_jvar)
#else
#define QSP_JV() (0)
#endif
#ifndef QSP_JCH
#define QSP_JCH 2  // columns of the shared row fetched per batch in the streaming (non-HOLD) path
#endif
#define QSP_JPAD 32  // rows are zero-padded to a multiple of this many columns

// Team geometry for an (nrows x ldw) problem (ldw a multiple of QSP_JPAD): tpp lanes per team,
// np pairs per team, cpl = ldw / tpp columns per lane (<= 16; up to 8 both rows of a visit are
// held in registers, beyond that the shared row is read twice).  Returns false when the register path cannot take the problem.
__host__ __device__ inline bool jacobi_oe_geometry(int nt, int nrows, int ldw, int* tpp_out, int* np_out) {
  const int pairs = nrows >> 1;
  if (ldw % QSP_JPAD) return false;
  int tpp = 32;
  while (tpp > 1 && nt / tpp < pairs) tpp >>= 1;  // one pair per team when possible
  int np = 1;
  if (nt / tpp < pairs) return false;
  while (ldw / tpp > 16 || ldw % tpp) {  // too many columns per lane: widen the team, two pairs per team
    if (tpp == 32) return false;
    tpp <<= 1;
  }
  while ((nt / tpp) * np < pairs) ++np;
  if (np > 2 || (np == 2 && ldw / tpp > 8)) return false;
  *tpp_out = tpp;
  *np_out = np;
  return true;
}

// rotation parameters from the Gram entries: x' = cs x - se y, y' = conj(se) x + cs y
// with se = s e^{i phi}.  With d = beta - alpha, h = sqrt(d^2 + 4|g|^2), u = h + |d|:
//   cs = sqrt(u / 2h) = (u/2) sw,   s e^{i phi} = sign(d) sw g,   sw = rsqrt(h u / 2)   (cs^2 + s^2 = 1 identically)
// -- two reciprocal square roots and no division or square root on the dependent chain of a round
// (h = q rsqrt(q)); the tangent, where needed, is t e^{i phi} = w g with w = sign(d) 2/u = sign(d) sw^2 h.
__device__ __forceinline__ void jacobi_rotation_sw(double d, double g2, double& cs, double& sws, double& h) {
  const double q = fma(d, d, 4.0 * g2);
  h = q * qsp_rsqrt(q);
  const double u = h + fabs(d);
  const double sw = qsp_rsqrt(0.5 * h * u);
  const bool tiny = !(q > 1e-290);  // both rows ~1e-70 and below (q under/near the denormals): leave them alone
  cs = tiny ? 1.0 : 0.5 * u * sw;
  sws = tiny ? 0.0 : (d >= 0.0 ? sw : -sw);
}
__device__ __forceinline__ void jacobi_rotation(double al, double be, double gx, double gy, double g2, double& cs,
                                                double& sx, double& sy) {
  double sws, h;
  jacobi_rotation_sw(be - al, g2, cs, sws, h);
  sx = sws * gx;
  sy = sws * gy;
}

// TPPC > 0 fixes the team width at compile time and FULL asserts jdot == jlen == CPL (no column
// predicates): the square power-of-two cases of the layer loop get immediate shared-memory offsets
// and noticeably lower register pressure.
//
// Scaled ("fast") rotations: every row is stored as d * x~ with a real scale d kept beside it
// (register rows: dr[], shared rows: dS[slot]).  With t e^{i phi} = w g the rotation
//     x' = c (x - w g y),  y' = c (y + w conj(g) x)
// becomes  x~' = x~ - (w dr^2 g~) y~,  d_x' = c ds ;   y~' = y~ + (w ds^2 conj(g~)) x~,  d_y' = c dr,
// i.e. two complex multiply-adds per column (8 DFMA) instead of 12 flops-heavy instructions.  The
// convergence test |g|^2 <= tol^2 alpha beta is scale free.  Scales are folded back into the data at
// the end of every sweep (they only shrink, by c >= 1/sqrt 2 per rotation).
// They are used where both rows of a visit are register resident (HOLD); the 16-columns-per-lane
// variant sits at the 128-register limit and keeps plain rotations (the extra scalars spill).
// Optional input / output of the register path (saves two passes over the matrix in the caller):
//   src        rows are loaded from here instead of Wg (row stride src_ld, src_cols live columns, zero beyond);
//              the orthogonalised rows are still written to Wg
//   norms_off  shared-memory offset of a double[nrows] that receives the SQUARED row norms over the Gram columns
//              (< 0: none); norms_done is set when they were written
//   tmem       base address of >= 192 allocated tensor-memory columns (tmem_ok set): the 128 x 128 problem keeps
//              48 of its 64 shared rows there (jacobi_rows_oe_tmem)
struct JacobiIO {
  const cplx* src;
  int src_ld, src_cols;
  int norms_off;
  bool norms_done;
  uint32_t tmem;
  bool tmem_ok;
  bool presort;  // lay the rows out by decreasing norm before every sweep (uses the norms_off area as scratch)
};

template <int CPL, int NP, bool HOLD, int TPPC = 0, bool FULL = false, int FASTSEL = -1>
__device__ __noinline__ bool jacobi_rows_oe(cplx* Wg, int nrows, int jdot_, int jlen_, int s_off, int tpp_, double tol2,
                                            int max_sweeps, int red_off, int* sweeps_out, JacobiIO* io = nullptr) {
  constexpr bool FAST = FASTSEL < 0 ? HOLD : (FASTSEL != 0);
  const int tpp = TPPC > 0 ? TPPC : tpp_;
  const int jlen = FULL ? CPL : jlen_;
  const int jdot = FULL ? CPL : jdot_;
  cplx* S = (cplx*)(QSP_SMEM_BASE() + s_off);
  double* red = (double*)(QSP_SMEM_BASE() + red_off);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int team = tid / tpp, tl = tid % tpp;
  const int ldw = jlen * tpp;
  const int ne = (nrows + 1) >> 1, no = nrows >> 1;
  double* dS = (double*)(S + (long)ne * ldw);  // scales of the shared (even-position) rows
  cplx R[NP][CPL];
  double dr[NP];
  {
    const bool ext = io && io->src;
    const cplx* src = ext ? io->src : Wg;
    const long sld = ext ? io->src_ld : ldw;
    const int scols = ext ? io->src_cols : ldw;
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const int p = team * NP + q;
      dr[q] = 1.0;
#pragma unroll
      for (int j = 0; j < CPL; ++j)
        R[q][j] = (p < no && j < jlen && tl + j * tpp < scols) ? src[(long)(2 * p + 1) * sld + tl + j * tpp] : cmake(0.0, 0.0);
    }
    for (int idx = tid; idx < ne * ldw; idx += nt) {
      const int r = idx / ldw, c = idx - r * ldw;
      S[idx] = (c < scols) ? src[(long)(2 * r) * sld + c] : cmake(0.0, 0.0);
    }
  }
  for (int idx = tid; idx < ne; idx += nt) dS[idx] = 1.0;
  __syncthreads();
  bool conv = false;
  int sweeps = 0;
  volatile int* flag = (volatile int*)(red + 48);  // set when a rotation above QSP_JAC_QUAD happened
  const int jv = QSP_JV();
  // Norm ordering: before every sweep the rows are laid out along the ring of the odd-even ordering by decreasing norm
  // (a permutation of the rows: the SVD is unchanged, the row order of the result is whatever the last sweep leaves,
  // as before).  Simulated and measured on 128 x 128 problems: ~15 % fewer sweeps on full-rank random matrices
  // (10.7 -> 9.0) and ~30 % fewer on graded ones (18 -> 13 at four decades); a sweep costs 128 rounds, the permutation
  // about one (norms, ranks by counting, one trip of the rows through Wg, which holds nothing the sweeps need).
  // Needs a double[nrows] scratch: the norms_off area.
  // (not for the few-row problems of the D=2 sweeps: there the permutation costs as much as the sweep)
  double* snrm = (io && io->norms_off >= 0 && io->presort && nrows >= 16) ? (double*)(QSP_SMEM_BASE() + io->norms_off) : nullptr;
  while (sweeps < max_sweeps) {
    if (snrm) {
      // squared norms over the Gram columns (scales are 1 here: folded at the end of every sweep)
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        const int p = team * NP + q;
        double n2 = 0.0;
#pragma unroll
        for (int j = 0; j < CPL; ++j)
          if (j < jdot) n2 = fma(R[q][j].x, R[q][j].x, fma(R[q][j].y, R[q][j].y, n2));
        for (int o = tpp >> 1; o > 0; o >>= 1) n2 += shfl_xor_d(n2, o);
        if (tl == 0 && p < no) snrm[2 * p + 1] = (n2 == n2) ? n2 : 0.0;
      }
      for (int s0 = 0; s0 < ne; s0 += nt / tpp) {  // warp-uniform trip count (shuffles inside)
        const int slot = s0 + team;
        const cplx* xr = S + (long)(slot < ne ? slot : 0) * ldw + tl;
        double n2 = 0.0;
        for (int j = 0; j < jdot; ++j) {
          const cplx v = xr[j * tpp];
          n2 = fma(v.x, v.x, fma(v.y, v.y, n2));
        }
        for (int o = tpp >> 1; o > 0; o >>= 1) n2 += shfl_xor_d(n2, o);
        if (tl == 0 && slot < ne) snrm[2 * slot] = (n2 == n2) ? n2 : 0.0;
      }
      __syncthreads();
      // new position of row i = number of rows ahead of it (larger norm; ties by index): a permutation for any data
      auto rank_of = [&](int i) {
        const double ni = snrm[i];
        int cnt = 0;
        for (int k = tl; k < nrows; k += tpp) {
          const double nk = snrm[k];
          cnt += (nk > ni || (nk == ni && k < i)) ? 1 : 0;
        }
        for (int o = tpp >> 1; o > 0; o >>= 1) cnt += __shfl_xor_sync(QSP_FULL_MASK, cnt, o);
        // rank -> place on the line: the larger half on the odd places (the register rows) in decreasing order from
        // the left, the smaller half on the even places in increasing order from the left, i.e. decreasing norm along
        // the REVERSED ring of the odd-even ordering (1, 3, ..., n-1, n-2, ..., 2, 0)
        return cnt < no ? 2 * cnt + 1 : 2 * (nrows - 1 - cnt);
      };
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        const int p = team * NP + q;
        const int pos = rank_of(p < no ? 2 * p + 1 : 0);
        if (p < no) {
#pragma unroll
          for (int j = 0; j < CPL; ++j)
            if (j < jlen) Wg[(long)pos * ldw + tl + j * tpp] = R[q][j];
        }
      }
      for (int s0 = 0; s0 < ne; s0 += nt / tpp) {
        const int slot = s0 + team;
        const int pos = rank_of(slot < ne ? 2 * slot : 0);
        if (slot < ne) {
          const cplx* xr = S + (long)slot * ldw + tl;
          for (int j = 0; j < jlen; ++j) Wg[(long)pos * ldw + tl + j * tpp] = xr[j * tpp];
        }
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        const int p = team * NP + q;
#pragma unroll
        for (int j = 0; j < CPL; ++j)
          if (p < no && j < jlen) R[q][j] = Wg[(long)(2 * p + 1) * ldw + tl + j * tpp];
      }
      for (int idx = tid; idx < ne * ldw; idx += nt) {
        const int r = idx / ldw, c = idx - r * ldw;
        S[idx] = Wg[(long)(2 * r) * ldw + c];
      }
      __syncthreads();
    }
    if (tid == 0) *flag = 0;
    __syncthreads();
    for (int r = 0; r < nrows; ++r) {
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        const int p = team * NP + q;
        const int slot = p + (r & 1);
        const bool valid = (p < no) && (slot < ne);
        cplx* xs = S + (long)(valid ? slot : 0) * ldw + tl;
        const double ds = FAST ? dS[valid ? slot : 0] : 1.0;
        cplx x[HOLD ? CPL : 1];
        if (HOLD) {
#pragma unroll
          for (int j = 0; j < CPL; ++j)
            if (j < jlen) x[HOLD ? j : 0] = xs[j * tpp];
        }
        double al = 0.0, be = 0.0, gx = 0.0, gy = 0.0;  // Gram entries of the scaled rows
        if (HOLD) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            if (j < jdot && !(jv & 1)) {
              const cplx y = R[q][j];
              const cplx xv = x[HOLD ? j : 0];
              al = fma(xv.x, xv.x, al);
              al = fma(xv.y, xv.y, al);
              be = fma(y.x, y.x, be);
              be = fma(y.y, y.y, be);
              gx = fma(xv.x, y.x, gx);
              gx = fma(xv.y, y.y, gx);
              gy = fma(xv.y, y.x, gy);
              gy = fma(-xv.x, y.y, gy);
            }
          }
        } else {
          // the shared row is streamed in chunks of QSP_JCH columns: the loads of a chunk are issued
          // together (independent registers), then consumed -- one shared-memory latency per chunk
#pragma unroll
          for (int j0 = 0; j0 < CPL; j0 += QSP_JCH) {
            cplx xt[QSP_JCH];
#pragma unroll
            for (int u = 0; u < QSP_JCH; ++u)
              if (j0 + u < jdot && !(jv & 1)) xt[u] = xs[(j0 + u) * tpp];
#pragma unroll
            for (int u = 0; u < QSP_JCH; ++u) {
              if (j0 + u < jdot && !(jv & 1)) {
                const cplx y = R[q][j0 + u];
                const cplx xv = xt[u];
                al = fma(xv.x, xv.x, al);
                al = fma(xv.y, xv.y, al);
                be = fma(y.x, y.x, be);
                be = fma(y.y, y.y, be);
                gx = fma(xv.x, y.x, gx);
                gx = fma(xv.y, y.y, gx);
                gy = fma(xv.y, y.x, gy);
                gy = fma(-xv.x, y.y, gy);
              }
            }
          }
        }
        if (!(jv & 2))
          for (int o = tpp >> 1; o > 0; o >>= 1) {
            al += shfl_xor_d(al, o);
            be += shfl_xor_d(be, o);
            gx += shfl_xor_d(gx, o);
            gy += shfl_xor_d(gy, o);
          }
        if (FAST) __syncwarp();  // every lane of the team has read dS[slot] before it is rewritten below
        if (valid) {
          const double g2t = gx * gx + gy * gy;
          const double abt = al * be;
          if ((jv & 8) ? false : ((jv & 4) ? true : (abt > 1e-280 && g2t > tol2 * abt))) {
            // true Gram entries: alpha = ds^2 al, beta = dr^2 be, |g|^2 = ds^2 dr^2 |g~|^2
            const double ds2 = FAST ? ds * ds : 1.0, dr2 = FAST ? dr[q] * dr[q] : 1.0;
            const double A = ds2 * al, Bq = dr2 * be, g2 = ds2 * dr2 * g2t;
            double cs, sws, h;
            jacobi_rotation_sw(Bq - A, g2, cs, sws, h);
            // FAST: tau1 = w dr^2 g~, tau2 = w ds^2 conj(g~) = (t2x, -t2y), w = sws sw h;  plain: s e^{i phi} = sws g
            const double wt = FAST ? sws * fabs(sws) * h : sws;
            const double w1 = FAST ? wt * dr2 : wt, w2 = FAST ? wt * ds2 : wt;
            const double t1x = w1 * gx, t1y = w1 * gy;
            const double t2x = w2 * gx, t2y = w2 * gy;
            const double c0 = FAST ? 1.0 : cs;  // plain rotations scale the kept term by c
            if (tl == 0 && g2t > (QSP_JAC_QUAD * QSP_JAC_QUAD) * abt) *flag = 1;
            if (HOLD) {
#pragma unroll
              for (int j = 0; j < CPL; ++j) {
                if (j < jlen) {
                  const cplx y = R[q][j];
                  const cplx xv = x[HOLD ? j : 0];
                  cplx xn, yn;
                  xn.x = fma(-t1x, y.x, fma(t1y, y.y, FAST ? xv.x : c0 * xv.x));
                  xn.y = fma(-t1x, y.y, fma(-t1y, y.x, FAST ? xv.y : c0 * xv.y));
                  yn.x = fma(t2x, xv.x, fma(t2y, xv.y, FAST ? y.x : c0 * y.x));
                  yn.y = fma(t2x, xv.y, fma(-t2y, xv.x, FAST ? y.y : c0 * y.y));
                  xs[j * tpp] = yn;  // the rows trade places
                  R[q][j] = xn;
                }
              }
            } else {
#pragma unroll
              for (int j0 = 0; j0 < CPL; j0 += QSP_JCH) {
                cplx xt[QSP_JCH];
#pragma unroll
                for (int u = 0; u < QSP_JCH; ++u)
                  if (j0 + u < jlen) xt[u] = xs[(j0 + u) * tpp];
#pragma unroll
                for (int u = 0; u < QSP_JCH; ++u) {
                  if (j0 + u < jlen) {
                    const cplx y = R[q][j0 + u];
                    const cplx xv = xt[u];
                    cplx xn, yn;
                    xn.x = fma(-t1x, y.x, fma(t1y, y.y, FAST ? xv.x : c0 * xv.x));
                    xn.y = fma(-t1x, y.y, fma(-t1y, y.x, FAST ? xv.y : c0 * xv.y));
                    yn.x = fma(t2x, xv.x, fma(t2y, xv.y, FAST ? y.x : c0 * y.x));
                    yn.y = fma(t2x, xv.y, fma(-t2y, xv.x, FAST ? y.y : c0 * y.y));
                    xs[(j0 + u) * tpp] = yn;  // the rows trade places
                    R[q][j0 + u] = xn;
                  }
                }
              }
            }
            if (FAST) {
              if (tl == 0) dS[slot] = cs * dr[q];  // scale of the row that moved to shared memory (y')
              dr[q] = cs * ds;                     // scale of the row that moved to registers (x')
            }
          } else {
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
              if (j < jlen) {
                const cplx xv = HOLD ? x[HOLD ? j : 0] : xs[j * tpp];
                xs[j * tpp] = R[q][j];
                R[q][j] = xv;
              }
            }
            if (FAST) {
              if (tl == 0) dS[slot] = dr[q];
              dr[q] = ds;
            }
          }
        }
      }
      if (!(jv & 16)) __syncthreads();
    }
    if (FAST) {  // fold the scales back into the data (keeps them away from underflow)
#pragma unroll
      for (int q = 0; q < NP; ++q) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) R[q][j] = cscale(R[q][j], dr[q]);
        dr[q] = 1.0;
      }
      for (int idx = tid; idx < ne * ldw; idx += nt) S[idx] = cscale(S[idx], dS[idx / ldw]);
      __syncthreads();
      for (int idx = tid; idx < ne; idx += nt) dS[idx] = 1.0;
    }
    ++sweeps;
    if (jv) {
      __syncthreads();
      if (sweeps >= 8) break;
      continue;
    }
    if (*flag == 0) {  // (ordered by the barrier above)
      conv = true;
      break;
    }
    __syncthreads();
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const int p = team * NP + q;
#pragma unroll
    for (int j = 0; j < CPL; ++j)
      if (p < no && j < jlen) Wg[(long)(2 * p + 1) * ldw + tl + j * tpp] = R[q][j];
  }
  for (int idx = tid; idx < ne * ldw; idx += nt) {
    const int r = idx / ldw, c = idx - r * ldw;
    Wg[(long)(2 * r) * ldw + c] = S[idx];
  }
  if (io && io->norms_off >= 0) {
    // squared row norms over the Gram columns while the rows are still on chip
    double* nrm = (double*)(QSP_SMEM_BASE() + io->norms_off);
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const int p = team * NP + q;
      double n2 = 0.0;
#pragma unroll
      for (int j = 0; j < CPL; ++j)
        if (j < jdot) n2 = fma(R[q][j].x, R[q][j].x, fma(R[q][j].y, R[q][j].y, n2));
      for (int o = tpp >> 1; o > 0; o >>= 1) n2 += shfl_xor_d(n2, o);
      if (tl == 0 && p < no) nrm[2 * p + 1] = n2;
    }
    for (int s0 = 0; s0 < ne; s0 += nt / tpp) {  // warp-uniform trip count (shuffles inside)
      const int slot = s0 + team;
      const cplx* xs = S + (long)(slot < ne ? slot : 0) * ldw + tl;
      double n2 = 0.0;
      for (int j = 0; j < jdot; ++j) {
        const cplx v = xs[j * tpp];
        n2 = fma(v.x, v.x, fma(v.y, v.y, n2));
      }
      for (int o = tpp >> 1; o > 0; o >>= 1) n2 += shfl_xor_d(n2, o);
      if (tl == 0 && slot < ne) nrm[2 * slot] = n2;
    }
    io->norms_done = true;  // (io is per-thread: every thread records it)
  }
  __syncthreads();
  if (sweeps_out) *sweeps_out = sweeps;
  return conv;
}

#ifdef QSP_DEV_BUILD
// ----------------------------------------------------------------------------------
// EXPERIMENT (development build only, scripts/jacobi_bench.py with QSP_DEBUG_FLAGS=2; profiles/r2_tmem_jacobi.md):
// the 128 x 128 problem (bulk sites of the chi = 64 layer loop) with three quarters of the shared rows in
// TENSOR MEMORY.  The round of jacobi_rows_oe is bound by its three shared-memory passes over the 64 rows at
// even positions (3 x 128 KB at ~125 B/clk, profiles/r1_ncu_right_compress.md); tcgen05.ld / tcgen05.st move
// the same bytes at ~300 B/clk on a pipe of their own (scripts/tmem_probe.py), but a thread only reaches its own
// TMEM lane, and only warps w, w+4, w+8, w+12 share a quadrant.  So the chain of teams is threaded through the
// CTA such that neighbours sit on the SAME lanes of the same quadrant in consecutive warps:
//     chain position p = ((quadrant * 4 + lane group) * 4 + warp / 4)
// and the row that teams p-1 and p hand back and forth (slot p) lives in the TMEM lanes they share when
// p % 4 != 0 (48 slots, 64 columns each) and in shared memory where the chain changes lanes (16 slots, 32 KB).
// Same arithmetic, ordering and convergence rule as jacobi_rows_oe<16, 1, false, 8, true>.
#include "qsp_tmem.cuh"
// timing ablations of the round loop (results invalid when set; development builds only)
#ifdef QSP_TJ_NOWAITLD
#define QSP_TJ_WAIT_LD()
#else
#define QSP_TJ_WAIT_LD() tmem_wait_ld()
#endif
#ifdef QSP_TJ_NOWAITST
#define QSP_TJ_WAIT_ST()
#else
#define QSP_TJ_WAIT_ST() tmem_wait_st()
#endif
#ifdef QSP_TJ_NOFENCE
#define QSP_TJ_FENCE_B()
#define QSP_TJ_FENCE_A()
#else
#define QSP_TJ_FENCE_B() tmem_fence_before()
#define QSP_TJ_FENCE_A() tmem_fence_after()
#endif
__device__ __noinline__ bool jacobi_rows_oe_tmem(cplx* Wg, int s_off, double tol2, int max_sweeps, int red_off,
                                                 int* sweeps_out, JacobiIO* io) {
  constexpr int CPL = 16, TPP = 8, NR = 128, LDW = 128, NSLOT = 64;
  cplx* S = (cplx*)(QSP_SMEM_BASE() + s_off);
  double* red = (double*)(QSP_SMEM_BASE() + red_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 3, tl = lane & 7, quad = warp & 3, jj = warp >> 2;
  const int p = (quad * 4 + grp) * 4 + jj;  // chain position: register row 2p+1, own slot p
  const uint32_t tq = io->tmem + ((uint32_t)(quad * 32) << 16);
  cplx R[CPL];
  {
    const bool ext = io->src != nullptr;
    const cplx* src = ext ? io->src : Wg;
    const long sld = ext ? io->src_ld : LDW;
    const int scols = ext ? io->src_cols : LDW;
    const cplx* ro = src + (long)(2 * p + 1) * sld;
    const cplx* re = src + (long)(2 * p) * sld;
#pragma unroll
    for (int j = 0; j < CPL; ++j) R[j] = (tl + j * TPP < scols) ? ro[tl + j * TPP] : cmake(0.0, 0.0);
    if (jj == 0) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) S[(p >> 2) * LDW + tl + j * TPP] = (tl + j * TPP < scols) ? re[tl + j * TPP] : cmake(0.0, 0.0);
    } else {
      const uint32_t ta = tq + (uint32_t)((jj - 1) * 64);
#pragma unroll
      for (int j = 0; j < CPL; j += 2) {
        const cplx a = (tl + j * TPP < scols) ? re[tl + j * TPP] : cmake(0.0, 0.0);
        const cplx b = (tl + (j + 1) * TPP < scols) ? re[tl + (j + 1) * TPP] : cmake(0.0, 0.0);
        tmem_st2(ta + 4 * j, a, b);
      }
      tmem_wait_st();
    }
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  bool conv = false;
  int sweeps = 0;
  volatile int* flag = (volatile int*)(red + 48);  // set when a rotation above QSP_JAC_QUAD happened
  while (sweeps < max_sweeps) {
    if (tid == 0) *flag = 0;
    __syncthreads();
    for (int r = 0; r < NR; ++r) {
      const int odd = r & 1;
      const int slot = p + odd;
      const bool in_tmem = odd ? (jj != 3) : (jj != 0);  // warp uniform
      const bool valid = slot < NSLOT;  // only the last team idles, in the odd rounds (a shared-memory warp)
      const uint32_t ta = tq + (uint32_t)((odd ? jj : jj - 1) * 64);
      cplx* xs = S + ((valid ? slot : 0) >> 2) * LDW + tl;
      // one instruction stream for both homes of the partner row; the chunk after the current one is in flight
      // while the current one is consumed (tcgen05.wait::ld waits for everything outstanding)
      cplx xa, xb;
#ifdef QSP_TJ_NOTMEM
#define QSP_TJ_LD2(a, x, y) (x = R[0], y = R[1])
#define QSP_TJ_ST2(a, x, y)
#else
#define QSP_TJ_LD2(a, x, y) tmem_ld2(a, x, y)
#define QSP_TJ_ST2(a, x, y) tmem_st2(a, x, y)
#endif
#define QSP_TJ_LOAD(j)                    \
  if (in_tmem) {                          \
    QSP_TJ_LD2(ta + 4 * (j), xa, xb);     \
  } else {                                \
    xa = xs[(j) * TPP];                   \
    xb = xs[((j) + 1) * TPP];             \
  }
      double al = 0.0, be = 0.0, gx = 0.0, gy = 0.0;
      QSP_TJ_LOAD(0);
#pragma unroll
      for (int j0 = 0; j0 < CPL; j0 += 2) {
        if (in_tmem) QSP_TJ_WAIT_LD();
        const cplx xt[2] = {xa, xb};
        if (j0 + 2 < CPL) { QSP_TJ_LOAD(j0 + 2); }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const cplx y = R[j0 + u];
          const cplx xv = xt[u];
          al = fma(xv.x, xv.x, al);
          al = fma(xv.y, xv.y, al);
          be = fma(y.x, y.x, be);
          be = fma(y.y, y.y, be);
          gx = fma(xv.x, y.x, gx);
          gx = fma(xv.y, y.y, gx);
          gy = fma(xv.y, y.x, gy);
          gy = fma(-xv.x, y.y, gy);
        }
      }
#pragma unroll
      for (int o = TPP >> 1; o > 0; o >>= 1) {
        al += shfl_xor_d(al, o);
        be += shfl_xor_d(be, o);
        gx += shfl_xor_d(gx, o);
        gy += shfl_xor_d(gy, o);
      }
      const double g2 = gx * gx + gy * gy, ab = al * be;
      const bool rot = valid && ab > 1e-280 && g2 > tol2 * ab;
      if (__any_sync(QSP_FULL_MASK, rot || !valid)) {
        // branch free (tcgen05.ld / st are warp-aligned: no divergence may reach them); teams that do not rotate
        // take the identity, i.e. their rows just trade places
        double cs, sws, h;
        jacobi_rotation_sw(be - al, g2, cs, sws, h);
        cs = rot ? cs : 1.0;
        sws = rot ? sws : 0.0;
        if (rot && tl == 0 && g2 > (QSP_JAC_QUAD * QSP_JAC_QUAD) * ab) *flag = 1;
        __syncwarp();
        double tx = sws * gx, ty = sws * gy;  // s e^{i phi}
        if (!valid) cs = 0.0, tx = -1.0, ty = 0.0;  // the idle team keeps its register row: x' = y (nothing is stored)
        QSP_TJ_LOAD(0);
#pragma unroll
        for (int j0 = 0; j0 < CPL; j0 += 2) {
          if (in_tmem) QSP_TJ_WAIT_LD();
          const cplx xt[2] = {xa, xb};
          if (j0 + 2 < CPL) { QSP_TJ_LOAD(j0 + 2); }
          cplx yn[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const cplx y = R[j0 + u];
            const cplx xv = xt[u];
            cplx xn;
            xn.x = fma(-tx, y.x, fma(ty, y.y, cs * xv.x));
            xn.y = fma(-tx, y.y, fma(-ty, y.x, cs * xv.y));
            yn[u].x = fma(tx, xv.x, fma(ty, xv.y, cs * y.x));
            yn[u].y = fma(tx, xv.y, fma(-ty, xv.x, cs * y.y));
            R[j0 + u] = xn;
          }
          if (in_tmem) {  // the rows trade places
            QSP_TJ_ST2(ta + 4 * j0, yn[0], yn[1]);
          } else if (valid) {
            xs[j0 * TPP] = yn[0];
            xs[(j0 + 1) * TPP] = yn[1];
          }
        }
      } else {
        QSP_TJ_LOAD(0);
#pragma unroll
        for (int j0 = 0; j0 < CPL; j0 += 2) {
          if (in_tmem) QSP_TJ_WAIT_LD();
          const cplx xt[2] = {xa, xb};
          if (j0 + 2 < CPL) { QSP_TJ_LOAD(j0 + 2); }
          if (in_tmem) {
            QSP_TJ_ST2(ta + 4 * j0, R[j0], R[j0 + 1]);
          } else if (valid) {
            xs[j0 * TPP] = R[j0];
            xs[(j0 + 1) * TPP] = R[j0 + 1];
          }
          R[j0] = xt[0];  // (every team of the warp is valid on this path)
          R[j0 + 1] = xt[1];
        }
      }
#undef QSP_TJ_LOAD
      if (in_tmem) QSP_TJ_WAIT_ST();
      QSP_TJ_FENCE_B();
      __syncthreads();
      QSP_TJ_FENCE_A();
    }
    ++sweeps;
    if (*flag == 0) {  // (ordered by the barrier above)
      conv = true;
      break;
    }
    __syncthreads();
  }
  __syncthreads();
  // rows out (a sweep has an even number of rounds: every team is back on its own slot) + squared norms
  double n_odd = 0.0, n_even = 0.0;
  {
    cplx* wo = Wg + (long)(2 * p + 1) * LDW;
    cplx* we = Wg + (long)(2 * p) * LDW;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
      wo[tl + j * TPP] = R[j];
      n_odd = fma(R[j].x, R[j].x, fma(R[j].y, R[j].y, n_odd));
    }
    if (jj == 0) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const cplx v = S[(p >> 2) * LDW + tl + j * TPP];
        we[tl + j * TPP] = v;
        n_even = fma(v.x, v.x, fma(v.y, v.y, n_even));
      }
    } else {
      const uint32_t ta = tq + (uint32_t)((jj - 1) * 64);
#pragma unroll
      for (int j = 0; j < CPL; j += 2) {
        cplx a, b;
        tmem_ld2(ta + 4 * j, a, b);
        tmem_wait_ld();
        we[tl + j * TPP] = a;
        we[tl + (j + 1) * TPP] = b;
        n_even = fma(a.x, a.x, fma(a.y, a.y, n_even));
        n_even = fma(b.x, b.x, fma(b.y, b.y, n_even));
      }
    }
  }
  if (io->norms_off >= 0) {
    double* nrm = (double*)(QSP_SMEM_BASE() + io->norms_off);
#pragma unroll
    for (int o = TPP >> 1; o > 0; o >>= 1) {
      n_odd += shfl_xor_d(n_odd, o);
      n_even += shfl_xor_d(n_even, o);
    }
    if (tl == 0) {
      nrm[2 * p + 1] = n_odd;
      nrm[2 * p] = n_even;
    }
    io->norms_done = true;
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  if (sweeps_out) *sweeps_out = sweeps;
  return conv;
}
#endif  // QSP_DEV_BUILD

// Can the register-resident path take an (nrows x ldw) problem?  (ldw already padded)
__host__ __device__ inline bool jacobi_oe_fits(int nt, int nrows, int ldw, int cap_elems) {
  int tpp, np;
  if (!jacobi_oe_geometry(nt, nrows, ldw, &tpp, &np)) return false;
  const long ne = (nrows + 1) >> 1;
  return ne * ldw + (ne + 1) / 2 <= cap_elems;  // even rows + their scales
}

// Dispatcher.  Wg: nrows x ldw (row stride ldw), first ndotp columns enter the Gram entries;
// for the register path ldw and ndotp are multiples of QSP_JPAD (zero padded).
__device__ inline bool jacobi_svd_rows(cplx* Wg, int nrows, int ndotp, int ldw, cplx* Wsm, int cap_elems, int tpp_hint,
                                       double tol2, int max_sweeps, double* red, int* sweeps_out, JacobiIO* io = nullptr) {
  if (nrows <= 1) {
    if (sweeps_out) *sweeps_out = 0;
    return true;
  }
  const int nt = blockDim.x;
  int tpp, np;
  if (jacobi_oe_geometry(nt, nrows, ldw, &tpp, &np) && jacobi_oe_fits(nt, nrows, ldw, cap_elems)) {
    const int jlen = ldw / tpp, jdot = (ndotp + tpp - 1) / tpp;
    const int s_off = (int)((char*)Wsm - QSP_SMEM_BASE()), r_off = (int)((char*)red - QSP_SMEM_BASE());
#define QSP_OE(C, N, H) \
  jacobi_rows_oe<C, N, H>(Wg, nrows, jdot, jlen, s_off, tpp, tol2, max_sweeps, r_off, sweeps_out, io)
#ifndef QSP_MAIN_FAST
#define QSP_MAIN_FAST 0  // scaled rotations in the 16-columns-per-lane variant (register limited: off)
#endif
#define QSP_OEF(C, H, T) \
  jacobi_rows_oe<C, 1, H, T, true, ((C) == 16 ? QSP_MAIN_FAST : -1)>(Wg, nrows, jdot, jlen, s_off, tpp, tol2, max_sweeps, \
                                                                       r_off, sweeps_out, io)
    if (np == 1 && jdot == jlen) {  // power-of-two problems of the layer loop and of the input pipeline
      if (tpp == 8) {
#ifdef QSP_DEV_BUILD
        if (jlen == 16 && nrows == 128 && nt == 512 && io && io->tmem_ok)  // experiment: rows in tensor memory
          return jacobi_rows_oe_tmem(Wg, s_off, tol2, max_sweeps, r_off, sweeps_out, io);
#endif
        if (jlen == 16) return QSP_OEF(16, false, 8);  // 128 x 128 (chi = 64 bulk)
        if (jlen == 8) return QSP_OEF(8, true, 8);     // 64 x 64 (chi = 32)
        if (jlen == 4) return QSP_OEF(4, true, 8);     // 32 x 32 (chi = 16)
      } else if (tpp == 16) {
        if (jlen == 8) return QSP_OEF(8, true, 16);    // 64 x 128 (prepare_target / second pass at chi = 64)
        if (jlen == 4) return QSP_OEF(4, true, 16);    // 32 x 64
      }
    }
    if (np == 1) {
      if (jlen <= 2) return QSP_OE(2, 1, true);
      if (jlen <= 4) return QSP_OE(4, 1, true);
      if (jlen <= 8) return QSP_OE(8, 1, true);
      return QSP_OE(16, 1, false);
    }
    if (jlen <= 4) return QSP_OE(4, 2, true);
    if (jlen <= 8) return QSP_OE(8, 2, false);
    return false;  // unreachable: geometry limits np = 2 to 8 columns per lane
#undef QSP_OE
#undef QSP_OEF
  }
  int cap_rows = cap_elems / ldw;
  if (cap_rows < 2) cap_rows = 2;
  int t2 = tpp_hint;
  while (t2 > 1 && t2 > ndotp) t2 >>= 1;
  return jacobi_rows(Wg, nrows, ndotp, ldw, Wsm, cap_rows, t2, tol2, max_sweeps, red, sweeps_out);
}
