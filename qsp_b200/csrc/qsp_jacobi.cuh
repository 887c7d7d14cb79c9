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

// One full tournament sweep over `nrows` rows of W (shared memory, row stride ldw).
// Returns the block-wide number of rotations applied.  All threads must call.
__device__ inline int jacobi_sweep_smem(cplx* W, int nrows, int ndot, int len, int ldw, int tpp, double tol2,
                                        double* red) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nteams = nt / tpp, team = tid / tpp, tl = tid % tpp;
  const int np = (nrows + 1) & ~1;  // padded to even
  const int npairs = np / 2, nrounds = np - 1;
  const int passes = (npairs + nteams - 1) / nteams;
  int nrot = 0;
  for (int r = 0; r < nrounds; ++r) {
    for (int ps = 0; ps < passes; ++ps) {
      const int pi = team + ps * nteams;
      int a = 0, b = 0;
      bool valid = pi < npairs;
      if (valid) {
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
        if (tl == 0) ++nrot;
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
  if (nrows <= cap_rows) {
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
    while (sweeps < max_sweeps) {
      int nrot = 0;
      for (int bi = 0; bi < nblk; ++bi) {
        for (int bj = bi + 1; bj < nblk; ++bj) {
          const int r0 = bi * rb, n0 = imin(rb, nrows - r0);
          const int r1 = bj * rb, n1 = imin(rb, nrows - r1);
          copy_rows(Wsm, len, Wg + (long)r0 * len, len, n0, len);
          copy_rows(Wsm + (long)n0 * len, len, Wg + (long)r1 * len, len, n1, len);
          __syncthreads();
          nrot += jacobi_sweep_smem(Wsm, n0 + n1, ndot, len, len, tpp, tol2, red);
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
