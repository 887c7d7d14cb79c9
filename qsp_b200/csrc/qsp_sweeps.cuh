// The two canonisation sweeps of the sequential MPS-encoding path, batched over
// independent MPS instances: one persistent CTA walks one instance along the chain
// (sites are strictly sequential), instances are strided over the grid.
//
//   k_left_canon      quimb MPS.left_canonize()  (first half of `compress`; tsp.py:72,
//                     core.py:17, sequential.py:54,57)  Householder QR of the
//                     (2a x b) site matrix held in shared memory, R carried to the next
//                     site through a fused R*A GEMM.
//   k_right_compress  quimb MPS.right_compress() on a LEFT-canonical input (second half of
//                     `compress`): per bond a one-sided Jacobi SVD of the (b x 2r') site
//                     matrix, rel. cutoff + max_bond truncation, A[i] <- Vh,
//                     A[i-1] <- A[i-1] (U s).  (T1 is an isometry after the left sweep, so
//                     the reference's QR(T1)/LQ(T2) reduction is the identity on T1 and
//                     the SVD of T2 alone gives the same factors -- DESIGN.md, I1.)
//
// MPS layout in HBM: [B][L][D][2][D] complex128, element (l,p,r) of a site at
// (l*2+p)*D + r; the live extent of every bond is in bonds[B][L+1] (int32).  Nothing
// outside the live extent is ever read.
#pragma once
#include "qsp_common.cuh"
#include "qsp_gemm.cuh"
#include "qsp_jacobi.cuh"

// ----------------------------------------------------------------------------------
// left canonisation
// ----------------------------------------------------------------------------------
__host__ __device__ inline size_t lcanon_smem_bytes(int Dm, int nt) {
  size_t s = 0;
  s += (size_t)Dm * (2 * Dm + 1) * sizeof(cplx);  // S
  s += (size_t)Dm * Dm * sizeof(cplx);            // R
  s += (size_t)Dm * sizeof(cplx);                 // tau
  s += (size_t)Dm * sizeof(double);               // beta
  s += 64 * sizeof(double);                       // red
  s += qsp_gemm_smem_bytes(nt);
  return s + 64;
}

// Householder QR of S (m x n, column-major, column stride ldS) in place.  On return the
// leading K = min(m,n) columns of S hold the thin Q and R (K x n, row-major, stride ldR)
// holds the triangular factor with a non-negative real diagonal.
//
// A warp owns a column while it is updated: the reflector (rows j..m-1 of column j, v_j = 1)
// and the column both sit in registers (QSP_QR_MAXR rows per lane, m - j <= 32*QSP_QR_MAXR),
// so a column is read and written once per reflector; the warp that finishes column j+1
// immediately builds reflector j+1 (look-ahead), so there is one barrier per reflector.
#define QSP_QR_MAXR 8

// reflector from column `col` (rows j..m-1): writes the scaled v below the diagonal, tau, beta.
__device__ __forceinline__ void qr_make_reflector(cplx* col, int j, int m, int lane, cplx* tau, double* beta) {
  double xn2 = 0.0;
  for (int r = j + 1 + lane; r < m; r += 32) xn2 += cabs2(col[r]);
  for (int o = 16; o > 0; o >>= 1) xn2 += shfl_xor_d(xn2, o);
  const cplx alpha = col[j];
  cplx t;
  double bt;
  if (xn2 == 0.0 && alpha.y == 0.0) {
    t = cmake(0.0, 0.0);
    bt = alpha.x;
  } else {
    // reciprocal square roots instead of sqrt + two divisions: this scalar chain sits on the critical path of the
    // factorisation (warp 0's look-ahead), and a dependent FP64 operation costs ~13 cycles here
    const double q = cabs2(alpha) + xn2;
    const double rq = qsp_rsqrt(q);
    const double nrm = q * rq;
    bt = alpha.x >= 0.0 ? -nrm : nrm;
    const double ibt = alpha.x >= 0.0 ? -rq : rq;
    t = cmake((bt - alpha.x) * ibt, -alpha.y * ibt);
    const double dx = alpha.x - bt;                   // |alpha - beta|^2 = dx^2 + alpha.y^2, no cancellation (signs agree)
    const double rd = qsp_rsqrt(fma(dx, dx, alpha.y * alpha.y));
    const double id2 = rd * rd;
    const cplx sc = cmake(dx * id2, -alpha.y * id2);  // 1 / (alpha - beta)
    for (int r = j + 1 + lane; r < m; r += 32) col[r] = cmul(col[r], sc);
  }
  if (lane == 0) {
    tau[j] = t;
    beta[j] = bt;
  }
  __syncwarp();
}

// col <- (I - f v v^H) col on rows j..m-1, v held in registers (vr), v_j = 1.
template <int MAXR>
__device__ __forceinline__ void qr_apply_reflector(const cplx (&vr)[MAXR], cplx* colc, int j, int m, int lane, cplx f) {
  cplx cr[MAXR];
  cplx w = cmake(0.0, 0.0);
#pragma unroll
  for (int k = 0; k < MAXR; ++k) {
    const int r = j + lane + 32 * k;
    if (r < m) {
      cr[k] = colc[r];
      cfma_conja(w, vr[k], cr[k]);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    w.x += shfl_xor_d(w.x, o);
    w.y += shfl_xor_d(w.y, o);
  }
  const cplx g = cmul(f, w);
#pragma unroll
  for (int k = 0; k < MAXR; ++k) {
    const int r = j + lane + 32 * k;
    if (r < m) colc[r] = csub(cr[k], cmul(g, vr[k]));
  }
}

// SHARED: S, R, tau, beta live in the dynamic shared window; they are rebuilt from their offsets inside the
// (out-of-line) function so that the accesses compile to LDS/STS instead of generic loads.
// Two columns at once: the loads, the two dot products and their shuffle reductions are independent,
// so their latencies overlap (a warp has nothing else to run between the barriers of the factorisation).
template <int MAXR>
__device__ __forceinline__ void qr_apply_reflector2(const cplx (&vr)[MAXR], cplx* c0, cplx* c1, int j, int m, int lane, cplx f) {
  cplx a[MAXR], b[MAXR];
  cplx wa = cmake(0.0, 0.0), wb = cmake(0.0, 0.0);
#pragma unroll
  for (int k = 0; k < MAXR; ++k) {
    const int r = j + lane + 32 * k;
    if (r < m) {
      a[k] = c0[r];
      b[k] = c1[r];
    }
  }
#pragma unroll
  for (int k = 0; k < MAXR; ++k) {
    const int r = j + lane + 32 * k;
    if (r < m) {
      cfma_conja(wa, vr[k], a[k]);
      cfma_conja(wb, vr[k], b[k]);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    wa.x += shfl_xor_d(wa.x, o);
    wa.y += shfl_xor_d(wa.y, o);
    wb.x += shfl_xor_d(wb.x, o);
    wb.y += shfl_xor_d(wb.y, o);
  }
  const cplx ga = cmul(f, wa), gb = cmul(f, wb);
#pragma unroll
  for (int k = 0; k < MAXR; ++k) {
    const int r = j + lane + 32 * k;
    if (r < m) {
      c0[r] = csub(a[k], cmul(ga, vr[k]));
      c1[r] = csub(b[k], cmul(gb, vr[k]));
    }
  }
}

// columns first, first + step, ... < end of S get H applied; in pairs while the registers allow (measured:
// pays up to 64 rows, costs 10 % at 128 rows where two columns and the reflector no longer fit)
template <int MAXR>
__device__ __forceinline__ void qr_apply_columns(const cplx (&vr)[MAXR], cplx* S, int ldS, int first, int step, int end, int j,
                                                 int m, int lane, cplx f) {
  int c = first;
  if (MAXR <= 2) {
    for (; c + step < end; c += 2 * step)
      qr_apply_reflector2(vr, S + (long)c * ldS, S + (long)(c + step) * ldS, j, m, lane, f);
  }
  for (; c < end; c += step) qr_apply_reflector(vr, S + (long)c * ldS, j, m, lane, f);
}

template <int MAXR, bool SHARED>
__device__ __noinline__ void cta_householder_qr_t(cplx* S, int m, int n, int ldS, cplx* R, int ldR, cplx* tau, double* beta,
                                                  long long* pc);

__device__ __forceinline__ void cta_householder_qr(cplx* S, int m, int n, int ldS, cplx* R, int ldR, cplx* tau, double* beta,
                                                   long long* pc) {
  if (m <= 32)
    cta_householder_qr_t<1, true>(S, m, n, ldS, R, ldR, tau, beta, pc);
  else if (m <= 64)
    cta_householder_qr_t<2, true>(S, m, n, ldS, R, ldR, tau, beta, pc);
  else if (m <= 128)
    cta_householder_qr_t<4, true>(S, m, n, ldS, R, ldR, tau, beta, pc);
  else
    cta_householder_qr_t<QSP_QR_MAXR, true>(S, m, n, ldS, R, ldR, tau, beta, pc);
}

template <int MAXR, bool SHARED>
__device__ __noinline__ void cta_householder_qr_t(cplx* S, int m, int n, int ldS, cplx* R, int ldR, cplx* tau, double* beta,
                                                  long long* pc) {
  QSP_ASSUME_SHARED(SHARED, S);
  QSP_ASSUME_SHARED(SHARED, R);
  QSP_ASSUME_SHARED(SHARED, tau);
  QSP_ASSUME_SHARED(SHARED, beta);
  long long tck = qsp_clock();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int K = imin(m, n);
  if (warp == 0) qr_make_reflector(S, 0, m, lane, tau, beta);
  __syncthreads();
  int jstart = 0;
  if (MAXR <= 4 && nw > 1 && K >= 2) {
    // Two reflectors per pass (registers allow it up to 128 rows): with z = v_{j+1}^H v_j,
    //   H_{j+1}^H H_j^H a = a - f0 w0 v_j - f1 (u1 - f0 z w0) v_{j+1},  w0 = v_j^H a,  u1 = v_{j+1}^H a,  f = conj(tau).
    // Warp 0 looks two reflectors ahead (columns j+2, j+3), the other warps share the columns from j+4 on;
    // one barrier per pair.
    if (warp == 0) {  // reflector 1: H_0^H on column 1, then build it
      cplx v0[MAXR];
      const cplx t0 = tau[0];
#pragma unroll
      for (int k = 0; k < MAXR; ++k) {
        const int r = lane + 32 * k;
        v0[k] = (r == 0) ? cmake(1.0, 0.0) : (r < m ? S[r] : cmake(0.0, 0.0));
      }
      qr_apply_reflector(v0, S + ldS, 0, m, lane, cconj(t0));
      __syncwarp();
      qr_make_reflector(S + ldS, 1, m, lane, tau, beta);
    }
    __syncthreads();
    int j = 0;
    for (; j + 1 < K; j += 2) {
      const cplx* colj = S + (long)j * ldS;
      const cplx* colq = S + (long)(j + 1) * ldS;
      const cplx f0 = cconj(tau[j]), f1 = cconj(tau[j + 1]);
      cplx vj[MAXR], vq[MAXR];
      cplx z = cmake(0.0, 0.0);
#pragma unroll
      for (int k = 0; k < MAXR; ++k) {
        const int r = j + lane + 32 * k;
        vj[k] = (r == j) ? cmake(1.0, 0.0) : (r < m ? colj[r] : cmake(0.0, 0.0));
        vq[k] = (r == j + 1) ? cmake(1.0, 0.0) : ((r > j + 1 && r < m) ? colq[r] : cmake(0.0, 0.0));
        cfma_conja(z, vq[k], vj[k]);
      }
      for (int o = 16; o > 0; o >>= 1) {
        z.x += shfl_xor_d(z.x, o);
        z.y += shfl_xor_d(z.y, o);
      }
      const cplx f0z = cmul(f0, z);
      auto pair_apply = [&](cplx* col) {
        cplx a[MAXR];
        cplx w0 = cmake(0.0, 0.0), u1 = cmake(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < MAXR; ++k) {
          const int r = j + lane + 32 * k;
          a[k] = (r < m) ? col[r] : cmake(0.0, 0.0);
          cfma_conja(w0, vj[k], a[k]);
          cfma_conja(u1, vq[k], a[k]);
        }
        for (int o = 16; o > 0; o >>= 1) {
          w0.x += shfl_xor_d(w0.x, o);
          w0.y += shfl_xor_d(w0.y, o);
          u1.x += shfl_xor_d(u1.x, o);
          u1.y += shfl_xor_d(u1.y, o);
        }
        const cplx g0 = cmul(f0, w0);
        const cplx g1 = cmul(f1, csub(u1, cmul(f0z, w0)));
#pragma unroll
        for (int k = 0; k < MAXR; ++k) {
          const int r = j + lane + 32 * k;
          if (r < m) col[r] = csub(csub(a[k], cmul(g0, vj[k])), cmul(g1, vq[k]));
        }
      };
      if (warp == 0) {
        if (j + 2 < n) {
          cplx* c2 = S + (long)(j + 2) * ldS;
          pair_apply(c2);
          if (j + 2 < K) {
            __syncwarp();
            qr_make_reflector(c2, j + 2, m, lane, tau, beta);
          }
        }
        if (j + 3 < n) {
          cplx* c3 = S + (long)(j + 3) * ldS;
          pair_apply(c3);
          if (j + 3 < K) {  // H_{j+2}^H as well, then reflector j+3 (otherwise the tail below applies H_{j+2})
            __syncwarp();
            const cplx* c2 = S + (long)(j + 2) * ldS;
            cplx v2[MAXR];
#pragma unroll
            for (int k = 0; k < MAXR; ++k) {
              const int r = j + 2 + lane + 32 * k;
              v2[k] = (r == j + 2) ? cmake(1.0, 0.0) : (r < m ? c2[r] : cmake(0.0, 0.0));
            }
            qr_apply_reflector(v2, c3, j + 2, m, lane, cconj(tau[j + 2]));
            __syncwarp();
            qr_make_reflector(c3, j + 3, m, lane, tau, beta);
          }
        }
      } else {
        for (int c = j + 4 + (warp - 1); c < n; c += nw - 1) pair_apply(S + (long)c * ldS);
      }
      __syncthreads();
    }
    jstart = j;  // K odd: reflector K-1 exists but has not been applied to the columns behind it
    if (jstart < K) {
      const cplx* colj = S + (long)jstart * ldS;
      const cplx tj = tau[jstart];
      cplx vr[MAXR];
#pragma unroll
      for (int k = 0; k < MAXR; ++k) {
        const int r = jstart + lane + 32 * k;
        vr[k] = (r == jstart) ? cmake(1.0, 0.0) : (r < m ? colj[r] : cmake(0.0, 0.0));
      }
      if (tj.x != 0.0 || tj.y != 0.0) qr_apply_columns(vr, S, ldS, jstart + 1 + warp, nw, n, jstart, m, lane, cconj(tj));
      __syncthreads();
    }
    jstart = K;
  }
  for (int j = jstart; j < K; ++j) {
    const cplx* colj = S + (long)j * ldS;
    const cplx tj = tau[j];
    cplx vr[MAXR];
#pragma unroll
    for (int k = 0; k < MAXR; ++k) {
      const int r = j + lane + 32 * k;
      vr[k] = (r == j) ? cmake(1.0, 0.0) : (r < m ? colj[r] : cmake(0.0, 0.0));
    }
    // apply H_j^H = I - conj(tau) v v^H to the trailing columns.  Warp 0 takes column j+1 and
    // then builds reflector j+1 (look-ahead, the critical path); the other warps share the rest.
    const bool nz = (tj.x != 0.0 || tj.y != 0.0);
    if (nw == 1) {
      for (int c = j + 1; c < n; ++c) {
        if (nz) qr_apply_reflector(vr, S + (long)c * ldS, j, m, lane, cconj(tj));
        if (c == j + 1 && j + 1 < K) {
          __syncwarp();
          qr_make_reflector(S + (long)c * ldS, j + 1, m, lane, tau, beta);
        }
      }
    } else if (warp == 0) {
      if (j + 1 < n) {
        cplx* colc = S + (long)(j + 1) * ldS;
        if (nz) qr_apply_reflector(vr, colc, j, m, lane, cconj(tj));
        if (j + 1 < K) {
          __syncwarp();
          qr_make_reflector(colc, j + 1, m, lane, tau, beta);
        }
      }
    } else if (nz) {
      qr_apply_columns(vr, S, ldS, j + 2 + (warp - 1), nw - 1, n, j, m, lane, cconj(tj));
    }
    __syncthreads();
  }
  pc[1] += qsp_clock() - tck, tck = qsp_clock();
  // ---- R out (upper trapezoid, diagonal = beta), sign-fixed --------------------------
  for (int idx = tid; idx < K * n; idx += blockDim.x) {
    const int r = idx / n, c = idx - r * n;
    cplx v = cmake(0.0, 0.0);
    if (c == r)
      v = cmake(beta[r], 0.0);
    else if (c > r)
      v = S[(long)c * ldS + r];
    if (beta[r] < 0.0) v = cmake(-v.x, -v.y);
    R[(long)r * ldR + c] = v;
  }
  __syncthreads();
  pc[2] += qsp_clock() - tck, tck = qsp_clock();
  // ---- thin Q in place (zung2r) ---------------------------------------------------------
  int jtop = K - 1;
  if (MAXR <= 4) {
    // Two reflectors per pass while the registers allow: H_{j-1} H_j a = a - tau_j w1 v_j - tau_{j-1} (w0 - tau_j z w1) v_{j-1}
    // with w1 = v_j^H a, w0 = v_{j-1}^H a, z = v_{j-1}^H v_j -- the two dot products of a column are independent,
    // a column is read and written once per pair, and a pair costs two barriers instead of four.
    for (; jtop >= 1; jtop -= 2) {
      const int j = jtop, b0 = j - 1;  // rows b0 .. m-1 are touched
      cplx* colj = S + (long)j * ldS;
      cplx* colp = S + (long)(j - 1) * ldS;
      const cplx tj = tau[j], tp = tau[j - 1];
      cplx vj[MAXR], vp[MAXR];
      cplx z = cmake(0.0, 0.0);
#pragma unroll
      for (int k = 0; k < MAXR; ++k) {
        const int r = b0 + lane + 32 * k;
        vj[k] = (r == j) ? cmake(1.0, 0.0) : ((r > j && r < m) ? colj[r] : cmake(0.0, 0.0));
        vp[k] = (r == b0) ? cmake(1.0, 0.0) : ((r > b0 && r < m) ? colp[r] : cmake(0.0, 0.0));
        cfma_conja(z, vp[k], vj[k]);
      }
      for (int o = 16; o > 0; o >>= 1) {
        z.x += shfl_xor_d(z.x, o);
        z.y += shfl_xor_d(z.y, o);
      }
      const cplx tz = cmul(tj, z);
      __syncthreads();  // every warp holds both reflectors: their columns may now be overwritten
      for (int c = j + 1 + warp; c < K; c += nw) {
        cplx* col = S + (long)c * ldS;
        cplx a[MAXR];
        cplx w1 = cmake(0.0, 0.0), w0 = cmake(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < MAXR; ++k) {
          const int r = b0 + lane + 32 * k;
          a[k] = (r < m) ? col[r] : cmake(0.0, 0.0);
          cfma_conja(w1, vj[k], a[k]);
          cfma_conja(w0, vp[k], a[k]);
        }
        for (int o = 16; o > 0; o >>= 1) {
          w1.x += shfl_xor_d(w1.x, o);
          w1.y += shfl_xor_d(w1.y, o);
          w0.x += shfl_xor_d(w0.x, o);
          w0.y += shfl_xor_d(w0.y, o);
        }
        const cplx g1 = cmul(tj, w1);
        const cplx g0 = cmul(tp, csub(w0, cmul(tz, w1)));
#pragma unroll
        for (int k = 0; k < MAXR; ++k) {
          const int r = b0 + lane + 32 * k;
          if (r < m) col[r] = csub(csub(a[k], cmul(g1, vj[k])), cmul(g0, vp[k]));
        }
      }
      if (warp == nw - 1) {
        // column j = H_{j-1} H_j e_j (w1 = 1, w0 = conj(v_{j-1}[j])), column j-1 = H_{j-1} e_{j-1}; sign-fixed like R
        const double sj = beta[j] < 0.0 ? -1.0 : 1.0, sp = beta[j - 1] < 0.0 ? -1.0 : 1.0;
        cplx vpj = cmake(0.0, 0.0);  // v_{j-1}[j]: lane 1, k = 0 holds it
        {
          const cplx mine = vp[0];
          vpj.x = __shfl_sync(QSP_FULL_MASK, mine.x, 1);
          vpj.y = __shfl_sync(QSP_FULL_MASK, mine.y, 1);
        }
        const cplx g0 = cmul(tp, csub(cconj(vpj), tz));
#pragma unroll
        for (int k = 0; k < MAXR; ++k) {
          const int r = b0 + lane + 32 * k;
          if (r < m) {
            cplx ej = (r == j) ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
            cplx ep = (r == b0) ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
            colj[r] = cscale(csub(csub(ej, cmul(tj, vj[k])), cmul(g0, vp[k])), sj);
            colp[r] = cscale(csub(ep, cmul(tp, vp[k])), sp);
          }
        }
        for (int r = lane; r < b0; r += 32) {
          colj[r] = cmake(0.0, 0.0);
          colp[r] = cmake(0.0, 0.0);
        }
      }
      __syncthreads();
    }
  }
  for (int j = jtop; j >= 0; --j) {
    cplx* colj = S + (long)j * ldS;
    const cplx tj = tau[j];
    if (tj.x != 0.0 || tj.y != 0.0) {
      cplx vr[MAXR];
#pragma unroll
      for (int k = 0; k < MAXR; ++k) {
        const int r = j + lane + 32 * k;
        vr[k] = (r == j) ? cmake(1.0, 0.0) : (r < m ? colj[r] : cmake(0.0, 0.0));
      }
      qr_apply_columns(vr, S, ldS, j + 1 + warp, nw, K, j, m, lane, tj);
    }
    __syncthreads();
    const double sg = beta[j] < 0.0 ? -1.0 : 1.0;
    for (int r = tid; r < m; r += blockDim.x) {
      cplx q;
      if (r < j)
        q = cmake(0.0, 0.0);
      else if (r == j)
        q = cmake(1.0 - tj.x, -tj.y);
      else
        q = cmul(cmake(-tj.x, -tj.y), colj[r]);
      colj[r] = cscale(q, sg);
    }
    __syncthreads();
  }
  pc[3] += qsp_clock() - tck;
}

// S and R live in shared memory (lcanon_smem_bytes); bonds beyond that take big_left_canonize (qsp_big_sweeps.cuh).
__device__ __forceinline__ void left_canon_impl(cplx* __restrict__ mps, int32_t* __restrict__ bonds, int B, int L, int D,
                                                int Dm, int32_t* __restrict__ status, long long* prof) {
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  char* sm = QSP_SMEM_BASE();
  const int ldS = 2 * Dm + 1;
  cplx* S = (cplx*)sm;
  cplx* R = S + (size_t)Dm * ldS;
  cplx* tau = R + (size_t)Dm * Dm;
  double* beta = (double*)(tau + Dm);
  double* red = beta + Dm;
  cplx* tiles = (cplx*)(((uintptr_t)(red + 64) + 15) & ~(uintptr_t)15);
  const int tid = threadIdx.x, nt = blockDim.x;
  const long site_stride = (long)D * 2 * D;

  for (int inst = blockIdx.x; inst < B; inst += gridDim.x) {
    cplx* A = mps + (long)inst * L * site_stride;
    int32_t* bd = bonds + (long)inst * (L + 1);
    bool bad = false;
    for (int i = 0; i <= L; ++i)
      if (bd[i] > Dm || bd[i] < 1) bad = true;
    if (bad) {
      if (tid == 0 && status) status[inst] |= QSP_ST_BOND_OVERFLOW;
      continue;
    }
    int a = bd[0];      // current (new) left bond
    int a_old = bd[0];  // left extent of the stored tensor = column count of the carried R
    for (int i = 0; i < L; ++i) {
      cplx* Ai = A + (long)i * site_stride;
      const int b = bd[i + 1];
      const int m = 2 * a, n = b;
      long long tck = qsp_clock();
      if (i == 0) {
        for (int idx = tid; idx < m * n; idx += nt) {
          const int r = idx / n, c = idx - r * n;
          S[(long)c * ldS + r] = Ai[(long)r * D + c];
        }
        __syncthreads();
      } else {
        // S[(k,p), r] = sum_x R[k,x] A[x,p,r]
        for (int p = 0; p < 2; ++p)
          cta_zgemm(a, n, a_old, R, Dm, 1, false, Ai + (long)p * D, 2L * D, 1, false, S + p, 2, ldS, false, tiles);
      }
      if (i == L - 1) {
        __syncthreads();
        for (int idx = tid; idx < m * n; idx += nt) {
          const int r = idx / n, c = idx - r * n;
          Ai[(long)r * D + c] = S[(long)c * ldS + r];
        }
        __syncthreads();
        break;
      }
      __syncthreads();
      pc[0] += qsp_clock() - tck;
      cta_householder_qr(S, m, n, ldS, R, Dm, tau, beta, pc);
      tck = qsp_clock();
      pc[6] += 1;
      const int K = imin(m, n);
      for (int idx = tid; idx < m * K; idx += nt) {
        const int r = idx / K, c = idx - r * K;
        Ai[(long)r * D + c] = S[(long)c * ldS + r];
      }
      __syncthreads();
      if (tid == 0) {
        bd[i] = a;
        bd[i + 1] = K;
      }
      a = K;
      a_old = n;
      __syncthreads();
      pc[4] += qsp_clock() - tck;
    }
  }
  (void)red;
  if (prof && tid == 0)
    for (int q = 0; q < 8; ++q) prof[(long)blockIdx.x * 8 + q] += pc[q];
}

__global__ void __launch_bounds__(512) k_left_canon(cplx* __restrict__ mps, int32_t* __restrict__ bonds, int B, int L, int D,
                                                    int Dm, int32_t* __restrict__ status, long long* prof) {
  left_canon_impl(mps, bonds, B, L, D, Dm, status, prof);
}
// ----------------------------------------------------------------------------------
// right compression (truncating SVD sweep) of a left-canonical MPS
// ----------------------------------------------------------------------------------
struct RCompressArgs {
  const cplx* in;           // [B][L][Din][2][Din]
  const int32_t* bonds_in;  // [B][L+1]
  int Din;
  cplx* out;                // [B][L][Dout][2][Dout]
  int32_t* bonds_out;       // [B][L+1]
  int Dout;
  int DinL, DoutL;          // largest LIVE bonds the launch was planned for (<= Din, Dout): scratch / shared sizing
  int B, L;
  int max_bond;             // 0 = unlimited (up to DoutL)
  double cutoff;            // relative (quimb cutoff_mode='rel'), no renormalisation
  int normalize;            // divide site 0 by its norm at the end
  double* svals;            // [B][L-1][nsv] largest singular values per bond (bond i between i-1,i at row i-1), or null
  int nsv;
  cplx* amp0;               // [B] <0...0|out> / ||out||  (null to skip)
  int32_t* status;          // [B]
  cplx* work;               // per-CTA scratch, rcompress_work_elems() each
  long work_stride;         // elements per CTA
  int* queue;               // instance queue head (zeroed before the launch): CTAs pull instances dynamically
  // Fused gate-apply-and-split (apply_inverse_unitary_layer_on_wfn + compress in one sweep): when G
  // is set, `in` is the LEFT-canonical chi-MPS psi itself (bonds_in its bonds, Din its padding) and
  // the sweep runs on U^dagger psi without materialising its doubled tensors:
  // site t of U^dagger psi is G_{t-1}^dagger o A_{t-1}, so M = A_{t-1} C needs half the flops.
  const cplx* G;            // [B][L][4][4] or null
  const int32_t* kind;      // [B][L]
  int cap_rows_elems;       // shared elements available for Jacobi rows
  int tpp;                  // lanes per Jacobi pair
  int max_sweeps;
  long long* prof;          // optional [grid][8] cycle counters (qsp_set_profile_buffer), else null
  int presort;              // Jacobi rows ordered by norm before every sweep (always on; the development build can switch it off)
};

// scratch (complex elements per CTA): Mg (Din x 2Dout), Wg (2Dout x (Din+2Dout)),
// Vg (Dout x 2Dout), Cg (Din x Dout), f (2 x Dout)
__host__ __device__ inline long rcompress_work_elems(int Din, int Dout) {
  return (long)Din * 2 * Dout + (long)2 * Dout * (Din + 2 * Dout + 64) + (long)Dout * 2 * Dout + (long)Din * Dout +
         2L * Dout + 64;
}
// shared: Jacobi rows | gemm tiles (aliased onto the same area) + norms/perm + red
__host__ __device__ inline size_t rcompress_smem_bytes(int jac_elems, int Dout, int nt) {
  size_t a = (size_t)jac_elems * sizeof(cplx);
  size_t g = qsp_gemm_smem_bytes(nt, QSP_GEMM_TK_DEEP);
  size_t s = a > g ? a : g;
  s += (size_t)(2 * Dout) * (sizeof(double) + sizeof(int));  // norms + perm
  s += 64 * sizeof(double);
  return s + 64;
}

__global__ void __launch_bounds__(512) k_right_compress(RCompressArgs P) {
  char* sm = QSP_SMEM_BASE();
  const int tid = threadIdx.x, nt = blockDim.x;
  const int Din = P.Din, Dout = P.Dout, L = P.L;  // padded strides of the in / out buffers
  const int DinL = P.DinL, DoutL = P.DoutL;       // live bounds: scratch and shared-memory sizing
  size_t jac_bytes = (size_t)P.cap_rows_elems * sizeof(cplx);
  size_t gem_bytes = qsp_gemm_smem_bytes(nt, QSP_GEMM_TK_DEEP);
  size_t big = jac_bytes > gem_bytes ? jac_bytes : gem_bytes;
  cplx* Wsm = (cplx*)sm;
  cplx* tiles = (cplx*)sm;  // aliased: GEMMs and Jacobi never overlap in time
  double* norms = (double*)(sm + big);
  int* perm = (int*)(norms + 2 * DoutL);
  double* red = (double*)(perm + 2 * DoutL);
  red = (double*)(((uintptr_t)red + 7) & ~(uintptr_t)7);

  cplx* Mg = P.work + (long)blockIdx.x * P.work_stride;
  cplx* Wg = Mg + (long)DinL * 2 * DoutL;
  cplx* Vg = Wg + (long)2 * DoutL * (DinL + 2 * DoutL + 64);
  cplx* Cg = Vg + (long)DoutL * 2 * DoutL;
  cplx* fg = Cg + (long)DinL * DoutL;  // two ping-pong vectors for the <0..0| amplitude

  const long sin = (long)Din * 2 * Din, sout = (long)Dout * 2 * Dout;
  const int kmax = (P.max_bond > 0) ? imin(P.max_bond, DoutL) : DoutL;
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tck = 0;
#define QSP_TICK() (tck = qsp_clock())
#define QSP_TOCK(slot) (pc[slot] += qsp_clock() - tck, tck = qsp_clock())

  // dynamic instance queue: Jacobi sweep counts differ per instance, so CTAs pull work instead of striding
  int* s_inst = (int*)(red + 56);
  int* s_bad = (int*)(red + 57);  // a thread saw a non-finite singular value
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      *s_inst = atomicAdd(P.queue, 1);
      *s_bad = 0;
    }
    __syncthreads();
    const int inst = *s_inst;
    if (inst >= P.B) break;
    const cplx* A = P.in + (long)inst * L * sin;
    const int32_t* bi = P.bonds_in + (long)inst * (L + 1);
    cplx* O = P.out + (long)inst * L * sout;
    int32_t* bo = P.bonds_out + (long)inst * (L + 1);
    int st = 0;
    bool bad = false;
    const bool fused = P.G != nullptr;
    const cplx* Gi = fused ? P.G + (long)inst * L * 16 : nullptr;
    const int32_t* kd = fused ? P.kind + (long)inst * L : nullptr;
    for (int i = 0; i <= L; ++i) {
      // live bond of the swept state: psi itself, or U^dagger psi (doubled inside a segment)
      int bb = bi[i];
      if (fused && i > 0 && i < L && kd[i - 1] == 2) bb = 2 * bi[i - 1];
      if (bb > DinL || bi[i] < 1 || (fused && bi[i] > Din)) bad = true;
    }
    if (bi[L] > DoutL || bi[0] > DoutL) bad = true;
    if (bad) {
      if (tid == 0 && P.status) P.status[inst] |= QSP_ST_BOND_OVERFLOW;
      continue;
    }
    int rp = bi[L];  // r' : current (already truncated) right bond of site i
    int fcur = 0;
    if (tid == 0) {
      bo[L] = bi[L];
      bo[0] = bi[0];
    }
    if (P.amp0)
      for (int k = tid; k < rp; k += nt) fg[k] = cmake(k == 0 ? 1.0 : 0.0, 0.0);
    __syncthreads();

    for (int i = L - 1; i >= 0; --i) {
      const cplx* Ai = A + (long)i * sin;
      cplx* Oi = O + (long)i * sout;
      // extents of site i of the swept state (left bond b, right bond ro)
      const bool start = fused && (i == 0 || kd[i - 1] == 1);
      const int b = (!fused || start) ? bi[i] : 2 * bi[i - 1];
      const int ro = (!fused || kd[i] == 1) ? bi[i + 1] : 2 * bi[i];
      const int c = 2 * rp;
      QSP_TICK();
      if (!fused) {
        // ---- M[(x,p),k] = sum_r A[x,p,r] C[r,k]   (C = identity at the last site) -----------
        if (i == L - 1) {
          for (int idx = tid; idx < 2 * b * rp; idx += nt) {
            const int row = idx / rp, k = idx - row * rp;
            Mg[idx] = Ai[(long)row * Din + k];
          }
          __syncthreads();
        } else {
          cta_zgemm_deep(2 * b, rp, ro, Ai, Din, 1, false, Cg, DoutL, 1, false, Mg, rp, 1, false, tiles);
          pc[7] += 8LL * (2 * b) * (rp) * (ro);
        }
      } else if (start && kd[i] == 2) {
        // first site of a multi-site segment: identity embedding, M[l,(a',k)] = C[(l,a'),k]
        for (int idx = tid; idx < b * c; idx += nt) {
          const int l = idx / c, col = idx - l * c;
          const int ap = col / rp, k = col - ap * rp;
          Mg[idx] = Cg[((long)l * 2 + ap) * DoutL + k];
        }
        __syncthreads();
      } else {
        // T[(x,p),k] = sum_r A_i[x,p,r] C[r,k] is needed for a one-qubit gate on site i
        // (single-site segment, or the end of a segment); Tg aliases the Jacobi staging area.
        cplx* Tg = Wg;
        const cplx* Cuse = Cg;  // carried U s with rows (x,b) [bulk] or r [after a one-qubit gate]
        long ldC = DoutL;
        if (kd[i] == 1) {
          const int xe = bi[i], re = bi[i + 1];
          if (i == L - 1) {
            for (int idx = tid; idx < 2 * xe * rp; idx += nt) {
              const int row = idx / rp, k = idx - row * rp;
              Tg[idx] = Ai[(long)row * Din + k];
            }
            __syncthreads();
          } else {
            cta_zgemm_deep(2 * xe, rp, re, Ai, Din, 1, false, Cg, DoutL, 1, false, Tg, rp, 1, false, tiles);
          pc[7] += 8LL * (2 * xe) * (rp) * (re);
          }
          // Q[(x,q),k] = sum_p conj(G_i[p][q]) T[(x,p),k]   (G_i^dagger on the physical leg)
          const cplx* ge = Gi + i * 16;
          cplx* Qg = start ? Mg : Vg;  // single-site segment: this already is M
          for (int idx = tid; idx < 2 * xe * rp; idx += nt) {
            const int row = idx / rp, k = idx - row * rp;
            const int x = row >> 1, q = row & 1;
            cplx acc = cmake(0.0, 0.0);
            for (int p = 0; p < 2; ++p) cfma_conja(acc, ge[p * 4 + q], Tg[((long)x * 2 + p) * rp + k]);
            Qg[idx] = acc;
          }
          __syncthreads();
          Cuse = Vg;
          ldC = rp;
        }
        if (!start) {
          // bulk gate G_{i-1} on (i-1, i):  P[(l,a),(b,k)] = sum_x A_{i-1}[l,a,x] C[(x,b),k],
          // M[(l,a'),(b',k)] = sum_{a,b} conj(G_{i-1}[(a b)][(a' b')]) P[(l,a),(b,k)]
          const cplx* Aj = A + (long)(i - 1) * sin;
          const cplx* g = Gi + (i - 1) * 16;
          const int lj = bi[i - 1], xj = bi[i];
          for (int bq = 0; bq < 2; ++bq)
            cta_zgemm_deep(2 * lj, rp, xj, Aj, Din, 1, false, Cuse + (long)bq * ldC, 2 * ldC, 1, false, Tg + (long)bq * rp,
                           2L * rp, 1, false, tiles);
          pc[7] += 16LL * (2 * lj) * rp * xj;
          for (int idx = tid; idx < lj * rp; idx += nt) {
            const int l = idx / rp, k = idx - l * rp;
            cplx pv[4];
#pragma unroll
            for (int ab = 0; ab < 4; ++ab) pv[ab] = Tg[((long)l * 2 + (ab >> 1)) * (2 * rp) + (ab & 1) * rp + k];
#pragma unroll
            for (int apbp = 0; apbp < 4; ++apbp) {
              cplx acc = cmake(0.0, 0.0);
#pragma unroll
              for (int ab = 0; ab < 4; ++ab) cfma_conja(acc, g[ab * 4 + apbp], pv[ab]);
              Mg[((long)l * 2 + (apbp >> 1)) * c + (apbp & 1) * rp + k] = acc;
            }
          }
          __syncthreads();
        }
      }
      if (i == 0) {
        // orthogonality centre: site 0 <- M (b x 2 x rp), optional normalisation
        double n2 = 0.0;
        for (int idx = tid; idx < 2 * b * rp; idx += nt) n2 += cabs2(Mg[idx]);
        n2 = block_sum(n2, red);
        double sc = 1.0;
        if (P.normalize) {
          if (!(n2 > 0.0) || !(n2 < 1e300)) st |= QSP_ST_BADNORM;
          else sc = 1.0 / sqrt(n2);
        }
        for (int idx = tid; idx < 2 * b * rp; idx += nt) {
          const int row = idx / rp, k = idx - row * rp;
          Oi[(long)row * Dout + k] = cscale(Mg[idx], sc);
        }
        if (P.amp0) {
          // amplitude <0..0|psi> = sum_k M[(0,0),k] f[k], divided by the norm of psi
          __syncthreads();
          if (tid == 0) {
            cplx acc = cmake(0.0, 0.0);
            const cplx* f = fg + fcur * DoutL;
            for (int k = 0; k < rp; ++k) cfma(acc, Mg[k], f[k]);
            const double nn = sqrt(n2);
            P.amp0[inst] = (nn > 0.0) ? cscale(acc, 1.0 / nn) : cmake(0.0, 0.0);
          }
        }
        __syncthreads();
        break;
      }
      QSP_TOCK(0);
      // ---- Jacobi SVD of M (b x c) ----------------------------------------------------
      const bool modeC = b > c;  // more rows than columns: rotate the columns, accumulate J
      const int nrows = modeC ? c : b;
      const int ndot = modeC ? b : c;
      // pad the Gram part and the whole row to a multiple of the team width (zero columns)
      const int tq = QSP_JPAD;
      const int ndotp = ((ndot + tq - 1) / tq) * tq;
      int len = modeC ? ndotp + c : ndotp;
      len = ((len + tq - 1) / tq) * tq;
      const bool oe = nrows >= 2 && jacobi_oe_fits(nt, nrows, len, P.cap_rows_elems);  // register-resident Jacobi takes it
      if (!oe && !modeC) len = c;  // generic path: no padding needed
      const int joff = modeC ? ndotp : 0;  // first column of the appended identity
      JacobiIO io;
      io.src = nullptr;
      io.src_ld = io.src_cols = 0;
      io.norms_off = (int)((char*)norms - sm);
      io.norms_done = false;
      io.tmem = 0u;
      io.tmem_ok = false;
      io.presort = P.presort != 0;
      if (oe && !modeC) {
        // the register path loads M itself (zero padded on the fly) and leaves s Vh in Wg: no staging copy
        io.src = Mg;
        io.src_ld = io.src_cols = c;
      } else if (modeC) {
        // Wg = [ M^T 0 | I 0 ]  (c x len)
        for (int idx = tid; idx < c * len; idx += nt) {
          const int r = idx / len, cc = idx - r * len;
          cplx v = cmake(0.0, 0.0);
          if (cc < b)
            v = Mg[(long)cc * c + r];
          else if (cc >= joff && cc - joff == r)
            v = cmake(1.0, 0.0);
          Wg[idx] = v;
        }
      } else {
        for (int idx = tid; idx < b * len; idx += nt) {
          const int r = idx / len, cc = idx - r * len;
          Wg[idx] = (cc < c) ? Mg[(long)r * c + cc] : cmake(0.0, 0.0);
        }
      }
      __syncthreads();
      if (P.cap_rows_elems / len < 2) st |= QSP_ST_BOND_OVERFLOW;
      const double tol = 4e-16 * sqrt((double)ndot);
      QSP_TOCK(1);
      int nsweeps = 0;
      if (!jacobi_svd_rows(Wg, nrows, modeC ? ndotp : len, len, Wsm, P.cap_rows_elems, P.tpp, tol * tol, P.max_sweeps, red,
                           &nsweeps, &io))
        st |= QSP_ST_JACOBI_NOCONV;
      QSP_TOCK(2);
      pc[5] += nsweeps;
      pc[6] += 1;
      // executed flops (bench.py: roofline.executed_frac): a Jacobi visit costs 8 DFMA per Gram column and
      // 4 DMUL + 8 DFMA per rotated column (16 + 20 flops); the GEMMs of the site count 8 m n k each (QSP_GEMM_FLOPS)
      pc[7] += (long long)nsweeps * ((long long)nrows * (nrows - 1) / 2) * (16LL * (modeC ? ndotp : len) + 20LL * len);
      // ---- singular values = row norms; rank them ---------------------------------------
      if (io.norms_done) {  // the register path left the squared norms behind
        for (int r = tid; r < nrows; r += nt) norms[r] = sqrt(norms[r]);
      } else {
        const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
        for (int r = warp; r < nrows; r += nw) {
          double s2 = 0.0;
          for (int cc = lane; cc < ndot; cc += 32) s2 += cabs2(Wg[(long)r * len + cc]);
          for (int o = 16; o > 0; o >>= 1) s2 += shfl_xor_d(s2, o);
          if (lane == 0) norms[r] = sqrt(s2);
        }
      }
      __syncthreads();
      // non-finite input: rank it last (a NaN would break the total order below and leave perm[] slots unwritten)
      for (int r = tid; r < nrows; r += nt)
        if (!(norms[r] <= 1e300)) {
          norms[r] = 0.0;
          *s_bad = 1;
        }
      __syncthreads();
      for (int r = tid; r < nrows; r += nt) {
        const double s = norms[r];
        int rank = 0;
        for (int q = 0; q < nrows; ++q) {
          const double t = norms[q];
          if (t > s || (t == s && q < r)) ++rank;
        }
        perm[rank] = r;
      }
      __syncthreads();
      const double smax = norms[perm[0]];
      int k = 0;
      for (int q = 0; q < nrows; ++q)
        if (norms[perm[q]] > P.cutoff * smax) ++k;  // monotone: counts the leading run
      if (k > kmax) {
        if (P.max_bond <= 0 || P.max_bond > DoutL) st |= QSP_ST_BOND_OVERFLOW;  // cutoff alone did not fit the buffer
        k = kmax;
      }
      k = imax(k, 1);
      if (P.max_bond == 2 && k == 2) {
        // D=2 truncation (core.py:16-18): a degenerate kept pair, or a cut through a degenerate pair, leaves the D=2
        // state / its bond gauge -- and with it the completed unitaries -- ill-defined: report it
        const double s0 = smax, s1 = norms[perm[1]], s2 = nrows > 2 ? norms[perm[2]] : -1.0;
        if (s0 - s1 < QSP_DEGENERATE_GAP * s0 || (s2 >= 0.0 && s1 - s2 < QSP_DEGENERATE_GAP * s0 && s2 > P.cutoff * s0))
          st |= QSP_ST_DEGENERATE;
      }
      if (P.svals) {
        double* sv = P.svals + ((long)inst * (L - 1) + (i - 1)) * P.nsv;
        for (int q = tid; q < P.nsv; q += nt) sv[q] = (q < nrows) ? norms[perm[q]] : 0.0;
      }
      // ---- Vh -> out site i (k x 2 x rp) and compact copy Vg (k x c); U s -> Cg (b x k) --
      QSP_TOCK(3);
      if (modeC) {
        for (int idx = tid; idx < k * c; idx += nt) {
          const int q = idx / c, cc = idx - q * c;
          const cplx v = cconj(Wg[(long)perm[q] * len + joff + cc]);
          const int p = cc / rp, kk = cc - p * rp;
          Oi[((long)q * 2 + p) * Dout + kk] = v;
          Vg[idx] = v;
        }
        for (int idx = tid; idx < b * k; idx += nt) {
          const int x = idx / k, q = idx - x * k;
          Cg[(long)x * DoutL + q] = Wg[(long)perm[q] * len + x];
        }
        __syncthreads();
      } else {
        for (int idx = tid; idx < k * c; idx += nt) {
          const int q = idx / c, cc = idx - q * c;
          const int r = perm[q];
          const double s = norms[r];
          const cplx v = (s > 0.0) ? cscale(Wg[(long)r * len + cc], 1.0 / s) : cmake(0.0, 0.0);
          const int p = cc / rp, kk = cc - p * rp;
          Oi[((long)q * 2 + p) * Dout + kk] = v;
          Vg[idx] = v;
        }
        __syncthreads();
        // U s = M Vh^H
        cta_zgemm_deep(b, k, c, Mg, c, 1, false, Vg, 1, c, true, Cg, DoutL, 1, false, tiles);
          pc[7] += 8LL * (b) * (k) * (c);
      }
      QSP_TOCK(4);
      // ---- <0...0| amplitude: f_new[q] = sum_kk Vh[q,(0,kk)] f[kk] ---------------------------
      if (P.amp0) {
        const cplx* f = fg + fcur * DoutL;
        cplx* fn = fg + (1 - fcur) * DoutL;
        for (int q = tid; q < k; q += nt) {
          cplx acc = cmake(0.0, 0.0);
          for (int kk = 0; kk < rp; ++kk) cfma(acc, Vg[(long)q * c + kk], f[kk]);
          fn[q] = acc;
        }
        fcur = 1 - fcur;
      }
      if (tid == 0) bo[i] = k;
      rp = k;
      __syncthreads();
    }
    if (*s_bad) st |= QSP_ST_BADNORM;
    if (tid == 0 && P.status && st) P.status[inst] |= st;
    __syncthreads();
  }
  if (P.prof && tid == 0)
    for (int q = 0; q < 8; ++q) P.prof[(long)blockIdx.x * 8 + q] += pc[q];
#undef QSP_TICK
#undef QSP_TOCK
}
