// Isometry -> unitary completion (core.py:25-123), exact (inverse) layer application as a
// bond-doubling MPO product (core.py:126-141, 224-240 without the re-balancing SVD) and the
// transfer-matrix overlap (helper_routines.py:65-73).
//
// Gate storage: G[B][L][4][4] complex128 row-major, the matrices the reference keeps in
// Gs[i].data (core.py:46,71-73,105-107): a two-qubit gate on (it, it+1) has row index
// (p_it, r) and column index (l, i_fresh); a one-qubit gate (segment end) sits in the
// top-left 2x2 block.  kind[B][L]: 2 = two-qubit gate starting at this site, 1 = one-qubit
// gate (this site ends a segment).  Segment starts are the sites following a kind-1 site.
#pragma once
#include "qsp_common.cuh"
#include "qsp_gemm.cuh"

#define QSP_TIE 0.999  // pivot = lowest index within TIE of the maximum (oracle: TIE)

__device__ inline int pivot_index(const double* mag, int n) {
  double mx = mag[0];
  for (int i = 1; i < n; ++i) mx = fmax(mx, mag[i]);
  for (int i = 0; i < n; ++i)
    if (mag[i] >= QSP_TIE * mx) return i;
  return 0;
}

// v: n x m (v[i*m + c]) orthonormal columns -> K: n x (n-m) (K[i*(n-m) + t]) orthonormal
// complement by Gram-Schmidt on the standard basis with pivoting (oracle: complete_pivoted).
__device__ inline void complete_pivoted(const cplx* v, int n, int m, cplx* K) {
  cplx P[16];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      cplx s = cmake(i == j ? 1.0 : 0.0, 0.0);
      for (int c = 0; c < m; ++c) s = csub(s, cmul(v[i * m + c], cconj(v[j * m + c])));
      P[i * n + j] = s;
    }
  const int nk = n - m;
  for (int t = 0; t < nk; ++t) {
    double rho[4];
    for (int i = 0; i < n; ++i) rho[i] = P[i * n + i].x;
    const int ip = pivot_index(rho, n);
    const double inv = 1.0 / sqrt(rho[ip]);
    cplx k[4];
    for (int i = 0; i < n; ++i) k[i] = cscale(P[i * n + ip], inv);
    for (int i = 0; i < n; ++i) K[i * nk + t] = k[i];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) P[i * n + j] = csub(P[i * n + j], cmul(k[i], cconj(k[j])));
  }
}

// One thread per instance: gauge-fix the D<=2 tensors (in place) and emit the gates.
// d2: [B][L][2][2][2] (element (l,p,r) at (l*2+p)*2 + r), right canonical, normalised.
__global__ void k_unitaries(cplx* __restrict__ d2, const int32_t* __restrict__ d2_bonds, int B, int L,
                            cplx* __restrict__ G, int32_t* __restrict__ kind, int32_t* __restrict__ status) {
  const int inst = blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= B) return;
  cplx* A = d2 + (long)inst * L * 8;
  const int32_t* bd = d2_bonds + (long)inst * (L + 1);
  cplx* Gi = G + (long)inst * L * 16;
  int32_t* kd = kind + (long)inst * L;
  int st = 0;
  // ---- phase gauge: right to left, pivot entry of every row real positive ------------------
  for (int j = L - 1; j >= 1; --j) {
    const int l = bd[j], r = bd[j + 1];
    for (int k = 0; k < l; ++k) {
      double mag[4];
      int n = 0;
      for (int p = 0; p < 2; ++p)
        for (int rr = 0; rr < r; ++rr) mag[n++] = cabs2(A[j * 8 + (k * 2 + p) * 2 + rr]);
      const int m = pivot_index(mag, n);
      const int p = m / r, rr = m - p * r;
      const cplx z = A[j * 8 + (k * 2 + p) * 2 + rr];
      const double az = sqrt(cabs2(z));
      if (!(az > 0.0)) {
        st |= QSP_ST_BADNORM;
        continue;
      }
      const cplx ph = cscale(z, 1.0 / az);
      for (int pp = 0; pp < 2; ++pp)
        for (int r2 = 0; r2 < r; ++r2) {
          cplx* e = &A[j * 8 + (k * 2 + pp) * 2 + r2];
          *e = cmul(*e, cconj(ph));
        }
      const int lprev = bd[j - 1];
      for (int ll = 0; ll < lprev; ++ll)
        for (int pp = 0; pp < 2; ++pp) {
          cplx* e = &A[(j - 1) * 8 + (ll * 2 + pp) * 2 + k];
          *e = cmul(*e, ph);
        }
    }
  }
  // ---- gates ---------------------------------------------------------------------------------
  for (int it = 0; it < L; ++it) {
    const int dl = bd[it], dr = bd[it + 1];
    const cplx* a = A + it * 8;
    cplx g[16];
    for (int q = 0; q < 16; ++q) g[q] = cmake(0.0, 0.0);
    int kn;
    if (dl < 2 && dr < 2) {  // single-site segment (core.py:36-41)
      cplx v[2], w[2];
      v[0] = a[0 * 2 + 0];
      v[1] = a[1 * 2 + 0];
      complete_pivoted(v, 2, 1, w);
      g[0 * 4 + 0] = v[0];
      g[1 * 4 + 0] = v[1];
      g[0 * 4 + 1] = w[0];
      g[1 * 4 + 1] = w[1];
      kn = 1;
    } else if (dl >= 2 && dr < 2) {  // segment end (core.py:43-46): G[p][l] = A[l,p,0]
      for (int p = 0; p < 2; ++p)
        for (int l = 0; l < 2; ++l) g[p * 4 + l] = a[(l * 2 + p) * 2 + 0];
      kn = 1;
    } else if (dl >= 2) {  // middle (core.py:55-73): rows (p,r), columns (l,i)
      cplx V[8], K[8];
      for (int p = 0; p < 2; ++p)
        for (int r = 0; r < 2; ++r)
          for (int l = 0; l < 2; ++l) V[(p * 2 + r) * 2 + l] = a[(l * 2 + p) * 2 + r];
      complete_pivoted(V, 4, 2, K);
      for (int row = 0; row < 4; ++row)
        for (int l = 0; l < 2; ++l) {
          g[row * 4 + (l * 2 + 0)] = V[row * 2 + l];
          g[row * 4 + (l * 2 + 1)] = K[row * 2 + l];
        }
      kn = 2;
    } else {  // segment start (core.py:87-107)
      cplx v[4], K[12];
      for (int p = 0; p < 2; ++p)
        for (int r = 0; r < 2; ++r) v[p * 2 + r] = a[(0 * 2 + p) * 2 + r];
      complete_pivoted(v, 4, 1, K);
      for (int row = 0; row < 4; ++row) {
        g[row * 4 + 0] = v[row];
        for (int l = 0; l < 2; ++l)
          for (int i = 0; i < 2; ++i) {
            if (l == 0 && i == 0) continue;
            g[row * 4 + (l * 2 + i)] = K[row * 3 + (i * 2 + l - 1)];
          }
      }
      kn = 2;
    }
    // unitarity (helper_routines.py:56-62)
    const int n = kn == 2 ? 4 : 2;
    double err = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) {
        cplx s = cmake(i == j ? -1.0 : 0.0, 0.0);
        for (int c = 0; c < n; ++c) s = cadd(s, cmul(g[i * 4 + c], cconj(g[j * 4 + c])));
        err = fmax(err, fmax(fabs(s.x), fabs(s.y)));
      }
    if (!(err < 1e-8)) st |= QSP_ST_NONUNITARY;
    for (int q = 0; q < 16; ++q) Gi[it * 16 + q] = g[q];
    kd[it] = kn;
  }
  if (st && status) status[inst] |= st;
}

// U^dagger |psi> for a LEFT-canonical psi as an exact bond-doubling product; the result is
// left canonical.  One CTA per instance.
__global__ void k_expand_inverse(const cplx* __restrict__ in, const int32_t* __restrict__ bonds_in, int Din,
                                 const cplx* __restrict__ G, const int32_t* __restrict__ kind, cplx* __restrict__ out,
                                 int32_t* __restrict__ bonds_out, int Dout, int B, int L,
                                 int32_t* __restrict__ status) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const long sin = (long)Din * 2 * Din, sout = (long)Dout * 2 * Dout;
  for (int inst = blockIdx.x; inst < B; inst += gridDim.x) {
    const cplx* A = in + (long)inst * L * sin;
    const int32_t* bi = bonds_in + (long)inst * (L + 1);
    const cplx* Gi = G + (long)inst * L * 16;
    const int32_t* kd = kind + (long)inst * L;
    cplx* O = out + (long)inst * L * sout;
    int32_t* bo = bonds_out + (long)inst * (L + 1);
    bool bad = false;
    for (int t = 0; t < L; ++t) {
      const bool start = (t == 0) || (kd[t - 1] == 1);
      const int nb = start ? bi[t] : 2 * bi[t - 1];
      if (nb > Dout || bi[t] > Din) bad = true;
    }
    if (bi[L] > Dout) bad = true;
    if (bad) {
      if (tid == 0 && status) status[inst] |= QSP_ST_BOND_OVERFLOW;
      continue;
    }
    for (int t = 0; t < L; ++t) {
      const bool start = (t == 0) || (kd[t - 1] == 1);
      cplx* Ot = O + (long)t * sout;
      if (tid == 0) bo[t] = start ? bi[t] : 2 * bi[t - 1];
      if (t == L - 1 && tid == 0) bo[L] = bi[L];
      if (start && kd[t] == 1) {
        // one-site segment: out[l,q,r] = sum_p conj(G[p][q]) A[l,p,r]
        const cplx* At = A + (long)t * sin;
        const cplx* g = Gi + t * 16;
        const int l = bi[t], r = bi[t + 1];
        for (int idx = tid; idx < l * 2 * r; idx += nt) {
          const int row = idx / r, rr = idx - row * r;
          const int ll = row >> 1, q = row & 1;
          cplx s = cmake(0.0, 0.0);
          for (int p = 0; p < 2; ++p) cfma_conja(s, g[p * 4 + q], At[((long)ll * 2 + p) * Din + rr]);
          Ot[(long)row * Dout + rr] = s;
        }
      } else if (start) {
        // identity embedding (l, a') -> k
        const int l = bi[t];
        for (int idx = tid; idx < l * 2 * 2 * l; idx += nt) {
          const int row = idx / (2 * l), k = idx - row * (2 * l);
          Ot[(long)row * Dout + k] = cmake(row == k ? 1.0 : 0.0, 0.0);
        }
      } else {
        const int j = t - 1;
        const cplx* Aj = A + (long)j * sin;
        const cplx* g = Gi + j * 16;  // Gd[(a'b'),(ab)] = conj(g[(ab)][(a'b')])
        const int l = bi[j], x = bi[j + 1];
        if (kd[t] == 2) {
          // out[(l,a'), b', (x,b)] = sum_a Gd[(a'b'),(a b)] A_j[l,a,x]
          const int ncol = 2 * x;
          for (int idx = tid; idx < l * 2 * 2 * ncol; idx += nt) {
            const int row = idx / ncol, col = idx - row * ncol;  // row = ((l,a'),b')
            const int bp = row & 1, ap = (row >> 1) & 1, ll = row >> 2;
            const int xx = col >> 1, b = col & 1;
            cplx s = cmake(0.0, 0.0);
            for (int a = 0; a < 2; ++a)
              cfma_conja(s, g[(a * 2 + b) * 4 + (ap * 2 + bp)], Aj[((long)ll * 2 + a) * Din + xx]);
            Ot[(long)row * Dout + col] = s;
          }
        } else {
          // segment end: out[(l,a'), b', r] = sum_{a,b,x} Gd_j[(a'b'),(ab)] A_j[l,a,x] (Gd_e A_e)[x,b,r]
          const cplx* Ae = A + (long)t * sin;
          const cplx* ge = Gi + t * 16;
          const int r = bi[t + 1];
          for (int idx = tid; idx < l * 2 * 2 * r; idx += nt) {
            const int row = idx / r, rr = idx - row * r;
            const int bp = row & 1, ap = (row >> 1) & 1, ll = row >> 2;
            cplx s = cmake(0.0, 0.0);
            for (int xx = 0; xx < x; ++xx) {
              cplx ae[2];  // (Gd_e A_e)[xx, b, rr] = sum_p conj(ge[p][b]) A_e[xx,p,rr]
              for (int b = 0; b < 2; ++b) {
                cplx u = cmake(0.0, 0.0);
                for (int p = 0; p < 2; ++p) cfma_conja(u, ge[p * 4 + b], Ae[((long)xx * 2 + p) * Din + rr]);
                ae[b] = u;
              }
              for (int a = 0; a < 2; ++a) {
                const cplx aj = Aj[((long)ll * 2 + a) * Din + xx];
                for (int b = 0; b < 2; ++b) {
                  const cplx gd = cconj(g[(a * 2 + b) * 4 + (ap * 2 + bp)]);
                  cfma(s, cmul(gd, aj), ae[b]);
                }
              }
            }
            Ot[(long)row * Dout + rr] = s;
          }
        }
      }
    }
    __syncthreads();
  }
}

// U |psi> for a RIGHT-canonical psi as an exact bond-doubling product (right canonical result).
__global__ void k_expand_forward(const cplx* __restrict__ in, const int32_t* __restrict__ bonds_in, int Din,
                                 const cplx* __restrict__ G, const int32_t* __restrict__ kind, cplx* __restrict__ out,
                                 int32_t* __restrict__ bonds_out, int Dout, int B, int L,
                                 int32_t* __restrict__ status) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const long sin = (long)Din * 2 * Din, sout = (long)Dout * 2 * Dout;
  for (int inst = blockIdx.x; inst < B; inst += gridDim.x) {
    const cplx* A = in + (long)inst * L * sin;
    const int32_t* bi = bonds_in + (long)inst * (L + 1);
    const cplx* Gi = G + (long)inst * L * 16;
    const int32_t* kd = kind + (long)inst * L;
    cplx* O = out + (long)inst * L * sout;
    int32_t* bo = bonds_out + (long)inst * (L + 1);
    bool bad = false;
    for (int t = 0; t < L; ++t) {
      const int nb = (kd[t] == 2) ? 2 * bi[t + 2] : bi[t + 1];
      if (nb > Dout || bi[t] > Din) bad = true;
    }
    if (bi[0] > Dout) bad = true;
    if (bad) {
      if (tid == 0 && status) status[inst] |= QSP_ST_BOND_OVERFLOW;
      continue;
    }
    for (int t = 0; t < L; ++t) {
      const bool start = (t == 0) || (kd[t - 1] == 1);
      cplx* Ot = O + (long)t * sout;
      const cplx* g = Gi + t * 16;
      if (tid == 0) {
        if (t == 0) bo[0] = bi[0];
        bo[t + 1] = (kd[t] == 2) ? 2 * bi[t + 2] : bi[t + 1];
      }
      if (start && kd[t] == 1) {
        const cplx* At = A + (long)t * sin;
        const int l = bi[t], r = bi[t + 1];
        for (int idx = tid; idx < l * 2 * r; idx += nt) {
          const int row = idx / r, rr = idx - row * r;
          const int ll = row >> 1, q = row & 1;
          cplx s = cmake(0.0, 0.0);
          for (int p = 0; p < 2; ++p) cfma(s, g[q * 4 + p], At[((long)ll * 2 + p) * Din + rr]);
          Ot[(long)row * Dout + rr] = s;
        }
      } else if (start) {
        // theta_s[l, a', (b', r)] = sum_{a,b,x} G[(a'b'),(ab)] A_s[l,a,x] A_{s+1}[x,b,r]
        const cplx* As = A + (long)t * sin;
        const cplx* An = A + (long)(t + 1) * sin;
        const int l = bi[t], x = bi[t + 1], r = bi[t + 2];
        const int ncol = 2 * r;
        for (int idx = tid; idx < l * 2 * ncol; idx += nt) {
          const int row = idx / ncol, col = idx - row * ncol;
          const int ll = row >> 1, ap = row & 1;
          const int bp = col / r, rr = col - bp * r;
          cplx s = cmake(0.0, 0.0);
          for (int xx = 0; xx < x; ++xx)
            for (int a = 0; a < 2; ++a) {
              const cplx as = As[((long)ll * 2 + a) * Din + xx];
              for (int b = 0; b < 2; ++b)
                cfma(s, cmul(g[(ap * 2 + bp) * 4 + (a * 2 + b)], as), An[((long)xx * 2 + b) * Din + rr]);
            }
          Ot[(long)row * Dout + col] = s;
        }
      } else if (kd[t] == 2) {
        // out[(a,x), a', (b', r)] = sum_b G_t[(a'b'),(a b)] A_{t+1}[x,b,r]
        const cplx* An = A + (long)(t + 1) * sin;
        const int x = bi[t + 1], r = bi[t + 2];
        const int ncol = 2 * r;
        for (int idx = tid; idx < 2 * x * 2 * ncol; idx += nt) {
          const int row = idx / ncol, col = idx - row * ncol;  // row = ((a,x), a')
          const int ap = row & 1, lx = row >> 1;
          const int a = lx / x, xx = lx - a * x;
          const int bp = col / r, rr = col - bp * r;
          cplx s = cmake(0.0, 0.0);
          for (int b = 0; b < 2; ++b) cfma(s, g[(ap * 2 + bp) * 4 + (a * 2 + b)], An[((long)xx * 2 + b) * Din + rr]);
          Ot[(long)row * Dout + col] = s;
        }
      } else {
        // segment end: out[(b, r), p', r2] = G_e[p'][b] delta(r, r2)
        const int r = bi[t + 1];
        for (int idx = tid; idx < 2 * r * 2 * r; idx += nt) {
          const int row = idx / r, r2 = idx - row * r;  // row = ((b,r), p')
          const int pp = row & 1, br = row >> 1;
          const int b = br / r, rr = br - b * r;
          Ot[(long)row * Dout + r2] = (rr == r2) ? g[pp * 4 + b] : cmake(0.0, 0.0);
        }
      }
    }
    __syncthreads();
  }
}

// raw inner product  sum_x a(x) conj(b(x))  by the transfer-matrix contraction.
// work per CTA: E0, E1 (Da x Db each) and T (Da x 2 Db).
__host__ __device__ inline long overlap_work_elems(int Da, int Db) { return 2L * Da * Db + 2L * Da * Db + 64; }

__global__ void __launch_bounds__(256, 2) k_overlap(const cplx* __restrict__ a, const int32_t* __restrict__ bonds_a, int Da,
                          const cplx* __restrict__ b, const int32_t* __restrict__ bonds_b, int Db, int B, int L,
                          cplx* __restrict__ result, cplx* __restrict__ work, long work_stride) {
  cplx* tiles = (cplx*)QSP_SMEM_BASE();
  const int tid = threadIdx.x;
  cplx* E0 = work + (long)blockIdx.x * work_stride;
  cplx* E1 = E0 + (long)Da * Db;
  cplx* T = E1 + (long)Da * Db;
  const long sa = (long)Da * 2 * Da, sb = (long)Db * 2 * Db;
  for (int inst = blockIdx.x; inst < B; inst += gridDim.x) {
    const cplx* A = a + (long)inst * L * sa;
    const cplx* Bm = b + (long)inst * L * sb;
    const int32_t* ba = bonds_a + (long)inst * (L + 1);
    const int32_t* bb = bonds_b + (long)inst * (L + 1);
    cplx* E = E0;
    cplx* En = E1;
    if (tid == 0) E[0] = cmake(1.0, 0.0);  // bonds[0] == 1 on both sides
    __syncthreads();
    for (int i = 0; i < L; ++i) {
      const int xa = ba[i], xb = bb[i], ya = ba[i + 1], yb = bb[i + 1];
      const cplx* Ai = A + (long)i * sa;
      const cplx* Bi = Bm + (long)i * sb;
      // T[x,(p,y')] = sum_y E[x,y] conj(B[y,p,y'])
      for (int p = 0; p < 2; ++p)
        cta_zgemm_deep(xa, yb, xb, E, xb, 1, false, Bi + (long)p * Db, 2L * Db, 1, true, T + (long)p * yb, 2L * yb, 1, false,
                       tiles);
      // E'[x',y'] = sum_{(x,p)} A[x,p,x'] T[(x,p),y']
      cta_zgemm_deep(ya, yb, 2 * xa, Ai, 1, Da, false, T, yb, 1, false, En, yb, 1, false, tiles);
      cplx* tmp = E;
      E = En;
      En = tmp;
    }
    if (tid == 0) result[inst] = E[0];
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------
// MPS linear combination (lcu/lcu.py:75,89-92: `mps - renom * approx`, `approx + kappa * residual`):
// out = alpha |x> + beta |y> as the direct sum of the bond spaces -- site 0 concatenates along the right
// bond (and carries the two scalars), the last site along the left bond, sites in between are block
// diagonal; a chain of one site is the plain sum.  Elementwise, one CTA per instance.
// ----------------------------------------------------------------------------------
__global__ void k_mps_axpby(const cplx* __restrict__ x, const int32_t* __restrict__ bx, int Dx, const cplx* __restrict__ y,
                            const int32_t* __restrict__ by, int Dy, const cplx* __restrict__ alpha,
                            const cplx* __restrict__ beta, cplx* __restrict__ out, int32_t* __restrict__ bo, int Do, int B,
                            int L, int32_t* __restrict__ status) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const long sx = (long)Dx * 2 * Dx, sy = (long)Dy * 2 * Dy, so = (long)Do * 2 * Do;
  for (int inst = blockIdx.x; inst < B; inst += gridDim.x) {
    const cplx* X = x + (long)inst * L * sx;
    const cplx* Y = y + (long)inst * L * sy;
    cplx* O = out + (long)inst * L * so;
    const int32_t* bxi = bx + (long)inst * (L + 1);
    const int32_t* byi = by + (long)inst * (L + 1);
    int32_t* boi = bo + (long)inst * (L + 1);
    const cplx al = alpha[inst], be = beta[inst];
    bool bad = false;
    for (int i = 0; i <= L; ++i) {
      const bool edge = (i == 0 || i == L);
      if (bxi[i] > Dx || byi[i] > Dy || (edge ? 1 : bxi[i] + byi[i]) > Do || (edge && (bxi[i] != 1 || byi[i] != 1))) bad = true;
    }
    if (bad) {
      if (tid == 0 && status) status[inst] |= QSP_ST_BOND_OVERFLOW;
      continue;
    }
    for (int i = 0; i < L; ++i) {
      const int lx = bxi[i], rx = bxi[i + 1], ly = byi[i], ry = byi[i + 1];
      const int lo = (i == 0) ? 1 : lx + ly, ro = (i == L - 1) ? 1 : rx + ry;
      const int loff = (i == 0) ? 0 : lx, roff = (i == L - 1) ? 0 : rx;  // where the y block starts
      const cplx* Xi = X + (long)i * sx;
      const cplx* Yi = Y + (long)i * sy;
      cplx* Oi = O + (long)i * so;
      for (int idx = tid; idx < lo * 2 * ro; idx += nt) {
        const int row = idx / ro, r = idx - row * ro;
        const int l = row >> 1, p = row & 1;
        cplx v = cmake(0.0, 0.0);
        if (l < lx && r < rx) {
          const cplx a = Xi[((long)l * 2 + p) * Dx + r];
          v = (i == 0) ? cmul(al, a) : a;
        }
        const int l2 = l - loff, r2 = r - roff;
        if (l2 >= 0 && l2 < ly && r2 >= 0 && r2 < ry) {
          const cplx b = Yi[((long)l2 * 2 + p) * Dy + r2];
          v = cadd(v, (i == 0) ? cmul(be, b) : b);
        }
        Oi[((long)l * 2 + p) * Do + r] = v;
      }
      if (tid == 0) {
        boi[i] = lo;
        boi[i + 1] = ro;
      }
    }
  }
}

// |x> <- alpha |x> in place (the scalar goes onto site 0)
__global__ void k_mps_scale(cplx* __restrict__ x, const int32_t* __restrict__ bx, int Dx, const cplx* __restrict__ alpha, int B,
                            int L) {
  const long sx = (long)Dx * 2 * Dx;
  for (int inst = blockIdx.x; inst < B; inst += gridDim.x) {
    cplx* X = x + (long)inst * L * sx;
    const int r = bx[(long)inst * (L + 1) + 1];
    const cplx al = alpha[inst];
    for (int idx = threadIdx.x; idx < 2 * r; idx += blockDim.x) {
      const int p = idx / r, rr = idx - p * r;
      X[(long)p * Dx + rr] = cmul(al, X[(long)p * Dx + rr]);
    }
  }
}
