// Right-compression sweep (quimb right_compress, and the fused gate-apply-and-split of the layer loop) for bonds
// that do not fit one SM: the same mathematics, argument block and scratch partition as k_right_compress
// (qsp_sweeps.cuh), but every step of a site is a batched launch over (instance x tile) -- GEMMs and the block
// Jacobi SVD on the FP64 tensor cores (qsp_big.cuh), small elementwise kernels in between.  The host side
// (qsp_capi.cu: big_right_compress) walks the chain and only knows upper bounds of the extents; the live
// extents of every instance stay on the device (BigRCState).
#pragma once
#include "qsp_big.cuh"
#include "qsp_sweeps.cuh"

__device__ __forceinline__ BigRCWork big_rc_work(const RCompressArgs& P, int inst) {
  BigRCWork w;
  w.Mg = P.work + (long)inst * P.work_stride;
  w.Wg = w.Mg + (long)P.DinL * 2 * P.DoutL;
  w.Vg = w.Wg + (long)2 * P.DoutL * (P.DinL + 2 * P.DoutL + 64);
  w.Cg = w.Vg + (long)P.DoutL * 2 * P.DoutL;
  w.fg = w.Cg + (long)P.DinL * P.DoutL;
  return w;
}
__host__ __device__ inline int big_r32(int x) { return (x + QSP_BIG_CK - 1) / QSP_BIG_CK * QSP_BIG_CK; }

__global__ void k_big_rc_init(RCompressArgs P, BigRCState* __restrict__ S) {
  const int inst = blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= P.B) return;
  const int L = P.L;
  const int32_t* bi = P.bonds_in + (long)inst * (L + 1);
  int32_t* bo = P.bonds_out + (long)inst * (L + 1);
  const bool fused = P.G != nullptr;
  const int32_t* kd = fused ? P.kind + (long)inst * L : nullptr;
  bool bad = false;
  for (int i = 0; i <= L; ++i) {
    int bb = bi[i];
    if (fused && i > 0 && i < L && kd[i - 1] == 2) bb = 2 * bi[i - 1];
    if (bb > P.DinL || bi[i] < 1 || (fused && bi[i] > P.Din)) bad = true;
  }
  if (bi[L] > P.DoutL || bi[0] > P.DoutL) bad = true;
  BigRCState s;
  memset(&s, 0, sizeof(s));
  s.bad = bad ? 1 : 0;
  s.rp = bi[L];
  S[inst] = s;
  if (bad) {
    if (P.status) P.status[inst] |= QSP_ST_BOND_OVERFLOW;
    return;
  }
  bo[L] = bi[L];
  bo[0] = bi[0];
  if (P.amp0) {
    BigRCWork w = big_rc_work(P, inst);
    for (int k = 0; k < s.rp; ++k) w.fg[k] = cmake(k == 0 ? 1.0 : 0.0, 0.0);
  }
}

// extents of site i and the operand descriptors of its GEMMs: g1 (T or M = A_i C), g2[2] (P = A_{i-1} C, two halves)
__global__ void k_big_rc_site(RCompressArgs P, BigRCState* __restrict__ S, int i, BigGemm* __restrict__ g1, BigGemm* __restrict__ g2) {
  const int inst = blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= P.B) return;
  BigGemm z;
  memset(&z, 0, sizeof(z));
  g1[inst] = z;
  g2[2 * inst] = z;
  g2[2 * inst + 1] = z;
  BigRCState s = S[inst];
  if (s.bad) return;
  const int L = P.L, Din = P.Din, DoutL = P.DoutL;
  const int32_t* bi = P.bonds_in + (long)inst * (L + 1);
  const bool fused = P.G != nullptr;
  const int32_t* kd = fused ? P.kind + (long)inst * L : nullptr;
  const long sin = (long)Din * 2 * Din;
  const cplx* A = P.in + (long)inst * L * sin;
  const cplx* Ai = A + (long)i * sin;
  BigRCWork w = big_rc_work(P, inst);
  const int rp = s.rp;
  const bool start = fused && (i == 0 || kd[i - 1] == 1);
  const int b = (!fused || start) ? bi[i] : 2 * bi[i - 1];
  const int ro = (!fused || kd[i] == 1) ? bi[i + 1] : 2 * bi[i];
  const int c = 2 * rp;
  s.b = b;
  s.c = c;
  s.ro = ro;
  s.start = start ? 1 : 0;
  s.kd = fused ? kd[i] : 0;
  BigGemm g;
  if (!fused) {
    if (i != L - 1) {  // M[(x,p),k] = sum_r A[x,p,r] C[r,k]
      g = z;
      g.A = Ai, g.sAr = Din, g.sAk = 1;
      g.B = w.Cg, g.sBk = DoutL, g.sBc = 1;
      g.C = w.Mg, g.sCr = rp, g.sCc = 1;
      g.m = 2 * b, g.n = rp, g.k = ro;
      g1[inst] = g;
    }
  } else if (!(start && kd[i] == 2)) {
    const cplx* Cuse = w.Cg;
    long ldC = DoutL;
    if (kd[i] == 1) {
      const int xe = bi[i], re = bi[i + 1];
      if (i != L - 1) {  // T[(x,p),k] = sum_r A_i[x,p,r] C[r,k]  (Tg aliases the Jacobi matrix)
        g = z;
        g.A = Ai, g.sAr = Din, g.sAk = 1;
        g.B = w.Cg, g.sBk = DoutL, g.sBc = 1;
        g.C = w.Wg, g.sCr = rp, g.sCc = 1;
        g.m = 2 * xe, g.n = rp, g.k = re;
        g1[inst] = g;
      }
      Cuse = w.Vg;
      ldC = rp;
    }
    if (!start) {  // P[(l,a),(bq,k)] = sum_x A_{i-1}[l,a,x] C[(x,bq),k]
      const cplx* Aj = A + (long)(i - 1) * sin;
      const int lj = bi[i - 1], xj = bi[i];
      for (int bq = 0; bq < 2; ++bq) {
        g = z;
        g.A = Aj, g.sAr = Din, g.sAk = 1;
        g.B = Cuse + (long)bq * ldC, g.sBk = 2 * ldC, g.sBc = 1;
        g.C = w.Wg + (long)bq * rp, g.sCr = 2L * rp, g.sCc = 1;
        g.m = 2 * lj, g.n = rp, g.k = xj;
        g2[2 * inst + bq] = g;
      }
    }
  }
  if (i > 0) {
    s.modeC = b > c ? 1 : 0;
    s.nrows = s.modeC ? c : b;
    s.ndot = s.modeC ? b : c;
    s.ndotp = big_r32(s.ndot);
    s.len = s.modeC ? s.ndotp + big_r32(c) : s.ndotp;
    s.ldw = s.len;
    s.joff = s.ndotp;
  }
  S[inst] = s;
}

// elementwise steps of a site; blockIdx.x = inst * parts + part
__global__ void __launch_bounds__(256) k_big_rc_elem(RCompressArgs P, const BigRCState* __restrict__ S, int i, int phase, int parts) {
  const int inst = blockIdx.x / parts, part = blockIdx.x - inst * parts;
  const BigRCState s = S[inst];
  if (s.bad) return;
  const int L = P.L, Din = P.Din, DoutL = P.DoutL;
  const bool fused = P.G != nullptr;
  const long sin = (long)Din * 2 * Din;
  const cplx* Ai = P.in + ((long)inst * L + i) * sin;
  const int32_t* bi = P.bonds_in + (long)inst * (L + 1);
  BigRCWork w = big_rc_work(P, inst);
  const int rp = s.rp, b = s.b, c = s.c;
  const int t0 = part * blockDim.x + threadIdx.x, stride = parts * blockDim.x;
  const bool start = s.start != 0;
  if (phase == 0) {
    if (!fused) {
      if (i == L - 1)
        for (int idx = t0; idx < 2 * b * rp; idx += stride) {
          const int row = idx / rp, k = idx - row * rp;
          w.Mg[idx] = Ai[(long)row * Din + k];
        }
    } else if (start && s.kd == 2) {
      // first site of a multi-site segment: identity embedding, M[l,(a',k)] = C[(l,a'),k]
      for (int idx = t0; idx < b * c; idx += stride) {
        const int l = idx / c, col = idx - l * c;
        const int ap = col / rp, k = col - ap * rp;
        w.Mg[idx] = w.Cg[((long)l * 2 + ap) * DoutL + k];
      }
    } else if (s.kd == 1 && i == L - 1) {
      const int xe = bi[i];
      for (int idx = t0; idx < 2 * xe * rp; idx += stride) {
        const int row = idx / rp, k = idx - row * rp;
        w.Wg[idx] = Ai[(long)row * Din + k];
      }
    }
  } else if (phase == 1) {
    if (fused && s.kd == 1 && !(start && s.kd == 2)) {
      // Q[(x,q),k] = sum_p conj(G_i[p][q]) T[(x,p),k]   (G_i^dagger on the physical leg)
      const int xe = bi[i];
      const cplx* ge = P.G + ((long)inst * L + i) * 16;
      const cplx* Tg = w.Wg;
      cplx* Qg = start ? w.Mg : w.Vg;
      for (int idx = t0; idx < 2 * xe * rp; idx += stride) {
        const int row = idx / rp, k = idx - row * rp;
        const int x = row >> 1, q = row & 1;
        cplx acc = cmake(0.0, 0.0);
        for (int p = 0; p < 2; ++p) cfma_conja(acc, ge[p * 4 + q], Tg[((long)x * 2 + p) * rp + k]);
        Qg[idx] = acc;
      }
    }
  } else {
    if (fused && !start) {
      // M[(l,a'),(b',k)] = sum_{a,b} conj(G_{i-1}[(a b)][(a' b')]) P[(l,a),(b,k)]
      const cplx* g = P.G + ((long)inst * L + (i - 1)) * 16;
      const cplx* Tg = w.Wg;
      const int lj = bi[i - 1];
      for (int idx = t0; idx < lj * rp; idx += stride) {
        const int l = idx / rp, k = idx - l * rp;
        cplx pv[4];
#pragma unroll
        for (int ab = 0; ab < 4; ++ab) pv[ab] = Tg[((long)l * 2 + (ab >> 1)) * (2 * rp) + (ab & 1) * rp + k];
#pragma unroll
        for (int apbp = 0; apbp < 4; ++apbp) {
          cplx acc = cmake(0.0, 0.0);
#pragma unroll
          for (int ab = 0; ab < 4; ++ab) cfma_conja(acc, g[ab * 4 + apbp], pv[ab]);
          w.Mg[((long)l * 2 + (apbp >> 1)) * c + (apbp & 1) * rp + k] = acc;
        }
      }
    }
  }
}

// site 0 (orthogonality centre): out <- M (b x 2 x rp), optional normalisation, <0..0| amplitude; one CTA per instance
__global__ void __launch_bounds__(256) k_big_rc_final(RCompressArgs P, BigRCState* __restrict__ S) {
  const int inst = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  BigRCState s = S[inst];
  if (s.bad) return;
  double* red = (double*)QSP_SMEM_BASE();  // 64 doubles of dynamic shared memory
  BigRCWork w = big_rc_work(P, inst);
  const int b = s.b, rp = s.rp, Dout = P.Dout;
  cplx* Oi = P.out + (long)inst * P.L * ((long)Dout * 2 * Dout);
  double n2 = 0.0;
  for (int idx = tid; idx < 2 * b * rp; idx += nt) n2 += cabs2(w.Mg[idx]);
  n2 = block_sum(n2, red);
  double sc = 1.0;
  int st = s.st;
  if (P.normalize) {
    if (!(n2 > 0.0) || !(n2 < 1e300)) st |= QSP_ST_BADNORM;
    else sc = 1.0 / sqrt(n2);
  }
  for (int idx = tid; idx < 2 * b * rp; idx += nt) {
    const int row = idx / rp, k = idx - row * rp;
    Oi[(long)row * Dout + k] = cscale(w.Mg[idx], sc);
  }
  if (tid == 0) {
    if (P.amp0) {
      cplx acc = cmake(0.0, 0.0);
      const cplx* f = w.fg + s.fcur * P.DoutL;
      for (int k = 0; k < rp; ++k) cfma(acc, w.Mg[k], f[k]);
      const double nn = sqrt(n2);
      P.amp0[inst] = (nn > 0.0) ? cscale(acc, 1.0 / nn) : cmake(0.0, 0.0);
    }
    if (P.status && st) P.status[inst] |= st;
  }
}

// Jacobi operand: mode R  W = [ M 0 ] (b x ldw);  mode C  W = [ M^T 0 | I 0 ] (c x ldw)
__global__ void __launch_bounds__(256) k_big_rc_jprep(RCompressArgs P, const BigRCState* __restrict__ S, BigJac* __restrict__ jd, int parts) {
  const int inst = blockIdx.x / parts, part = blockIdx.x - inst * parts;
  const BigRCState s = S[inst];
  BigRCWork w = big_rc_work(P, inst);
  if (part == 0 && threadIdx.x == 0) {
    BigJac j;
    memset(&j, 0, sizeof(j));
    if (!s.bad && s.nrows >= 2) {
      j.W = w.Wg;
      j.nrows = s.nrows;
      j.ndotp = s.ndotp;
      j.len = s.len;
      j.ldw = s.ldw;
      j.nblk = ((s.nrows + QSP_BIG_RB - 1) / QSP_BIG_RB + 1) & ~1;
      j.inner = j.nblk == 2 ? 8 : 1;
      const double tol = 4e-16 * sqrt((double)s.ndot);
      j.tol2 = tol * tol;
    }
    jd[inst] = j;
  }
  if (s.bad) return;
  const int b = s.b, c = s.c, ldw = s.ldw;
  const long total = (long)s.nrows * ldw;
  for (long idx = (long)part * blockDim.x + threadIdx.x; idx < total; idx += (long)parts * blockDim.x) {
    const int r = (int)(idx / ldw), cc = (int)(idx - (long)r * ldw);
    cplx v = cmake(0.0, 0.0);
    if (s.modeC) {
      if (cc < b)
        v = w.Mg[(long)cc * c + r];
      else if (cc >= s.joff && cc - s.joff == r)
        v = cmake(1.0, 0.0);
    } else if (cc < c) {
      v = w.Mg[(long)r * c + cc];
    }
    w.Wg[idx] = v;
  }
}

__host__ __device__ inline size_t big_rank_smem_bytes(int nrows_max) { return (size_t)nrows_max * (sizeof(double) + sizeof(int)) + 64 * sizeof(double) + 64; }

// singular values = row norms: rank, truncate, emit Vh (and U s in mode C), amplitude vector; one CTA per instance.
// gU receives the operands of U s = M Vh^H (mode R).
__global__ void __launch_bounds__(256) k_big_rc_rank(RCompressArgs P, BigRCState* __restrict__ S, const BigJac* __restrict__ jd, int i,
                                                     BigGemm* __restrict__ gU) {
  const int inst = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  BigGemm z;
  memset(&z, 0, sizeof(z));
  BigRCState s = S[inst];
  if (s.bad) {
    if (tid == 0) gU[inst] = z;
    return;
  }
  char* sm = QSP_SMEM_BASE();
  const int nrows = s.nrows, ndot = s.ndot, len = s.ldw, b = s.b, c = s.c, rp = s.rp;
  double* norms = (double*)sm;
  int* perm = (int*)(norms + nrows);
  int* s_bad = perm + nrows;
  BigRCWork w = big_rc_work(P, inst);
  const int L = P.L, Dout = P.Dout, DoutL = P.DoutL;
  cplx* Oi = P.out + ((long)inst * L + i) * ((long)Dout * 2 * Dout);
  int32_t* bo = P.bonds_out + (long)inst * (L + 1);
  const int kmax = (P.max_bond > 0) ? imin(P.max_bond, DoutL) : DoutL;
  int st = s.st;
  if (jd[inst].nblk != 0) st |= QSP_ST_JACOBI_NOCONV;  // still unconverged when the sweep limit was reached
  if (tid == 0) *s_bad = 0;
  const cplx* Wsrc = w.Wg;  // staged by k_big_rc_jprep for every row count
  {
    const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    for (int r = warp; r < nrows; r += nw) {
      double s2 = 0.0;
      if (Wsrc) {
        for (int cc = lane; cc < ndot; cc += 32) s2 += cabs2(Wsrc[(long)r * len + cc]);
      } else {  // a single row was never staged: M itself (mode R: row of M; mode C: column of M)
        for (int cc = lane; cc < ndot; cc += 32) s2 += cabs2(s.modeC ? w.Mg[(long)cc * c + r] : w.Mg[(long)r * c + cc]);
      }
      for (int o = 16; o > 0; o >>= 1) s2 += shfl_xor_d(s2, o);
      if (lane == 0) norms[r] = sqrt(s2);
    }
  }
  __syncthreads();
  for (int r = tid; r < nrows; r += nt)
    if (!(norms[r] <= 1e300)) {
      norms[r] = 0.0;
      *s_bad = 1;
    }
  __syncthreads();
  for (int r = tid; r < nrows; r += nt) {
    const double sv = norms[r];
    int rank = 0;
    for (int q = 0; q < nrows; ++q) {
      const double t = norms[q];
      if (t > sv || (t == sv && q < r)) ++rank;
    }
    perm[rank] = r;
  }
  __syncthreads();
  if (*s_bad) st |= QSP_ST_BADNORM;
  const double smax = norms[perm[0]];
  int k = 0;
  for (int q = 0; q < nrows; ++q)
    if (norms[perm[q]] > P.cutoff * smax) ++k;
  if (k > kmax) {
    if (P.max_bond <= 0 || P.max_bond > DoutL) st |= QSP_ST_BOND_OVERFLOW;
    k = kmax;
  }
  k = imax(k, 1);
  if (P.max_bond == 2 && k == 2) {
    const double s0 = smax, s1 = norms[perm[1]], s2 = nrows > 2 ? norms[perm[2]] : -1.0;
    if (s0 - s1 < QSP_DEGENERATE_GAP * s0 || (s2 >= 0.0 && s1 - s2 < QSP_DEGENERATE_GAP * s0 && s2 > P.cutoff * s0))
      st |= QSP_ST_DEGENERATE;
  }
  if (P.svals) {
    double* sv = P.svals + ((long)inst * (L - 1) + (i - 1)) * P.nsv;
    for (int q = tid; q < P.nsv; q += nt) sv[q] = (q < nrows) ? norms[perm[q]] : 0.0;
  }
  // ---- Vh -> out site i (k x 2 x rp) and compact copy Vg (k x c); mode C: U s -> Cg (b x k) ----
  const int joff = s.joff;
  if (s.modeC) {
    for (int idx = tid; idx < k * c; idx += nt) {
      const int q = idx / c, cc = idx - q * c;
      const int r = perm[q];
      const cplx v = Wsrc ? cconj(Wsrc[(long)r * len + joff + cc]) : cmake(cc == r ? 1.0 : 0.0, 0.0);
      const int p = cc / rp, kk = cc - p * rp;
      Oi[((long)q * 2 + p) * Dout + kk] = v;
      w.Vg[idx] = v;
    }
    for (int idx = tid; idx < b * k; idx += nt) {
      const int x = idx / k, q = idx - x * k;
      const int r = perm[q];
      w.Cg[(long)x * DoutL + q] = Wsrc ? Wsrc[(long)r * len + x] : w.Mg[(long)x * c + r];
    }
  } else {
    for (int idx = tid; idx < k * c; idx += nt) {
      const int q = idx / c, cc = idx - q * c;
      const int r = perm[q];
      const double sv = norms[r];
      const cplx x = Wsrc ? Wsrc[(long)r * len + cc] : w.Mg[(long)r * c + cc];
      const cplx v = (sv > 0.0) ? cscale(x, 1.0 / sv) : cmake(0.0, 0.0);
      const int p = cc / rp, kk = cc - p * rp;
      Oi[((long)q * 2 + p) * Dout + kk] = v;
      w.Vg[idx] = v;
    }
  }
  __syncthreads();
  if (P.amp0) {
    const cplx* f = w.fg + s.fcur * DoutL;
    cplx* fn = w.fg + (1 - s.fcur) * DoutL;
    for (int q = tid; q < k; q += nt) {
      cplx acc = cmake(0.0, 0.0);
      for (int kk = 0; kk < rp; ++kk) cfma(acc, w.Vg[(long)q * c + kk], f[kk]);
      fn[q] = acc;
    }
  }
  if (tid == 0) {
    BigGemm g = z;
    if (!s.modeC) {  // U s = M Vh^H
      g.A = w.Mg, g.sAr = c, g.sAk = 1;
      g.B = w.Vg, g.sBk = 1, g.sBc = c, g.conjB = 1;
      g.C = w.Cg, g.sCr = DoutL, g.sCc = 1;
      g.m = b, g.n = k, g.k = c;
    }
    gU[inst] = g;
    bo[i] = k;
    s.k = k;
    s.rp = k;
    s.st = st;
    if (P.amp0) s.fcur = 1 - s.fcur;
    S[inst] = s;
  }
}

// ----------------------------------------------------------------------------------
// left canonisation, large bonds: blocked Householder QR (compact WY), trailing updates and the thin Q as
// batched tensor-core GEMMs; the panel factorisation is one CTA per instance.
//   A_i <- thin Q of (R_{i-1} A_i),  R with a non-negative diagonal carried to site i+1  (quimb left_canonize)
// ----------------------------------------------------------------------------------
#define QSP_BIG_NBW 16  // panel width

struct BigLCArgs {
  cplx* mps;        // [B][L][D][2][D], canonised in place
  int32_t* bonds;   // [B][L+1]
  int B, L, D, Dm;  // Dm: largest live bond the launch was planned for
  int32_t* status;
  cplx* work;       // per-instance scratch, big_lc_work_elems(Dm) each
  long work_stride;
};
struct BigLCState {
  int bad, a, a_old;
  int m, n, K;  // current site: S is m x n, K = min(m, n)
  int pad[2];
};
struct BigLCWork {
  cplx *S, *Q, *V, *VT, *VTh, *Wt, *T, *R, *tau;
  double* beta;
  int ldS, ldw;
};
__host__ __device__ inline long big_lc_work_elems(int Dm) {
  const long ldS = 2L * Dm;
  return 4 * ldS * Dm + ldS * QSP_BIG_NBW + (long)QSP_BIG_NBW * Dm + QSP_BIG_NBW * QSP_BIG_NBW + (long)Dm * Dm + 2L * Dm + 64;
}
__device__ __forceinline__ BigLCWork big_lc_work(const BigLCArgs& P, int inst) {
  BigLCWork w;
  const long ldS = 2L * P.Dm;
  w.ldS = (int)ldS;
  w.ldw = P.Dm;
  w.S = P.work + (long)inst * P.work_stride;  // m x n, column-major (ldS)
  w.Q = w.S + ldS * P.Dm;                     // m x K, column-major
  w.V = w.Q + ldS * P.Dm;                     // explicit reflectors (unit diagonal, zeros above), column-major
  w.VT = w.V + ldS * P.Dm;                    // V_p T_p of every panel
  w.VTh = w.VT + ldS * P.Dm;                  // V_p T_p^H of the current panel
  w.Wt = w.VTh + ldS * QSP_BIG_NBW;           // NBW x n row-major (ldw)
  w.T = w.Wt + (long)QSP_BIG_NBW * P.Dm;
  w.R = w.T + QSP_BIG_NBW * QSP_BIG_NBW;      // K x n row-major (ld Dm)
  w.tau = w.R + (long)P.Dm * P.Dm;
  w.beta = (double*)(w.tau + P.Dm);
  return w;
}

__global__ void k_big_lc_init(BigLCArgs P, BigLCState* __restrict__ S) {
  const int inst = blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= P.B) return;
  const int32_t* bd = P.bonds + (long)inst * (P.L + 1);
  bool bad = false;
  for (int i = 0; i <= P.L; ++i)
    if (bd[i] > P.Dm || bd[i] < 1) bad = true;
  BigLCState s;
  memset(&s, 0, sizeof(s));
  s.bad = bad ? 1 : 0;
  s.a = s.a_old = bd[0];
  S[inst] = s;
  if (bad && P.status) P.status[inst] |= QSP_ST_BOND_OVERFLOW;
}

// extents of site i and the operands of S = R A_i (two GEMMs, p = 0, 1)
__global__ void k_big_lc_site(BigLCArgs P, BigLCState* __restrict__ S, int i, BigGemm* __restrict__ g2) {
  const int inst = blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= P.B) return;
  BigGemm z;
  memset(&z, 0, sizeof(z));
  g2[2 * inst] = g2[2 * inst + 1] = z;
  BigLCState s = S[inst];
  if (s.bad) return;
  const int32_t* bd = P.bonds + (long)inst * (P.L + 1);
  const int D = P.D;
  s.m = 2 * s.a;
  s.n = bd[i + 1];
  s.K = imin(s.m, s.n);
  S[inst] = s;
  if (i == 0) return;  // copied by k_big_lc_elem
  BigLCWork w = big_lc_work(P, inst);
  const cplx* Ai = P.mps + ((long)inst * P.L + i) * ((long)D * 2 * D);
  for (int p = 0; p < 2; ++p) {  // S[(k,p), r] = sum_x R[k,x] A[x,p,r]
    BigGemm g = z;
    g.A = w.R, g.sAr = P.Dm, g.sAk = 1;
    g.B = Ai + (long)p * D, g.sBk = 2L * D, g.sBc = 1;
    g.C = w.S + p, g.sCr = 2, g.sCc = w.ldS;
    g.m = s.a, g.n = s.n, g.k = s.a_old;
    g2[2 * inst + p] = g;
  }
}

// phase 0: S <- A_0 (first site);  phase 1: last site, A_{L-1} <- S;  phase 2: R out (sign fixed), Q <- [I; 0];
// phase 3: A_i <- Q (sign fixed), bonds, carry
__global__ void __launch_bounds__(256) k_big_lc_elem(BigLCArgs P, BigLCState* __restrict__ S, int i, int phase, int parts) {
  const int inst = blockIdx.x / parts, part = blockIdx.x - inst * parts;
  BigLCState s = S[inst];
  if (s.bad) return;
  BigLCWork w = big_lc_work(P, inst);
  const int D = P.D, m = s.m, n = s.n, K = s.K, ldS = w.ldS;
  cplx* Ai = P.mps + ((long)inst * P.L + i) * ((long)D * 2 * D);
  const long t0 = (long)part * blockDim.x + threadIdx.x, stride = (long)parts * blockDim.x;
  if (phase == 0) {
    for (long idx = t0; idx < (long)m * n; idx += stride) {
      const int r = (int)(idx / n), c = (int)(idx - (long)r * n);
      w.S[(long)c * ldS + r] = Ai[(long)r * D + c];
    }
  } else if (phase == 1) {
    for (long idx = t0; idx < (long)m * n; idx += stride) {
      const int r = (int)(idx / n), c = (int)(idx - (long)r * n);
      Ai[(long)r * D + c] = w.S[(long)c * ldS + r];
    }
  } else if (phase == 2) {
    for (long idx = t0; idx < (long)K * n; idx += stride) {
      const int r = (int)(idx / n), c = (int)(idx - (long)r * n);
      cplx v = cmake(0.0, 0.0);
      if (c == r)
        v = cmake(w.beta[r], 0.0);
      else if (c > r)
        v = w.S[(long)c * ldS + r];
      if (w.beta[r] < 0.0) v = cmake(-v.x, -v.y);
      w.R[(long)r * P.Dm + c] = v;
    }
    for (long idx = t0; idx < (long)m * K; idx += stride) {
      const int c = (int)(idx / m), r = (int)(idx - (long)c * m);
      w.Q[(long)c * ldS + r] = cmake(r == c ? 1.0 : 0.0, 0.0);
    }
  } else {
    for (long idx = t0; idx < (long)m * K; idx += stride) {
      const int r = (int)(idx / K), c = (int)(idx - (long)r * K);
      const cplx q = w.Q[(long)c * ldS + r];
      Ai[(long)r * D + c] = w.beta[c] < 0.0 ? cmake(-q.x, -q.y) : q;
    }
    if (part == 0 && threadIdx.x == 0) {
      int32_t* bd = P.bonds + (long)inst * (P.L + 1);
      bd[i] = s.a;
      bd[i + 1] = K;
      s.a = K;
      s.a_old = n;
      S[inst] = s;
    }
  }
}

// Householder factorisation of the panel S[p0:m, p0:p0+jb] (in place, zgeqr2 conventions: H_j = I - tau_j v_j v_j^H,
// v_j[j] = 1, real diagonal beta_j), the triangular factor T of the compact WY form Q_p = H_1..H_jb = I - V T V^H,
// explicit V, V T and V T^H for the GEMMs, and the operands of the trailing update
//   C <- C - (V T^H) (V^H C),   C = S[p0:m, p0+jb:n].      One CTA per instance.
__global__ void __launch_bounds__(256) k_big_qr_panel(BigLCArgs P, const BigLCState* __restrict__ Sst, int p0, BigGemm* __restrict__ gW,
                                                      BigGemm* __restrict__ gC) {
  const int inst = blockIdx.x, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  BigGemm z;
  memset(&z, 0, sizeof(z));
  const BigLCState s = Sst[inst];
  if (s.bad || p0 >= s.K) {
    if (tid == 0) gW[inst] = gC[inst] = z;
    return;
  }
  double* red = (double*)QSP_SMEM_BASE();       // 64 doubles
  cplx* Z = (cplx*)(red + 64);                  // NBW x NBW Gram of the reflectors
  cplx* Ts = Z + QSP_BIG_NBW * QSP_BIG_NBW;     // NBW x NBW
  BigLCWork w = big_lc_work(P, inst);
  const int m = s.m, n = s.n, K = s.K, ldS = w.ldS;
  const int jb = imin(QSP_BIG_NBW, K - p0);
  const int mm = m - p0;                        // rows of the panel
  cplx* Pn = w.S + (long)p0 * ldS + p0;         // panel, column-major
  for (int j = 0; j < jb; ++j) {
    cplx* col = Pn + (long)j * ldS;
    // ---- reflector j from rows j .. mm-1 of column j ----
    double xn2 = 0.0;
    for (int r = j + 1 + tid; r < mm; r += nt) xn2 += cabs2(col[r]);
    xn2 = block_sum(xn2, red);
    const cplx alpha = col[j];
    cplx t, sc = cmake(0.0, 0.0);
    double bt;
    if (xn2 == 0.0 && alpha.y == 0.0) {
      t = cmake(0.0, 0.0);
      bt = alpha.x;
    } else {
      const double nrm = sqrt(cabs2(alpha) + xn2);
      bt = alpha.x >= 0.0 ? -nrm : nrm;
      const double ibt = 1.0 / bt;
      t = cmake((bt - alpha.x) * ibt, -alpha.y * ibt);
      const double dx = alpha.x - bt;
      const double id2 = 1.0 / fma(dx, dx, alpha.y * alpha.y);
      sc = cmake(dx * id2, -alpha.y * id2);  // 1 / (alpha - beta)
    }
    __syncthreads();  // everybody has read alpha
    if (t.x != 0.0 || t.y != 0.0)
      for (int r = j + 1 + tid; r < mm; r += nt) col[r] = cmul(col[r], sc);
    if (tid == 0) {
      w.tau[p0 + j] = t;
      w.beta[p0 + j] = bt;
      col[j] = cmake(bt, 0.0);
    }
    __syncthreads();
    // ---- H_j^H on the remaining panel columns: one warp per column ----
    if (t.x != 0.0 || t.y != 0.0) {
      const cplx f = cconj(t);
      for (int c = j + 1 + warp; c < jb; c += nw) {
        cplx* cc = Pn + (long)c * ldS;
        cplx wv = cmake(0.0, 0.0);
        for (int r = j + lane; r < mm; r += 32) {
          const cplx v = (r == j) ? cmake(1.0, 0.0) : col[r];
          cfma_conja(wv, v, cc[r]);
        }
        for (int o = 16; o > 0; o >>= 1) {
          wv.x += shfl_xor_d(wv.x, o);
          wv.y += shfl_xor_d(wv.y, o);
        }
        const cplx g = cmul(f, wv);
        for (int r = j + lane; r < mm; r += 32) {
          const cplx v = (r == j) ? cmake(1.0, 0.0) : col[r];
          cc[r] = csub(cc[r], cmul(g, v));
        }
      }
    }
    __syncthreads();
  }
  // ---- explicit V (mm x jb) ----
  cplx* Vp = w.V + (long)p0 * ldS + p0;
  for (int idx = tid; idx < mm * jb; idx += nt) {
    const int j = idx / mm, r = idx - j * mm;
    Vp[(long)j * ldS + r] = (r == j) ? cmake(1.0, 0.0) : (r > j ? Pn[(long)j * ldS + r] : cmake(0.0, 0.0));
  }
  __syncthreads();
  // ---- Z = V^H V (upper part), T by the forward recurrence ----
  for (int pr = warp; pr < jb * jb; pr += nw) {
    const int a = pr / jb, b = pr - a * jb;
    if (a >= b) continue;
    cplx acc = cmake(0.0, 0.0);
    for (int r = b + lane; r < mm; r += 32) cfma_conja(acc, Vp[(long)a * ldS + r], Vp[(long)b * ldS + r]);
    for (int o = 16; o > 0; o >>= 1) {
      acc.x += shfl_xor_d(acc.x, o);
      acc.y += shfl_xor_d(acc.y, o);
    }
    if (lane == 0) Z[a * QSP_BIG_NBW + b] = acc;
  }
  for (int idx = tid; idx < QSP_BIG_NBW * QSP_BIG_NBW; idx += nt) Ts[idx] = cmake(0.0, 0.0);
  __syncthreads();
  for (int j = 0; j < jb; ++j) {
    const cplx tj = w.tau[p0 + j];
    if (tid < j) {  // T[i][j] = -tau_j sum_{l=i}^{j-1} T[i][l] Z[l][j]
      cplx acc = cmake(0.0, 0.0);
      for (int l = tid; l < j; ++l) cfma(acc, Ts[tid * QSP_BIG_NBW + l], Z[l * QSP_BIG_NBW + j]);
      Ts[tid * QSP_BIG_NBW + j] = cmul(cmake(-tj.x, -tj.y), acc);
    } else if (tid == j) {
      Ts[j * QSP_BIG_NBW + j] = tj;
    }
    __syncthreads();
  }
  // ---- V T and V T^H ----
  cplx* VTp = w.VT + (long)p0 * ldS + p0;
  for (int idx = tid; idx < mm * jb; idx += nt) {
    const int j = idx / mm, r = idx - j * mm;
    cplx a1 = cmake(0.0, 0.0), a2 = cmake(0.0, 0.0);
    for (int l = 0; l < jb; ++l) {
      const cplx v = Vp[(long)l * ldS + r];
      if (l <= j) cfma(a1, v, Ts[l * QSP_BIG_NBW + j]);
      if (l >= j) cfma(a2, v, cconj(Ts[j * QSP_BIG_NBW + l]));
    }
    VTp[(long)j * ldS + r] = a1;
    w.VTh[(long)j * ldS + r] = a2;
  }
  if (tid == 0) {
    BigGemm g = z, h = z;
    const int ntr = n - (p0 + jb);
    if (ntr > 0) {
      cplx* C = w.S + (long)(p0 + jb) * ldS + p0;
      g.A = Vp, g.sAr = ldS, g.sAk = 1, g.conjA = 1;  // W = V^H C
      g.B = C, g.sBk = 1, g.sBc = ldS;
      g.C = w.Wt, g.sCr = w.ldw, g.sCc = 1;
      g.m = jb, g.n = ntr, g.k = mm;
      h.A = w.VTh, h.sAr = 1, h.sAk = ldS;            // C -= (V T^H) W
      h.B = w.Wt, h.sBk = w.ldw, h.sBc = 1;
      h.C = C, h.sCr = 1, h.sCc = ldS;
      h.m = mm, h.n = ntr, h.k = jb, h.mode = 2;
    }
    gW[inst] = g;
    gC[inst] = h;
  }
}

// operands of the backward accumulation of the thin Q for panel p0:  Q[p0:m, p0:K] -= (V T) (V^H Q[p0:m, p0:K])
__global__ void k_big_lc_qdesc(BigLCArgs P, const BigLCState* __restrict__ S, int p0, BigGemm* __restrict__ gW, BigGemm* __restrict__ gC) {
  const int inst = blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= P.B) return;
  BigGemm g, h;
  memset(&g, 0, sizeof(g));
  memset(&h, 0, sizeof(h));
  const BigLCState s = S[inst];
  if (!s.bad && p0 < s.K) {
    BigLCWork w = big_lc_work(P, inst);
    const int ldS = w.ldS, jb = imin(QSP_BIG_NBW, s.K - p0), mm = s.m - p0, nc = s.K - p0;
    cplx* Vp = w.V + (long)p0 * ldS + p0;
    cplx* VTp = w.VT + (long)p0 * ldS + p0;
    cplx* Qs = w.Q + (long)p0 * ldS + p0;
    g.A = Vp, g.sAr = ldS, g.sAk = 1, g.conjA = 1;
    g.B = Qs, g.sBk = 1, g.sBc = ldS;
    g.C = w.Wt, g.sCr = w.ldw, g.sCc = 1;
    g.m = jb, g.n = nc, g.k = mm;
    h.A = VTp, h.sAr = 1, h.sAk = ldS;
    h.B = w.Wt, h.sBk = w.ldw, h.sBc = 1;
    h.C = Qs, h.sCr = 1, h.sCc = ldS;
    h.m = mm, h.n = nc, h.k = jb, h.mode = 2;
  }
  gW[inst] = g;
  gC[inst] = h;
}

// ----------------------------------------------------------------------------------
// transfer-matrix overlap sum_x a(x) conj(b(x)) for large bonds (norm_mps_ovrlap, helper_routines.py:65-73): the two
// environment updates of a site as batched tensor-core GEMMs,
//   T[x,(p,y')] = sum_y E[x,y] conj(B[y,p,y']),     E'[x',y'] = sum_(x,p) A[x,p,x'] T[(x,p),y']
// ----------------------------------------------------------------------------------
struct BigOvArgs {
  const cplx* a;
  const int32_t* bonds_a;
  int Da;
  const cplx* b;
  const int32_t* bonds_b;
  int Db, B, L;
  cplx* result;
  cplx* work;  // per instance: E0, E1 (Da x Db), T (Da x 2 Db)
  long work_stride;
};
__host__ __device__ inline long big_ov_work_elems(int Da, int Db) { return 4L * Da * Db + 64; }

// gT[2 * inst + p], gE[inst]: operands of site i (ping-pong environments: E_i lives in slot i & 1)
__global__ void k_big_ov_site(BigOvArgs P, int i, BigGemm* __restrict__ gT, BigGemm* __restrict__ gE) {
  const int inst = blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= P.B) return;
  const int32_t* ba = P.bonds_a + (long)inst * (P.L + 1);
  const int32_t* bb = P.bonds_b + (long)inst * (P.L + 1);
  cplx* E0 = P.work + (long)inst * P.work_stride;
  cplx* E1 = E0 + (long)P.Da * P.Db;
  cplx* T = E1 + (long)P.Da * P.Db;
  cplx* E = (i & 1) ? E1 : E0;
  cplx* En = (i & 1) ? E0 : E1;
  if (i == 0) E[0] = cmake(1.0, 0.0);  // bonds[0] == 1 on both sides
  const int xa = ba[i], xb = bb[i], ya = ba[i + 1], yb = bb[i + 1];
  const cplx* Ai = P.a + ((long)inst * P.L + i) * ((long)P.Da * 2 * P.Da);
  const cplx* Bi = P.b + ((long)inst * P.L + i) * ((long)P.Db * 2 * P.Db);
  BigGemm g;
  for (int p = 0; p < 2; ++p) {
    memset(&g, 0, sizeof(g));
    g.A = E, g.sAr = xb, g.sAk = 1;
    g.B = Bi + (long)p * P.Db, g.sBk = 2L * P.Db, g.sBc = 1, g.conjB = 1;
    g.C = T + (long)p * yb, g.sCr = 2L * yb, g.sCc = 1;
    g.m = xa, g.n = yb, g.k = xb;
    gT[2 * inst + p] = g;
  }
  memset(&g, 0, sizeof(g));
  g.A = Ai, g.sAr = 1, g.sAk = P.Da;
  g.B = T, g.sBk = yb, g.sBc = 1;
  g.C = En, g.sCr = yb, g.sCc = 1;
  g.m = ya, g.n = yb, g.k = 2 * xa;
  gE[inst] = g;
  if (i == P.L - 1) {
    // nothing: the result is read by k_big_ov_done after the last GEMM
  }
}
__global__ void k_big_ov_done(BigOvArgs P) {
  const int inst = blockIdx.x * blockDim.x + threadIdx.x;
  if (inst >= P.B) return;
  const cplx* E0 = P.work + (long)inst * P.work_stride;
  const cplx* E1 = E0 + (long)P.Da * P.Db;
  P.result[inst] = (P.L & 1) ? E1[0] : E0[0];  // E_L lives in slot L & 1
}
