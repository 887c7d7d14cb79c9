// CTA-level complex128 GEMM on generic (strided) operands.
//
//   C[i*sCr + j*sCc]  (+)=  sum_k  opA(A[i*sAr + k*sAk]) * opB(B[k*sBk + j*sBc])
//
// The matrices of the small-bond sweeps (<= 128 x 128) live in L2 / shared memory, so one CTA multiplies one
// instance's operands.  Two variants: cta_zgemm_t (DFMA: k-major padded smem tiles, 4x4 register micro-tile per
// thread, 64 DFMA per 8 LDS.128) and cta_zgemm_dmma_t (FP64 tensor cores, 8x8 complex tiles as four real DMMA
// m8n8k4; used by the 256 / 512-thread sweeps).  tcgen05 has no FP64 kind and the DMMA peak equals the DFMA peak on
// B200, so DMMA buys issue slots and operand bandwidth, not a higher ceiling.
#pragma once
#include "qsp_common.cuh"
#include "qsp_mma.cuh"

#define QSP_GEMM_TK 8
#define QSP_GEMM_TN 64

// bytes of shared memory cta_zgemm needs for a block of nt threads
__host__ __device__ inline size_t qsp_gemm_smem_bytes(int nt, int tk = QSP_GEMM_TK) {
  int tm = 4 * (nt / 16);
  return (size_t)tk * ((tm + 1) + (QSP_GEMM_TN + 1)) * sizeof(cplx);
}

// Must be called by all threads of the block (contains __syncthreads).
// NT = blockDim.x at compile time (tile geometry and shared strides become immediates).
template <int NT, int TK = QSP_GEMM_TK>
__device__ __forceinline__ void cta_zgemm_t(int m, int n, int kdim, const cplx* __restrict__ A, long sAr, long sAk,
                                            bool conjA, const cplx* __restrict__ B, long sBk, long sBc, bool conjB,
                                            cplx* __restrict__ C, long sCr, long sCc, bool accumulate, cplx* tiles) {
  constexpr int TM = 4 * (NT / 16), TN = QSP_GEMM_TN;
  constexpr int ldA = TM + 1, ldB = TN + 1;
  const int tid = threadIdx.x;
  cplx* As = tiles;             // [TK][ldA]
  cplx* Bs = tiles + TK * ldA;  // [TK][ldB]
  const int tx = tid & 15, ty = tid >> 4;
  const bool unitA = (sAk == 1), unitB = (sBk == 1);

  for (int i0 = 0; i0 < m; i0 += TM) {
    for (int j0 = 0; j0 < n; j0 += TN) {
      cplx acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = cmake(0.0, 0.0);
      const bool active = (i0 + ty * 4 < m) && (j0 + tx * 4 < n);

      for (int k0 = 0; k0 < kdim; k0 += TK) {
        __syncthreads();
#pragma unroll
        for (int idx = tid; idx < TM * TK; idx += NT) {
          int r, kk;
          if (unitA) {
            kk = idx % TK;
            r = idx / TK;
          } else {
            r = idx % TM;
            kk = idx / TM;
          }
          cplx v = cmake(0.0, 0.0);
          if (i0 + r < m && k0 + kk < kdim) {
            v = A[(long)(i0 + r) * sAr + (long)(k0 + kk) * sAk];
            if (conjA) v.y = -v.y;
          }
          As[kk * ldA + r] = v;
        }
#pragma unroll
        for (int idx = tid; idx < TN * TK; idx += NT) {
          int c, kk;
          if (unitB) {
            kk = idx % TK;
            c = idx / TK;
          } else {
            c = idx % TN;
            kk = idx / TN;
          }
          cplx v = cmake(0.0, 0.0);
          if (j0 + c < n && k0 + kk < kdim) {
            v = B[(long)(k0 + kk) * sBk + (long)(j0 + c) * sBc];
            if (conjB) v.y = -v.y;
          }
          Bs[kk * ldB + c] = v;
        }
        __syncthreads();
        if (active) {
#pragma unroll
          for (int kk = 0; kk < TK; ++kk) {
            cplx a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk * ldA + ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk * ldB + tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) cfma(acc[i][j], a[i], b[j]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int gi = i0 + ty * 4 + i;
        if (gi >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int gj = j0 + tx * 4 + j;
          if (gj >= n) continue;
          cplx* dst = &C[(long)gi * sCr + (long)gj * sCc];
          if (accumulate) {
            cplx o = *dst;
            *dst = cadd(o, acc[i][j]);
          } else {
            *dst = acc[i][j];
          }
        }
      }
    }
  }
  __syncthreads();
}

__device__ inline void cta_zgemm(int m, int n, int kdim, const cplx* __restrict__ A, long sAr, long sAk, bool conjA,
                                 const cplx* __restrict__ B, long sBk, long sBc, bool conjB, cplx* __restrict__ C,
                                 long sCr, long sCc, bool accumulate, cplx* tiles) {
  switch (blockDim.x) {
    case 512: cta_zgemm_t<512>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles); return;
    case 256: cta_zgemm_t<256>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles); return;
    case 128: cta_zgemm_t<128>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles); return;
    default: cta_zgemm_t<64>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles); return;
  }
}

// Deeper k-tiles (half the barriers and exposed load latencies per flop) for kernels that can
// spare the shared memory: qsp_gemm_smem_bytes(nt, QSP_GEMM_TK_DEEP).
#ifndef QSP_GEMM_TK_DEEP
#define QSP_GEMM_TK_DEEP 16
#endif
// Tensor-core variant for the 256- and 512-thread launches (chi >= 24): the same strided operands, 8 x 8 complex output
// tiles as four real DMMA m8n8k4 each (qsp_mma.cuh), warp tile 16 x 32, CTA tile (16 NT/64) x 64 -- 128 x 64 at 512
// threads, the shape of the A C and M Vh^H products of the chi = 64 sweep.  Operand tiles: As[row][k], Bs[col][k] with a
// stride of 12 (two rows of a quarter-warp's 128-bit loads hit distinct banks); needs (TM + 64) * 12 elements of `tiles`
// (less than the DFMA variant's k-major tiles).
template <int NT>
__device__ __forceinline__ void cta_zgemm_dmma_t(int m, int n, int kdim, const cplx* __restrict__ A, long sAr, long sAk, bool conjA,
                                                 const cplx* __restrict__ B, long sBk, long sBc, bool conjB, cplx* __restrict__ C,
                                                 long sCr, long sCc, bool accumulate, cplx* tiles) {
  constexpr int TM = 16 * (NT / 64), TN = 64, TK = 8, LD = 12;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  cplx* As = tiles;            // [TM][LD]
  cplx* Bs = tiles + TM * LD;  // [TN][LD]
  const int wm = (warp >> 1) * 16, wn = (warp & 1) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  const bool kfastA = (sAk == 1), kfastB = (sBk == 1);
  for (int i0 = 0; i0 < m; i0 += TM) {
    for (int j0 = 0; j0 < n; j0 += TN) {
      ctile acc[2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = ctile_zero();
      const bool active = (i0 + wm < m) && (j0 + wn < n);  // warp-uniform
      for (int k0 = 0; k0 < kdim; k0 += TK) {
        __syncthreads();
        for (int idx = tid; idx < TM * TK; idx += NT) {
          int r, kk;
          if (kfastA) {
            kk = idx % TK;
            r = idx / TK;
          } else {
            r = idx % TM;
            kk = idx / TM;
          }
          cplx v = cmake(0.0, 0.0);
          if (i0 + r < m && k0 + kk < kdim) {
            v = A[(long)(i0 + r) * sAr + (long)(k0 + kk) * sAk];
            if (conjA) v.y = -v.y;
          }
          As[r * LD + kk] = v;
        }
        for (int idx = tid; idx < TN * TK; idx += NT) {
          int c, kk;
          if (kfastB) {
            kk = idx % TK;
            c = idx / TK;
          } else {
            c = idx % TN;
            kk = idx / TN;
          }
          cplx v = cmake(0.0, 0.0);
          if (j0 + c < n && k0 + kk < kdim) {
            v = B[(long)(k0 + kk) * sBk + (long)(j0 + c) * sBc];
            if (conjB) v.y = -v.y;
          }
          Bs[c * LD + kk] = v;
        }
        __syncthreads();
        if (active) {
#pragma unroll
          for (int ks = 0; ks < TK; ks += 4) {
            cplx af[2], bf[4];
#pragma unroll
            for (int a = 0; a < 2; ++a) af[a] = As[(wm + a * 8 + fr) * LD + ks + fk];
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = Bs[(wn + b * 8 + fr) * LD + ks + fk];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) cdmma(acc[a][b], af[a].x, af[a].y, bf[b].x, bf[b].y);
          }
        }
      }
      if (active) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int gi = i0 + wm + a * 8 + fr;
          if (gi >= m) continue;
#pragma unroll
          for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int gj = j0 + wn + b * 8 + 2 * fk + e;
              if (gj >= n) continue;
              cplx v = cmake(e ? acc[a][b].r1 : acc[a][b].r0, e ? acc[a][b].i1 : acc[a][b].i0);
              cplx* dst = &C[(long)gi * sCr + (long)gj * sCc];
              *dst = accumulate ? cadd(*dst, v) : v;
            }
        }
      }
    }
  }
  __syncthreads();
}

__device__ inline void cta_zgemm_deep(int m, int n, int kdim, const cplx* __restrict__ A, long sAr, long sAk, bool conjA,
                                      const cplx* __restrict__ B, long sBk, long sBc, bool conjB, cplx* __restrict__ C,
                                      long sCr, long sCc, bool accumulate, cplx* tiles) {
#ifndef QSP_GEMM_NO_DMMA
  if (blockDim.x == 512) {
    cta_zgemm_dmma_t<512>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles);
    return;
  }
  if (blockDim.x == 256) {
    cta_zgemm_dmma_t<256>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles);
    return;
  }
#endif
  switch (blockDim.x) {
    case 512: cta_zgemm_t<512, QSP_GEMM_TK_DEEP>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles); return;
    case 256: cta_zgemm_t<256, QSP_GEMM_TK_DEEP>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles); return;
    case 128: cta_zgemm_t<128, QSP_GEMM_TK_DEEP>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles); return;
    default: cta_zgemm_t<64, QSP_GEMM_TK_DEEP>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles); return;
  }
}
