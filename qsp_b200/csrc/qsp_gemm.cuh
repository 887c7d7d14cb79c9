// CTA-level complex128 GEMM on generic (strided) operands.
//
//   C[i*sCr + j*sCc]  (+)=  sum_k  opA(A[i*sAr + k*sAk]) * opB(B[k*sBk + j*sBc])
//
// All matrices on the hot path are small (<= 256 x 256) and live in L2 / shared memory,
// so one CTA multiplies one instance's operands: operand tiles are staged through shared
// memory (k-major, padded against bank conflicts), every thread owns a 4x4 register
// micro-tile and issues 64 DFMA per 8 LDS.128.  tcgen05 has no FP64 kind; on B200 the
// FP64 DMMA peak equals the DFMA peak, so the register-tiled DFMA loop is the right tool
// ("why no tensor cores").
#pragma once
#include "qsp_common.cuh"

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
__device__ inline void cta_zgemm_deep(int m, int n, int kdim, const cplx* __restrict__ A, long sAr, long sAk, bool conjA,
                                      const cplx* __restrict__ B, long sBk, long sBc, bool conjB, cplx* __restrict__ C,
                                      long sCr, long sCc, bool accumulate, cplx* tiles) {
  switch (blockDim.x) {
    case 512: cta_zgemm_t<512, QSP_GEMM_TK_DEEP>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles); return;
    case 256: cta_zgemm_t<256, QSP_GEMM_TK_DEEP>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles); return;
    case 128: cta_zgemm_t<128, QSP_GEMM_TK_DEEP>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles); return;
    default: cta_zgemm_t<64, QSP_GEMM_TK_DEEP>(m, n, kdim, A, sAr, sAk, conjA, B, sBk, sBc, conjB, C, sCr, sCc, accumulate, tiles); return;
  }
}
