/* qsp_b200_dev -- development / measurement entry points of libqsp_b200_dev.so.
 *
 * NOT part of the drop-in boundary (include/qsp_b200.h).  libqsp_b200_dev.so is the product ABI compiled
 * with its development hooks plus the probes below; bench.py takes the FP64 roofline denominators from it,
 * scripts/ use the micro-benchmarks, and the CPU test emulation (tests/emu) is built from the same source
 * (qsp_b200/csrc/qsp_dev.cu).  Nothing in the product package's hot path loads it.
 */
#ifndef QSP_B200_DEV_H
#define QSP_B200_DEV_H
#include "qsp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Development aid: Jacobi micro-benchmark, W[B][nrows][ldw] orthogonalised in place (ldw a
 * multiple of 32), nt threads per CTA. */
int qsp_jacobi_bench_fat(qsp_c128* W, int nrows, int ldw, int B, int* sweeps, int* ok, void* stream);
int qsp_smem_probe(double* out, int blocks, int iters, int mode, void* stream); /* shared-memory bandwidth probe */
int qsp_tmem_probe(double* out, int blocks, int iters, int mode, void* stream); /* tensor-memory (tcgen05.ld/st) bandwidth probe */
void qsp_jacobi_variant(int bits); /* timing experiments: disables parts of the Jacobi round (results invalid) */
void qsp_debug_flags(int flags);    /* test hook: bit 0 = take the large-bond left-canonise path at any size */
int qsp_jacobi_bench(qsp_c128* W, int nrows, int ndotp, int ldw, int B, int nt, int* sweeps, int* ok, void* stream);

/* Measurement aid: blocks x threads threads each run `iters` rounds of 16 independent
 * double-precision FMAs (flops = blocks*threads*iters*32).  bench.py times it to obtain the
 * FP64 roofline denominator, which MEASURED_PEAKS.json does not carry. */
int qsp_dfma_probe(double* out, int blocks, int threads, int iters, void* stream);
/* Same for the FP64 tensor pipe: `iters` rounds of 8 independent warp-level DMMA (mma.sync m8n8k4.f64),
 * flops = blocks*(threads/32)*iters*8*512. */
int qsp_dmma_probe(double* out, int blocks, int threads, int iters, void* stream);

/* Large-bond path pieces on their own (tests, micro-benchmarks, ncu captures).
 * qsp_big_jacobi: block one-sided Jacobi SVD on the FP64 tensor cores; the rows of W[B][nrows][ldw] are orthogonalised
 * in place (ldw and ndotp multiples of 32, zero padded); sweeps[B], ok[B] (device); work: qsp_big_jacobi_work_bytes().
 * qsp_big_gemm: batched DMMA GEMM C[b] = op(A[b]) op(B[b]) (row-major, trans / conj flags); work: batch * 128 bytes. */
void qsp_big_tuning(int update_splits, int gram_splits); /* tuning hook: column splits of the block-Jacobi launches (0 = automatic) */
size_t qsp_big_jacobi_work_bytes(int nrows, int ldw, int B);
int qsp_big_jacobi(qsp_c128* W, int nrows, int ndotp, int ldw, int B, int* sweeps, int* ok, void* work, void* stream);
int qsp_big_gemm(const qsp_c128* A, int transA, int conjA, const qsp_c128* Bm, int transB, int conjB, qsp_c128* C, int m, int n,
                 int k, int batch, void* work, void* stream);

#ifdef __cplusplus
}
#endif
#endif
