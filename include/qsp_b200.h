/* qsp_b200 -- C ABI of the B200-native sequential MPS-encoding path.
 *
 * Drop-in boundary for mohsin-0/qsp's `MPSPreparation.sequential_unitary_circuit`
 * (qsp/tsp.py:79-149 -> qsp/sequential/sequential.py:19-102 -> qsp/sequential/core.py and
 * qsp/tsp_helper_routines/helper_routines.py:29-73).  The reference has no FFI layer of its
 * own (it is pure Python on top of quimb); these entry points are what a binding for this
 * path would call -- INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (torch tensors in the python
 *    host); nothing is allocated behind the caller's back: scratch comes from `work`,
 *    sized by the matching *_workspace_bytes() query;
 *  - calls are asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant per
 *    stream and per host thread (the library keeps no mutable process-wide state besides an atomic
 *    launch tally; the last-error string and the optional profile buffer are per host thread), one
 *    device per call;
 *  - complex128 = two doubles (re, im);
 *  - a batch of B independent MPS of L sites is stored as mps[B][L][D][2][D]: element
 *    (l, p, r) of a site at (l*2+p)*D + r, D = padded bond; the live extent of bond j of
 *    instance b is bonds[b][j] (int32, j = 0..L, bonds[.][0] = bonds[.][L] = 1).  Nothing
 *    outside the live extent is read; padding is never initialised;
 *  - gates: G[..][L][4][4] row-major, exactly the matrices the reference keeps in
 *    Gs[i].data (core.py:46,71-73,105-107); a one-qubit gate sits in the top-left 2x2
 *    block.  kind[..][L]: 2 = two-qubit gate on (it, it+1), 1 = one-qubit gate on it (the
 *    site closes a segment, core.py:243-270);
 *  - status[B] is OR-ed with QSP_ST_* bits; the caller zeroes it.
 *  - return value: 0 on success, a QSP_ERR_* code otherwise (nothing was launched when
 *    the code is QSP_ERR_UNSUPPORTED / QSP_ERR_WORKSPACE / QSP_ERR_ARG).
 */
#ifndef QSP_B200_H
#define QSP_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  double re, im;
} qsp_c128;

#define QSP_OK 0
#define QSP_ERR_ARG 1         /* null pointer / non-positive size */
#define QSP_ERR_UNSUPPORTED 2 /* shape not supported by this build (no bond-dimension limit since the batched large-bond path) */
#define QSP_ERR_WORKSPACE 3   /* work_bytes smaller than the query says */
#define QSP_ERR_CUDA 4        /* a launch failed; see qsp_last_cuda_error() */

#define QSP_ST_JACOBI_NOCONV 1
#define QSP_ST_BOND_OVERFLOW 2
#define QSP_ST_NONUNITARY 4
#define QSP_ST_DEGENERATE 8
#define QSP_ST_BADNORM 16

int qsp_version(void);
/* 1 when the library was built for a CUDA device, 0 for the CPU emulation used by the
 * test-suite (tests/emu) -- the python package refuses to run on the latter. */
int qsp_is_device_build(void);
const char* qsp_last_cuda_error(void);
/* number of kernel launches issued by this library since load (bench.py: gpu_launches) */
long long qsp_launch_count(void);

/* quimb MPS.left_canonize() in place (first half of `compress`, tsp.py:72, core.py:17,
 * sequential.py:54,57).  R has a non-negative real diagonal (quimb's stabilised QR).
 * Bonds whose (2D x D) site matrix fits shared memory (D <= ~85) need no workspace (the query
 * returns 0, work may be NULL); larger bonds (any size) take the batched blocked-Householder path
 * (compact-WY panels, trailing updates and the thin Q as tensor-core GEMMs) out of the workspace. */
size_t qsp_left_canonize_workspace_bytes(int B, int L, int D);
int qsp_left_canonize(qsp_c128* mps, int32_t* bonds, int B, int L, int D, int32_t* status, void* work,
                      size_t work_bytes, void* stream);

/* quimb MPS.right_compress(max_bond, cutoff, cutoff_mode='rel') on a LEFT-canonical input:
 * right-to-left truncating SVD sweep (second half of `compress`).  max_bond = 0: cutoff
 * only.  normalize != 0: the orthogonality centre (site 0) is divided by its norm
 * (right_canonize(normalize=True), sequential.py:53,56).  svals (optional): the nsv
 * largest singular values seen at each bond, svals[b][i-1][0..nsv) for the bond between
 * sites i-1 and i.  amp0 (optional): <0...0|out>/||out||, i.e. norm_mps_ovrlap(out,
 * cl_zero_mps(L)) (helper_routines.py:65-73, sequential.py:60).
 * Site matrices up to 128 x 128 (bonds <= 64) stay on one SM for the whole sweep (register-resident Jacobi SVD); larger ones
 * take batched launches per site: block one-sided Jacobi SVD on the FP64 tensor cores (DMMA) with TMA-staged panels and
 * DMMA GEMMs -- that path reads one convergence flag per Jacobi sweep on the host (it synchronises `stream`). */
size_t qsp_right_compress_workspace_bytes(int B, int L, int D_in, int D_out);
int qsp_right_compress(const qsp_c128* in, const int32_t* bonds_in, int D_in, qsp_c128* out, int32_t* bonds_out,
                       int D_out, int B, int L, int max_bond, double cutoff, int normalize, double* svals, int nsv,
                       qsp_c128* amp0, int32_t* status, void* work, size_t work_bytes, void* stream);

/* generate_unitaries (core.py:25-123) on the D<=2 right-canonical, normalised MPS
 * d2[B][L][2][2][2] (gauge-fixed in place), with the deterministic 'canonical' completion
 *  in place of scipy.linalg.null_space. */
int qsp_generate_unitaries(qsp_c128* d2, const int32_t* d2_bonds, int B, int L, qsp_c128* G, int32_t* kind,
                           int32_t* status, void* stream);

/* apply_inverse_unitary_layer_on_wfn (core.py:224-240) on a LEFT-canonical MPS, exact
 * (bonds double inside a segment, D_out >= 2*max bond); the result is left canonical. */
int qsp_apply_inverse_layer(const qsp_c128* in, const int32_t* bonds_in, int D_in, const qsp_c128* G,
                            const int32_t* kind, qsp_c128* out, int32_t* bonds_out, int D_out, int B, int L,
                            int32_t* status, void* stream);
/* Fused gate-apply-and-split: apply_inverse_unitary_layer_on_wfn (core.py:224-240) followed by
 * [right_canonize; compress('right', max_bond)] (sequential.py:52-54) in one right-to-left sweep over
 * a LEFT-canonical psi (D-padded in and out, out != in).  The doubled tensors of U^dagger psi are never
 * stored: site t of U^dagger psi is G_{t-1}^dagger o A_{t-1}, so the sweep forms M = A_{t-1} C, mixes it
 * with the 4x4 gate and splits it with the Jacobi SVD.  amp0 as in qsp_right_compress. */
size_t qsp_apply_inverse_and_compress_workspace_bytes(int B, int L, int D);
int qsp_apply_inverse_and_compress(const qsp_c128* in, const int32_t* bonds_in, const qsp_c128* G, const int32_t* kind,
                                   qsp_c128* out, int32_t* bonds_out, int D, int B, int L, int max_bond, double cutoff,
                                   int normalize, qsp_c128* amp0, int32_t* status, void* work, size_t work_bytes,
                                   void* stream);

/* lcu/lcu.py:75,89-92: out = alpha |x> + beta |y>, the direct sum of the bond spaces
 * (quimb MatrixProductState.__add__ / __sub__ / scalar __mul__); alpha[B], beta[B] complex, D_out >= live
 * bond of x + live bond of y, out distinct from x and y.  qsp_mps_scale: |x> <- alpha |x> in place. */
int qsp_mps_axpby(const qsp_c128* x, const int32_t* bonds_x, int D_x, const qsp_c128* y, const int32_t* bonds_y, int D_y,
                  const qsp_c128* alpha, const qsp_c128* beta, qsp_c128* out, int32_t* bonds_out, int D_out, int B, int L,
                  int32_t* status, void* stream);
int qsp_mps_scale(qsp_c128* x, const int32_t* bonds_x, int D_x, const qsp_c128* alpha, int B, int L, void* stream);

/* the gate application of apply_unitary_layer_on_wfn (core.py:126-138) on a
 * RIGHT-canonical MPS, exact; the `compress('right')` of core.py:140 is
 * qsp_left_canonize + qsp_right_compress on the result. */
int qsp_apply_forward_layer(const qsp_c128* in, const int32_t* bonds_in, int D_in, const qsp_c128* G,
                            const int32_t* kind, qsp_c128* out, int32_t* bonds_out, int D_out, int B, int L,
                            int32_t* status, void* stream);

/* sum_x a(x) conj(b(x)) -- the numerator of norm_mps_ovrlap(a, b) (helper_routines.py:69). */
size_t qsp_overlap_workspace_bytes(int B, int L, int D_a, int D_b);
int qsp_overlap(const qsp_c128* a, const int32_t* bonds_a, int D_a, const qsp_c128* b, const int32_t* bonds_b,
                int D_b, int B, int L, qsp_c128* result, void* work, size_t work_bytes, void* stream);

/* MPSPreparation.__init__ + the first lines of sequential_unitary_circuit (tsp.py:63-77,
 * sequential.py:38-40): normalise, compress('right'), right_canonize(normalize=True).
 * In place on mps/bonds; `tmp` has the size of mps. */
size_t qsp_prepare_target_workspace_bytes(int B, int L, int D);
int qsp_prepare_target(qsp_c128* mps, int32_t* bonds, int B, int L, int D, double cutoff, int32_t* status,
                       void* work, size_t work_bytes, void* stream);

/* apply_unitary_layers_on_wfn(unitaries, |0..0>) + right_canonize(normalize) + norm_mps_ovrlap with the
 * target (core.py:217-221, sequential.py:67-70, tsp.py:130-132): result[b] = <target_b| U_0 ... U_{n-1} |0..0>
 * for G[n][B][L][4][4], kind[n][B][L].  Every forward layer is applied exactly (bonds double); the
 * reference's cutoff-only compress('right') after each layer does not change the state and is only run
 * when the doubled bonds would not fit (more than 8 layers on chains of at most 15 sites). */
size_t qsp_encoded_overlap_workspace_bytes(int B, int L, int D, int layers);
int qsp_encoded_overlap(const qsp_c128* G, const int32_t* kind, int n, const qsp_c128* target,
                        const int32_t* target_bonds, int B, int L, int D, double cutoff, qsp_c128* result,
                        int32_t* status, void* work, size_t work_bytes, void* stream);

/* Statevector of the emitted circuit on the device -- the cross-check the reference runs on qiskit-aer for small
 * registers (qsp/q_circs/qiskit_circuit.py:52-67), batched.  state[B][2^L] complex128, axis i of the (2,)*L tensor =
 * qubit i.  qsp_circuit_statevector: |psi_b> = U_0 .. U_{n-1} |0..0> from the layer-major outputs of qsp_seq_prepare
 * (layers in reverse order, sites ascending, a two-qubit gate of site it on the qiskit qubit list [it+1, it], a
 * one-qubit gate on [it]: qiskit_circuit.py:38-48).  qsp_statevector_apply: one gate on qubits [q_lo, q_hi] (matrix
 * index 2*bit(q_hi) + bit(q_lo); q_lo = -1: one-qubit gate on q_hi, top-left 2x2 block) on every instance; gates is one
 * shared 4x4 (per_instance = 0) or [B][4][4]; kind (optional, [B]): 0 skip / 1 one-qubit on q_hi / 2 two-qubit. */
int qsp_statevector_apply(qsp_c128* state, int B, int L, const qsp_c128* gates, int per_instance, const int32_t* kind, int q_hi,
                          int q_lo, void* stream);
int qsp_circuit_statevector(const qsp_c128* G, const int32_t* kind, int n, int B, int L, qsp_c128* state, void* stream);

/* The whole layer loop of sequential_unitary_circuit (sequential.py:42-85) for a batch.
 *   target / target_bonds : right-canonical normalised MPS (qsp_prepare_target output), D-padded;
 *                           every bond must stay <= D at every layer start (with
 *                           max_bond = 0 the caller pads for the doubling: D >= min(chi*2^layers, 2^(L/2)))
 *   max_bond              : do_compression ? max_bond_dim : 0
 *   flags                 : QSP_FLAG_*
 * Outputs (layer-major so that every layer is one contiguous batch slab):
 *   overlaps[layers][B]      encoded-state overlaps <target|U_0..U_it|0> at it % overlaps_gap == 0
 *                            and at the last layer (sequential.py:67-70); NaN elsewhere
 *   dis_overlaps[layers][B]  norm_mps_ovrlap(disentangled mps, |0..0>)  (sequential.py:60)
 *   G[layers][B][L][4][4], kind[layers][B][L]
 *   d2_svals[layers][B][L-1][2]  the two largest singular values at every bond of the D=2 truncation
 *   final_mps / final_bonds  (optional) the disentangled MPS after the last layer, D-padded
 */
#define QSP_FLAG_SECOND_PASS 1 /* also run the trailing cutoff-only compress() of sequential.py:56-57 */
#define QSP_FLAG_NO_ENCODED 2  /* skip the encoded-state overlaps (overlaps[] stays NaN) */
size_t qsp_seq_prepare_workspace_bytes(int B, int L, int D, int layers, int max_bond, int flags);
int qsp_seq_prepare(const qsp_c128* target, const int32_t* target_bonds, int B, int L, int D, int layers,
                    int max_bond, double cutoff, int overlaps_gap, int flags, qsp_c128* overlaps,
                    qsp_c128* dis_overlaps, qsp_c128* G, int32_t* kind, double* d2_svals, qsp_c128* final_mps,
                    int32_t* final_bonds, int32_t* status, void* work, size_t work_bytes, void* stream);

/* Measurement aid (state of the calling host thread): when set (device pointer to >= 8*grid int64, zeroed by the
 * caller), every qsp_right_compress launch issued by this thread adds per-CTA clock64 cycles: [0] A*C GEMM, [1] staging, [2] Jacobi,
 * [3] ranking, [4] Vh/U s, [5] Jacobi sweeps, [6] sites, [7] floating-point operations executed (Jacobi visits + GEMMs).  NULL (default) disables it. */
void qsp_set_profile_buffer(void* buf);

#ifdef __cplusplus
}
#endif
#endif
