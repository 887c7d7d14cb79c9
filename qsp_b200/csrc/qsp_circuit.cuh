// Statevector simulation of the emitted circuits on the device: the cross-check the reference runs on qiskit-aer for
// small registers (qsp/q_circs/qiskit_circuit.py:52-67), batched over instances and extended to L <= 24.
//
// state[B][2^L] complex128; axis i of the (2,)*L tensor is qubit i, i.e. qubit i is bit (L-1-i) of the flat index.
// A two-qubit gate on the qiskit qubit list [q0, q1] has matrix index 2*bit(q1) + bit(q0)
// (qiskit_circuit.py:45: `circ.unitary(G, [it+1, it])`); a one-qubit gate sits in the top-left 2x2 block.
#pragma once
#include "qsp_common.cuh"

// gates: instance b uses the 4x4 at gates + b * gstride (gstride = 0: one shared gate); kind (optional): instance b reads
// kind[b * kstride]: 1 = one-qubit gate on q_hi, 2 = two-qubit gate on [q_lo, q_hi], 0 = skip.  One thread per group of
// four amplitudes (a one-qubit gate handles them as two pairs).
__global__ void __launch_bounds__(256) k_sv_gate(cplx* __restrict__ state, int B, int L, const cplx* __restrict__ gates, long gstride,
                                                 const int32_t* __restrict__ kind, long kstride, int q_hi, int q_lo) {
  const long groups = 1L << (L - 2 > 0 ? L - 2 : 0);
  const long tot = (long)B * groups;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += (long)gridDim.x * blockDim.x) {
    const int b = (int)(t / groups);
    const long grp = t - (long)b * groups;
    const int kd = kind ? kind[(long)b * kstride] : (q_lo < 0 ? 1 : 2);
    if (kd == 0 || (kd == 2 && q_lo < 0)) continue;
    const cplx* g = gates + (long)b * gstride;
    cplx* psi = state + ((long)b << L);
    if (L == 1) {
      if (grp == 0) {
        const cplx a0 = psi[0], a1 = psi[1];
        psi[0] = cadd(cmul(g[0], a0), cmul(g[1], a1));
        psi[1] = cadd(cmul(g[4], a0), cmul(g[5], a1));
      }
      continue;
    }
    // second bit position: the partner qubit of a two-qubit gate, or any other qubit for a one-qubit gate (the four
    // amplitudes of the group are then two independent pairs)
    const int pa = L - 1 - q_hi;
    int pb = (kd == 2) ? L - 1 - q_lo : (pa == 0 ? 1 : 0);
    const int p0 = pa < pb ? pa : pb, p1 = pa < pb ? pb : pa;
    long base = grp;
    base = ((base >> p0) << (p0 + 1)) | (base & ((1L << p0) - 1));
    base = ((base >> p1) << (p1 + 1)) | (base & ((1L << p1) - 1));
    const long sa = 1L << pa, sb = 1L << pb;
    cplx v[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) v[m] = psi[base + ((m >> 1) ? sa : 0) + ((m & 1) ? sb : 0)];
    if (kd == 2) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        cplx acc = cmake(0.0, 0.0);
#pragma unroll
        for (int m = 0; m < 4; ++m) cfma(acc, g[r * 4 + m], v[m]);
        psi[base + ((r >> 1) ? sa : 0) + ((r & 1) ? sb : 0)] = acc;
      }
    } else {
#pragma unroll
      for (int y = 0; y < 2; ++y) {
        const cplx a0 = v[y], a1 = v[2 + y];
        psi[base + (y ? sb : 0)] = cadd(cmul(g[0], a0), cmul(g[1], a1));
        psi[base + sa + (y ? sb : 0)] = cadd(cmul(g[4], a0), cmul(g[5], a1));
      }
    }
  }
}

__global__ void k_sv_init(cplx* __restrict__ state, int B, int L) {
  const long n = (long)B << L;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x)
    state[t] = cmake((t & ((1L << L) - 1)) == 0 ? 1.0 : 0.0, 0.0);
}
