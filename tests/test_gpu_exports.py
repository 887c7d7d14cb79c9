"""Every public name of the drop-in packages (qsp_b200.sequential.__all__, qsp_b200.tsp_helper_routines.__all__)
called on the GPU and compared with the oracle's same-named function (qsp/sequential/core.py:12-22, 217-240,
qsp/tsp_helper_routines/helper_routines.py:29-73), the input pipeline (tsp.py:63-77) at chi = 16 / 64, and the
BASELINE.json configurations at their stated per-instance sizes."""
import numpy as np
import pytest

from helpers import gates_to_layers, pack_batch, unpack_batch
from oracle import qsp_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def drv():
    import torch

    assert torch.cuda.is_available(), "run with -m gpu on a CUDA box"
    from helpers import CudaDriver

    d = CudaDriver()
    assert d.lib.is_device_build()
    return d


def _prepared(mps):
    t = O.prepare_target(mps, "lpr")
    O.right_canonize(t, normalize=True)
    return t


def _plain(gs_lst):
    """reference-style Gs_lst (GateTensor entries) -> the oracle's [[start, end, [ndarray...]], ...]"""
    return [[seg[0], seg[1], [np.asarray(getattr(g, "data", g)) for g in seg[2]]] for seg in gs_lst]


def _fid(a, b):
    return abs(O.inner(a, b)) / np.sqrt(O.inner(a, a).real * O.inner(b, b).real)


# ---------------------------------------------------------------------------------------
# qsp_b200.sequential / qsp_b200.tsp_helper_routines: every exported name
# ---------------------------------------------------------------------------------------
def test_all_exports_are_covered_here():
    import qsp_b200.sequential as S
    import qsp_b200.tsp_helper_routines as H

    assert set(S.__all__) == {"sequential_unitary_circuit", "generate_bond_d_unitary", "apply_unitary_layers_on_wfn",
                              "apply_inverse_unitary_layer_on_wfn"}
    assert set(H.__all__) == {"cl_zero_mps", "norm_mps_ovrlap", "unitaries_specs", "unitaries_sanity_check"}


@pytest.mark.parametrize("L,chi", [(10, 8), (12, 100)])
def test_generate_bond_d_unitary_mirror(drv, L, chi):
    """core.py:12-22; chi = 100 takes the large-bond kernels (needs the workspace the broken round-1 call omitted)."""
    from qsp_b200.mps import MPSState
    from qsp_b200.sequential import generate_bond_d_unitary
    from qsp_b200.tsp_helper_routines import unitaries_sanity_check, unitaries_specs

    psi = _prepared(O.random_mps(L, chi, 201))
    got = generate_bond_d_unitary(MPSState(O.copy_mps(psi)))
    ref = O.generate_bond_d_unitary(psi, completion="canonical")
    unitaries_sanity_check(got)
    assert unitaries_specs(got) == O.unitaries_specs(ref)
    assert [(s[0], s[1]) for s in got] == [(s[0], s[1]) for s in ref]
    for sg, sr in zip(_plain(got), ref):
        for x, y in zip(sg[2], sr[2]):
            assert np.abs(x - y).max() < 1e-10
    # gauge-invariant check: the layer prepares the D=2 truncation of psi
    enc = O.apply_unitary_layers_on_wfn([_plain(got)], O.cl_zero_mps(L))
    assert abs(_fid(enc, O.d2_truncate(psi)) - 1.0) < 1e-12


def test_apply_layer_mirrors(drv):
    """core.py:217-221 and 224-240 as states (the bond gauge of the reference's SVD splits is not defined)."""
    from qsp_b200.mps import MPSState, gates_to_gs_lst
    from qsp_b200.sequential import apply_inverse_unitary_layer_on_wfn, apply_unitary_layers_on_wfn
    from qsp_b200.tsp_helper_routines import cl_zero_mps, norm_mps_ovrlap
    from helpers import layers_to_gates

    L, chi = 10, 8
    psi = _prepared(O.random_mps(L, chi, 211))
    lay1 = O.generate_bond_d_unitary(psi, completion="canonical")
    psi1 = O.apply_inverse_unitary_layer_on_wfn(lay1, psi)
    O.compress(psi1, max_bond=chi)
    O.right_canonize(psi1, normalize=True)
    lay2 = O.generate_bond_d_unitary(psi1, completion="canonical")
    gs = [gates_to_gs_lst(*layers_to_gates(l, L)) for l in (lay1, lay2)]

    inv = apply_inverse_unitary_layer_on_wfn(gs[0], MPSState(O.copy_mps(psi)))
    ref = O.apply_inverse_unitary_layer_on_wfn(lay1, psi)
    assert abs(O.inner(inv.arrays, ref) / O.inner(ref, ref) - 1.0) < 1e-12  # same state incl. phase and norm
    assert max(inv.bond_sizes()) <= 2 * chi

    zero = cl_zero_mps(L)
    assert [a.shape for a in zero.arrays] == [a.shape for a in O.cl_zero_mps(L)]
    enc = apply_unitary_layers_on_wfn(gs, zero)
    ref = O.apply_unitary_layers_on_wfn([lay1, lay2], O.cl_zero_mps(L))
    assert abs(O.inner(enc.arrays, ref) / O.inner(ref, ref) - 1.0) < 1e-12
    assert enc.bond_sizes() == O.bonds_of(ref)[1:-1]  # compress('right') after every layer, core.py:140

    # helper_routines.py:65-73
    got = norm_mps_ovrlap(enc, MPSState(O.copy_mps(psi)))
    want = O.norm_mps_ovrlap(ref, psi)
    assert abs(got - want) < 1e-12
    unnorm = MPSState([a * 1.7 for a in psi])
    assert abs(norm_mps_ovrlap(unnorm, unnorm) - 1.0) < 1e-12


def test_sequential_unitary_circuit_mirror(drv):
    """sequential.py:19-102 called directly with an un-prepared MPSState: same dict as the oracle."""
    from qsp_b200.mps import MPSState
    from qsp_b200.sequential import sequential_unitary_circuit

    raw = O.random_mps(9, 6, 221)
    mps = MPSState(O.copy_mps(raw))
    d = sequential_unitary_circuit(mps, 5, do_compression=True, max_bond_dim=6, verbose=False)
    tgt = _prepared(raw)
    ref = O.sequential_unitary_circuit(tgt, 5, do_compression=True, max_bond_dim=6, completion="canonical")
    assert d["it_Ds"] == ref["it_Ds"] == [0, 4] and d["depths"] == ref["depths"] and d["gate_counts"] == ref["gate_counts"]
    assert np.abs(np.array(d["overlaps"]) - np.array(ref["overlaps"])).max() < 1e-10
    assert np.abs(np.array(d["disentangled_mps_overlaps"]) - np.array(ref["disentangled_mps_overlaps"])).max() < 1e-10
    assert d["degenerate_d2_cut"] is False
    # the input object is replaced by its right-canonical normalised form (sequential.py:38-39)
    assert abs(_fid(mps.arrays, tgt) - 1.0) < 1e-12
    x = mps.arrays[3].reshape(mps.arrays[3].shape[0], -1)
    assert np.abs(x @ x.conj().T - np.eye(x.shape[0])).max() < 1e-12


# ---------------------------------------------------------------------------------------
# input pipeline (MPSPreparation.__init__, tsp.py:63-77 + sequential.py:38-40)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("L,chi,B", [(20, 16, 3), (14, 64, 2), (12, 100, 2)])
def test_prepare_target_matches_oracle(drv, L, chi, B):
    raws = [O.random_mps(L, chi, 230 + s) for s in range(B)]
    # one un-normalised instance with rank-deficient (duplicated) bond directions: the cutoff must trim them
    raws[0] = [a * 3.0 for a in raws[0]]
    raws[0][L // 2][:, :, -1] = raws[0][L // 2][:, :, 0]
    a, b = pack_batch(raws, chi)
    a, b = drv.from_host(a), drv.from_host(b)
    st = drv.prepare_target(a, b)
    assert drv.to_host(st).max() == 0
    got = unpack_batch(drv.to_host(a), drv.to_host(b))
    for raw, g in zip(raws, got):
        ref = _prepared(raw)
        assert O.bonds_of(g) == O.bonds_of(ref)
        assert abs(O.inner(g, ref) - 1.0) < 1e-12 and abs(O.inner(g, g).real - 1.0) < 1e-12
        for i in range(1, L):  # right canonical
            x = g[i].reshape(g[i].shape[0], -1)
            assert np.abs(x @ x.conj().T - np.eye(x.shape[0])).max() < 1e-12


def test_prepare_target_flags_bad_input(drv):
    """a zero / non-finite target is reported (QSP_ST_BADNORM), never a crash; the batched facade keeps the flag"""
    import torch

    from qsp_b200.tsp import BatchedMPSPreparation

    raws = [O.random_mps(8, 4, 240 + s) for s in range(3)]
    raws[1] = [a * 0.0 for a in raws[1]]
    raws[2][3][0, 0, 0] = np.nan
    prep = BatchedMPSPreparation(raws, shape="lpr")
    st = prep.status.cpu().numpy()
    assert st[0] == 0 and st[1] & 16 and st[2] & 16
    _, res = prep.sequential_unitary_circuit(2, do_compression=True, max_bond_dim=4)
    torch.cuda.synchronize()
    assert int(res.status[0]) == 0 and int(res.status[1]) & 16 and int(res.status[2]) & 16
    with pytest.raises(RuntimeError):
        prep.sequential_unitary_circuit(2, do_compression=True, max_bond_dim=4, check=True)


# ---------------------------------------------------------------------------------------
# BASELINE.json configurations at their stated per-instance sizes
# ---------------------------------------------------------------------------------------
def _run(drv, tgts, layers, chi):
    a, b = pack_batch(tgts, chi)
    r = drv.seq_prepare(drv.from_host(a), drv.from_host(b), layers, max_bond=chi)
    r = {k: drv.to_host(v) for k, v in r.items()}
    assert r["status"].max() == 0, r["status"]
    return r


def _check_named_config(drv, L, chi, layers, n_inst, seed, fid_tol=1e-9):
    """at the full per-instance size: same-convention oracle run (per-layer overlaps 1e-10, fidelity of the final
    disentangled state, bond profile) and teacher-forced run (the device's unitaries through the oracle's literal
    steps: per-layer overlaps 1e-10, D=2 singular values of EVERY layer)."""
    tgts = [_prepared(O.random_mps(L, chi, seed + s)) for s in range(n_inst)]
    r = _run(drv, tgts, layers, chi)
    ev = [i for i in range(layers) if i % 4 == 0 or i == layers - 1]
    kw = dict(do_compression=True, max_bond_dim=chi)
    for k, tgt in enumerate(tgts):
        sc = O.sequential_unitary_circuit(tgt, layers, completion="canonical", **kw)
        assert np.abs(np.array(sc["disentangled_mps_overlaps"]) - r["dis"][:, k]).max() < 1e-10
        assert np.abs(np.array(sc["overlaps"]) - r["overlaps"][ev, k]).max() < 1e-10
        fin = unpack_batch(r["final"], r["final_bonds"])[k]
        assert O.bonds_of(fin) == O.bonds_of(sc["final_mps"])
        ov = O.inner(fin, sc["final_mps"])
        # the same state to rounding; its global phase (fixed by the unitaries' conventions) may drift by ~1e-9 per layer
        assert abs(abs(ov) - 1.0) < 1e-11 and abs(np.angle(ov)) < fid_tol
        assert abs(O.inner(fin, fin).real - 1.0) < 1e-12
        dev_layers = [gates_to_layers(r["G"][j, k], r["kind"][j, k]) for j in range(layers)]
        for lay in dev_layers:
            O.unitaries_sanity_check(lay)
        tf = O.sequential_unitary_circuit(tgt, layers, unitaries_override=dev_layers, record_svals=True, **kw)
        assert np.abs(np.array(tf["disentangled_mps_overlaps"]) - r["dis"][:, k]).max() < 1e-10
        assert np.abs(np.array(tf["overlaps"]) - r["overlaps"][ev, k]).max() < 1e-10
        for j in range(layers):
            for q, s in enumerate(tf["d2_svals"][j]):
                i = L - 1 - q
                n = min(2, len(s))
                # 1e-11 relative where few truncations precede the bond; the values further along are the product of
                # q chained rank-2 truncations (each conditioned by the Schmidt gap s0 / (s1 - s2) of a dense
                # spectrum, magnitudes down to 1e-8), where LAPACK and Jacobi drift apart by their own rounding
                rtol = 1e-11 if q < 16 else 1e-8
                assert np.allclose(r["d2_svals"][j, k, i - 1, :n], s[:n], rtol=rtol, atol=1e-13 * s[0]), (j, i)


def test_config2_L50_chi16_4layers(drv):
    """BASELINE configs[1]: random MPS L=50, chi=16, 4 layers (two instances of the batch)."""
    _check_named_config(drv, 50, 16, 4, 2, 300)


def test_config3c_L100_chi32(drv):
    """BASELINE configs[2] shape with random inputs (SURVEY 8d cfg 3c): L=100, chi=32, 2 instances x 2 layers."""
    _check_named_config(drv, 100, 32, 2, 2, 310)


def test_config4_L100_chi64_4layers(drv):
    """BASELINE configs[3] as named: L=100, chi=64, all 4 layers, one instance (about a minute of oracle time).
    The later layers are where the Jacobi sweeps see graded spectra."""
    _check_named_config(drv, 100, 64, 4, 1, 320, fid_tol=1e-8)


# ---------------------------------------------------------------------------------------
# Jacobi SVD vs LAPACK on graded spectra (the convergence rule of qsp_jacobi.cuh:17-20)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("decades", [0, 4, 9])
def test_jacobi_graded_spectrum_vs_lapack(drv, decades):
    """128 x 128 complex matrices (the chi = 64 bulk problem, register-resident path, 512 threads) with singular
    values graded over `decades` decades, as the disentangled states of the deeper layers have them."""
    import ctypes

    import torch

    from qsp_b200 import _runtime

    dev = _runtime.dev_lib()
    n, B = 128, 6
    rng = np.random.default_rng(7 + decades)
    W = np.empty((B, n, n), np.complex128)
    for b in range(B):
        u, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        v, _ = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        s = 10.0 ** (-decades * np.sort(rng.random(n)))
        if b == 1:
            s[5:9] = s[5]  # a cluster
        W[b] = (u * s) @ v.conj().T
    ref = np.stack([np.linalg.svd(W[b], compute_uv=False) for b in range(B)])
    Wd = torch.from_numpy(W).cuda()
    sw = torch.zeros(B, dtype=torch.int32, device="cuda")
    ok = torch.zeros(B, dtype=torch.int32, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    dev.qsp_jacobi_bench(p(Wd), n, n, n, B, 512, p(sw), p(ok), drv.stream)
    torch.cuda.synchronize()
    # (cyclic Jacobi needs more sweeps the more strongly graded and mixed the spectrum: 9 decades take 25..28 here)
    assert int(ok.min()) == 1 and int(sw.max()) <= 14 + 2 * decades, (ok.tolist(), sw.tolist())
    R = Wd.cpu().numpy()
    for b in range(B):
        nr = np.linalg.norm(R[b], axis=1)
        got = np.sort(nr)[::-1]
        # LAPACK's own accuracy is absolute (eps * s0); Jacobi's is relative -- compare at the weaker of the two
        assert np.abs(got - ref[b]).max() < 1e-12 * ref[b][0]
        big = ref[b] > 1e-3 * ref[b][0]
        assert np.abs(got[big] / ref[b][big] - 1.0).max() < 1e-12
        Q = R[b] / nr[:, None]
        assert np.abs(Q @ Q.conj().T - np.eye(n)).max() < 1e-13 * np.sqrt(n) * 8


def test_tensor_memory_row_store_experiment_is_bit_identical(drv):
    """The development-build experiment that keeps 48 of the 64 shared rows of the 128 x 128 Jacobi in tensor memory
    (tcgen05.ld / st, profiles/r2_tmem_jacobi.md) performs the same arithmetic in the same order as the product's
    shared-memory variant: identical sweep counts and bit-identical rows."""
    import ctypes

    import torch

    from qsp_b200 import _runtime

    dev = _runtime.dev_lib()
    n, B = 128, 5
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    W0 = torch.complex(torch.randn((B, n, n), generator=g, device="cuda", dtype=torch.float64),
                       torch.randn((B, n, n), generator=g, device="cuda", dtype=torch.float64))
    W0[1, 3] = W0[1, 2]   # dependent rows
    W0[2] *= torch.logspace(0, -8, n, dtype=torch.float64, device="cuda")  # graded columns
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    out = []
    try:
        for flags in (4, 6):  # bit 2: no norm ordering (the experiment does not have it), bit 1: rows in tensor memory
            dev.qsp_debug_flags(flags)
            W = W0.clone()
            sw = torch.zeros(B, dtype=torch.int32, device="cuda")
            ok = torch.zeros(B, dtype=torch.int32, device="cuda")
            dev.qsp_jacobi_bench(p(W), n, n, n, B, 512, p(sw), p(ok), drv.stream)  # (the wrapper raises on an error code)
            torch.cuda.synchronize()
            assert int(ok.min()) == 1
            out.append((W, sw))
    finally:
        dev.qsp_debug_flags(0)
    assert torch.equal(out[0][1], out[1][1])
    assert torch.equal(torch.view_as_real(out[0][0]), torch.view_as_real(out[1][0]))


def test_overlapped_chunks_match_single_call(drv):
    """Engine.seq_prepare splits a batch into chunks that alternate between two streams; HostBatchPipeline streams host
    chunks through the device: both give bit-identical results to one resident call."""
    import torch

    from bench import synth_batch
    from qsp_b200.engine import Engine
    from qsp_b200.tsp import HostBatchPipeline

    eng = Engine()
    B, L, chi, layers = 10, 12, 8, 3
    raw, bonds = synth_batch(torch, B, L, chi, eng.device, 5)
    host = [(raw[c:c + 4].cpu().pin_memory(), bonds[c:c + 4].cpu().pin_memory()) for c in range(0, B, 4)]
    assert int(eng.prepare_target(raw, bonds).max()) == 0
    one = eng.seq_prepare(raw, bonds, layers, max_bond=chi)
    many = eng.seq_prepare(raw, bonds, layers, max_bond=chi, chunk=3)
    torch.cuda.synchronize()
    for k in ("overlaps", "dis_overlaps", "G", "kind", "d2_svals", "status"):
        a, b = getattr(one, k), getattr(many, k)
        assert torch.equal(torch.view_as_real(a) if a.is_complex() else a, torch.view_as_real(b) if b.is_complex() else b) or \
            bool(((a == b) | (a != a)).all()), k  # (NaN marks the layers without an encoded-state evaluation)
    pipe = HostBatchPipeline(L, chi, layers, chi, 4)
    out = HostBatchPipeline.concat(pipe.run(host))
    assert pipe.h2d_bytes == B * (L * chi * 2 * chi * 16 + (L + 1) * 4)
    assert torch.equal(out["G"], one.G.cpu()) and torch.equal(out["kind"], one.kind.cpu()) and torch.equal(out["status"], one.status.cpu())
    assert torch.equal(torch.view_as_real(out["dis_overlaps"]), torch.view_as_real(one.dis_overlaps.cpu()))


def test_overlapped_chunks_large_bonds(drv):
    """bonds > 64: the chunks of a batch run on two streams fed by two host threads (the large-bond sweeps synchronise their
    stream once per Jacobi sweep); same results as one call up to rounding (the column splits of the Gram launches, and with
    them the summation order, depend on the chunk size)"""
    import torch

    from bench import synth_batch
    from qsp_b200.engine import Engine

    eng = Engine()
    B, L, chi, layers = 6, 14, 70, 2
    raw, bonds = synth_batch(torch, B, L, chi, eng.device, 9)
    assert int(eng.prepare_target(raw, bonds).max()) == 0
    one = eng.seq_prepare(raw, bonds, layers, max_bond=chi)
    many = eng.seq_prepare(raw, bonds, layers, max_bond=chi, chunk=2)
    torch.cuda.synchronize()
    assert int(one.status.max()) == 0 and int(many.status.max()) == 0
    # (bit-identical in every isolated run, scripts/stress_chunks.py; the bound is that of "equal up to rounding" after
    # two layers of D=2 truncations: summation order and sweep counts may differ between chunkings)
    assert float((one.dis_overlaps - many.dis_overlaps).abs().max()) < 1e-9
    assert float((one.overlaps[layers - 1] - many.overlaps[layers - 1]).abs().max()) < 1e-9
    assert float((one.d2_svals - many.d2_svals).abs().max()) < 1e-9 and torch.equal(one.kind, many.kind)


def test_batched_circuits_and_statevector_check(drv):
    """(overlap, circ) for a batch: circuits are materialised on demand from the device-resident gates, and the
    reference's statevector cross-check (qiskit_circuit.py:52-67) runs on the device for all instances at 18 qubits
    (the reference stops at 16)."""
    import torch

    from qsp_b200.tsp import BatchedMPSPreparation

    L, chi, layers, B = 18, 8, 3, 5
    raws = [O.random_mps(L, chi, 400 + s) for s in range(B)]
    prep = BatchedMPSPreparation(raws, shape="lpr")
    ov, res = prep.sequential_unitary_circuit(layers, do_compression=True, max_bond_dim=chi, check=True)
    co = prep.circuit_overlaps(chunk=2)
    torch.cuda.synchronize()
    assert float((co - ov.abs()).abs().max()) < 1e-11
    circ = prep.circuit(3)
    assert prep.circuit(3) is circ and set(circ.count_ops()) == {"u3", "cx"}
    assert circ.source.size() == layers * L and circ.source.num_nonlocal_gates() == layers * (L - 1)
    many = prep.circuits(workers=2)  # bulk synthesis over host processes
    assert len(many) == B and many[3] is circ and all(c.source.size() == layers * L for c in many)
    assert many[1].count_ops() == prep.circuit(1).count_ops()
    sv = circ.statevector().reshape(-1)  # the synthesised {cx, u3} circuit on the device kernel
    tgt = prep.engine.dense_from_packed(prep.mps[3:4], prep.bonds[3:4])[0]
    assert abs(abs(complex((tgt.conj() * sv).sum())) - abs(complex(ov[3]))) < 1e-10


def test_heisenberg_twelve_layers_reference_benchmark(drv):
    """the reference's own benchmark (examples/benchmarks.py:57-60): Heisenberg L=32 ground state, max_bond_dim=256,
    12 layers -- working bond 256 (large-bond kernels), encoded state of rank 517 beyond the doubling buffers.  Golden
    values: the oracle with the same completion convention (tests/golden/make_heisenberg_canonical.py, four CPU minutes)."""
    import json
    import os

    from qsp_b200.mps import MPSState
    from qsp_b200.tsp import MPSPreparation

    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    ref = json.load(open(os.path.join(gold, "heisenberg_canonical_12layers.json")))
    z = np.load(os.path.join(gold, "heisenberg_L32.npz"))
    b, m = z["bonds"], z["mps"]
    raw = [m[i, : b[i], :, : b[i + 1]].astype(np.complex128) for i in range(len(b) - 1)]
    prep = MPSPreparation(MPSState(raw), shape="lpr")
    ov, circ = prep.sequential_unitary_circuit(12, do_compression=True, max_bond_dim=256)
    d = prep.seq_data
    assert d["it_Ds"] == ref["it_Ds"]
    assert np.abs(np.abs(d["overlaps"]) - np.array(ref["abs_overlaps"])).max() < 1e-7
    assert np.abs(np.abs(d["disentangled_mps_overlaps"]) - np.array(ref["abs_disentangled_overlaps"])).max() < 1e-7
    assert abs(abs(ov) - ref["abs_overlaps"][-1]) < 1e-7 and 0.70 < abs(ov) < 0.80  # pickle (LAPACK convention): 0.748
    assert circ.source.size() == 12 * 32
