"""Kernel LOGIC on the CPU: the very sources under qsp_b200/csrc compiled against the fiber
emulator (tests/emu/cuda_emu.h) and driven through the same C ABI, compared with the oracle
on small seeded cases.  (The GPU parity tests in test_gpu_parity.py are the ones that
count for parity; these catch indexing / ordering bugs where no GPU is available.)"""
import numpy as np
import pytest

import proto_fused as P
from helpers import HostDriver, edge_cases, gates_to_layers, layers_to_gates, pack_batch, unpack_batch
from oracle import qsp_oracle as O


@pytest.fixture(scope="module")
def drv():
    d = HostDriver()
    assert not d.lib.is_device_build()
    return d


def _prepared(mps):
    t = O.prepare_target(mps, "lpr")
    O.right_canonize(t, normalize=True)
    return t


def test_left_canonize(drv):
    mpss = [O.random_mps(7, 5, s) for s in range(2)]
    for m in mpss:
        for i in range(7):
            m[i] = m[i] / np.linalg.norm(m[i])
    a, b = pack_batch(mpss, 6)
    assert drv.left_canonize(a, b).max() == 0
    for m, g in zip(mpss, unpack_batch(a, b)):
        ref = O.left_canonize(O.copy_mps(m))
        assert O.bonds_of(ref) == O.bonds_of(g)
        assert max(np.abs(x - y).max() for x, y in zip(ref, g)) < 1e-13


@pytest.fixture
def big_path(drv):
    """qsp_debug_flags bit 0: left-canonise out of the L2 scratch (the chi > 85 path) at any size"""
    drv.lib.qsp_debug_flags(1)
    yield drv
    drv.lib.qsp_debug_flags(0)


def test_large_bond_left_canonize_path(big_path):
    drv = big_path
    assert drv.lib.qsp_left_canonize_workspace_bytes(2, 7, 6) > 0
    test_left_canonize(drv)
    test_prepare_target(drv)
    test_overlap(drv)  # environments as batched tensor-core GEMMs
    test_layer_loop(drv, "compressed")
    assert drv.lib._qsp_left_canonize(drv.p(np.zeros(4 * 2 * 4 * 4, np.complex128)), drv.p(np.ones(5, np.int32)), 1, 4, 4,
                                      None, None, 0, None) == 3  # QSP_ERR_WORKSPACE


@pytest.mark.parametrize("max_bond", [0, 3, 2])
def test_right_compress(drv, max_bond):
    mpss = []
    for s in range(2):
        m = O.random_mps(7, 6, 10 + s)
        m[0] = m[0] / np.sqrt(O.inner(m, m).real)
        mpss.append(O.left_canonize(m))
    a, b = pack_batch(mpss, 6)
    out, bo, st, sv, amp = drv.right_compress(a, b, 6, max_bond=max_bond, normalize=1, nsv=3, amp0=True)
    assert st.max() == 0
    for k, (m, g) in enumerate(zip(mpss, unpack_batch(out, bo))):
        svl = []
        ref = O.right_compress(O.copy_mps(m), max_bond=max_bond or None, svals=svl)
        O.right_canonize(ref, normalize=True)
        assert O.bonds_of(ref) == O.bonds_of(g)
        assert abs(O.inner(g, ref) - 1.0) < 1e-13
        for j, s in enumerate(svl):
            n = min(3, len(s))
            assert np.allclose(sv[k, 7 - 1 - j - 1, :n], s[:n], rtol=1e-12)
        assert abs(amp[k] - O.norm_mps_ovrlap(ref, O.cl_zero_mps(7))) < 1e-13


def test_block_jacobi_path(drv):
    """more rows than fit the (emulated) shared buffer are rotated block-pair by block-pair:
    force it by shrinking the row capacity through a tall matrix (b > capacity)."""
    # L=2 chain: site 1 is (b, 2, 1) -> mode C (b > c); site rows live in mode R for the bulk case below
    m = O.random_mps(8, 16, 77)
    m[0] = m[0] / np.sqrt(O.inner(m, m).real)
    m = O.left_canonize(m)
    a, b = pack_batch([m], 16)
    out, bo, st, sv, _ = drv.right_compress(a, b, 16, max_bond=8, normalize=1, nsv=2)
    assert st.max() == 0
    ref = O.right_compress(O.copy_mps(m), max_bond=8)
    O.right_canonize(ref, normalize=True)
    g = unpack_batch(out, bo)[0]
    assert abs(O.inner(g, ref) - 1.0) < 1e-12


def _jacobi_cases(nrows, ncols, seed=3):
    rng = np.random.default_rng(seed)
    g = lambda *sh: rng.standard_normal(sh) + 1j * rng.standard_normal(sh)
    full = g(nrows, ncols)
    low = g(nrows, 3) @ g(3, ncols)                          # rank 3: most rows collapse to roundoff
    graded = g(nrows, ncols) * np.logspace(0, -9, ncols)     # graded columns (the carried U s)
    dup = full.copy(); dup[1] = dup[0]; dup[5] = 2 * dup[4]  # exactly dependent rows
    zero = full.copy(); zero[2] = 0                          # a zero row
    # clustered singular values (relative gaps 1e-6 and 1e-9): the quadratic-convergence stop must not leave the
    # near-degenerate directions mixed
    k = min(nrows, ncols)
    u, _ = np.linalg.qr(g(nrows, k))
    v, _ = np.linalg.qr(g(ncols, k))
    sv = 1.0 + 1e-6 * np.arange(k)
    sv[k // 2:] = 0.5 + 1e-9 * np.arange(k - k // 2)
    clustered = (u * sv) @ v.conj().T
    return {"full": full, "rank3": low, "graded": graded, "duplicates": dup, "zero_row": zero, "clustered": clustered}


@pytest.mark.parametrize("nrows,ldw,nt", [(16, 128, 64), (15, 128, 64), (8, 64, 64), (128, 128, 512)])
def test_jacobi_register_variants(drv, nrows, ldw, nt):
    """the register-resident odd-even Jacobi on its own: (16,128,64) is the 16-columns-per-lane geometry of the
    headline shape (plain rotations, carried row norms); (128,128,512) is the headline problem itself, which keeps 48 of
    its 64 shared rows in (emulated) tensor memory; singular values to 1e-12, rows orthogonal."""
    cases = _jacobi_cases(nrows, ldw)
    if nrows >= 128:  # (emulation time: 512 fibers; the full-rank and graded cases run on the GPU, test_gpu_exports.py)
        cases = {k: cases[k] for k in ("clustered",)}
    W = np.ascontiguousarray(np.stack(list(cases.values())))
    B = W.shape[0]
    ref = W.copy()
    sw, ok = np.zeros(B, np.int32), np.zeros(B, np.int32)
    assert drv.lib._qsp_jacobi_bench(drv.p(W), nrows, ldw, ldw, B, nt, drv.p(sw), drv.p(ok), None) == 0
    assert ok.min() == 1, (ok, sw)
    for k, name in enumerate(cases):
        s_ref = np.linalg.svd(ref[k], compute_uv=False)
        s = np.sort(np.linalg.norm(W[k], axis=1))[::-1]
        # 64 singular values inside a relative window of 6e-8: the quadratic-convergence stop (largest rotation of the
        # last sweep below 3e-8) leaves ~1e-11 there, the same in the shared-memory and the tensor-memory variant
        rtol = 1e-10 if (name == "clustered" and nrows >= 128) else 1e-12
        assert np.allclose(s, s_ref, rtol=rtol, atol=1e-13 * s_ref[0]), name
        keep = np.linalg.norm(W[k], axis=1) > 1e-8 * s_ref[0]
        V = W[k][keep] / np.linalg.norm(W[k][keep], axis=1)[:, None]
        # (the same 64-fold cluster: every rotation inside it is a large-angle one however small the Gram entry, so the
        # last sweep's rotations leave ~1e-9 between its members; multiplets of a few states are not affected)
        otol = 1e-13 if name != "clustered" else (1e-8 if nrows >= 128 else 1e-12)
        assert np.abs(V @ V.conj().T - np.eye(keep.sum())).max() < otol, name
        # the rotations are unitary: the row space (Gram matrix of columns) is unchanged
        assert np.abs(W[k].conj().T @ W[k] - ref[k].conj().T @ ref[k]).max() < 1e-12 * s_ref[0] ** 2, name


def test_unitaries_and_layers(drv):
    L = 8
    tg = [_prepared(O.random_mps(L, 4, 20 + s)) for s in range(2)]
    d2s = [O.d2_truncate(t) for t in tg]
    a, b = pack_batch(d2s, 2)
    G, kind, st = drv.generate_unitaries(a, b)
    assert st.max() == 0
    for k, d2 in enumerate(d2s):
        ref = O.generate_unitaries(d2, completion="canonical")
        got = gates_to_layers(G[k], kind[k])
        for (s0, e0, gr), (s1, e1, gg) in zip(ref, got):
            assert (s0, e0) == (s1, e1)
            assert max(np.abs(x - y).max() for x, y in zip(gr, gg)) < 1e-13
    lc = [O.left_canonize(O.copy_mps(t)) for t in tg]
    a, b = pack_batch(lc, 4)
    out, bo, st = drv.apply_layer(a, b, G, kind, 8, inverse=True)
    assert st.max() == 0
    for t, k, g in zip(tg, range(2), unpack_batch(out, bo)):
        ref = O.apply_inverse_unitary_layer_on_wfn(gates_to_layers(G[k], kind[k]), t)
        assert abs(O.inner(g, ref) - 1.0) < 1e-13
    a, b = pack_batch(tg, 4)
    out, bo, st = drv.apply_layer(a, b, G, kind, 8, inverse=False)
    for t, k, g in zip(tg, range(2), unpack_batch(out, bo)):
        ref = O.apply_unitary_layer_on_wfn(gates_to_layers(G[k], kind[k]), t)
        assert abs(O.inner(g, ref) - 1.0) < 1e-13


def test_segmented_layer(drv):
    """a D=2 state with collapsed bonds: one-site segments, one-qubit segment ends (H5)."""
    L = 7
    one = np.zeros((1, 2, 1), complex)
    one[0, :, 0] = [0.6, 0.8j]
    bell = np.zeros((1, 2, 2), complex)
    bell[0, 0, 0], bell[0, 1, 1] = 0.8, 0.6
    mid = np.zeros((2, 2, 2), complex)
    mid[0, 0, 0], mid[0, 1, 1], mid[1, 1, 0], mid[1, 0, 1] = 0.6, 0.8, 0.8, -0.6
    end = np.zeros((2, 2, 1), complex)
    end[0, 0, 0] = end[1, 1, 0] = 1
    d2 = [one.copy(), bell.copy(), mid.copy(), end.copy(), one.copy(), bell.copy(), end.copy()]
    O.right_canonize(d2, normalize=True)
    a, b = pack_batch([d2], 2)
    G, kind, st = drv.generate_unitaries(a, b)
    assert st.max() == 0 and kind[0].tolist() == [1, 2, 2, 1, 1, 2, 1]
    ref = O.generate_unitaries(d2, completion="canonical")
    got = gates_to_layers(G[0], kind[0])
    assert [(s, e) for s, e, _ in got] == [(s, e) for s, e, _ in ref] == [(0, 0), (1, 3), (4, 4), (5, 6)]
    for (_, _, gr), (_, _, gg) in zip(ref, got):
        assert max(np.abs(x - y).max() for x, y in zip(gr, gg)) < 1e-13
    # apply to a chi=4 state: inverse on left-canonical, forward on right-canonical
    psi = _prepared(O.random_mps(L, 4, 5))
    a, b = pack_batch([O.left_canonize(O.copy_mps(psi))], 4)
    out, bo, st = drv.apply_layer(a, b, G, kind, 8, inverse=True)
    assert st.max() == 0
    g = unpack_batch(out, bo)[0]
    assert abs(O.inner(g, O.apply_inverse_unitary_layer_on_wfn(ref, psi)) - 1.0) < 1e-13
    a, b = pack_batch([psi], 4)
    out, bo, st = drv.apply_layer(a, b, G, kind, 8, inverse=False)
    g = unpack_batch(out, bo)[0]
    assert abs(O.inner(g, O.apply_unitary_layer_on_wfn(ref, psi)) - 1.0) < 1e-13


def test_fused_inverse_and_compress(drv):
    """qsp_apply_inverse_and_compress == apply_inverse_unitary_layer_on_wfn + right_canonize +
    compress('right', max_bond) of the reference (core.py:224-240, sequential.py:51-54), segmented layer included."""
    L = 7
    psis = [_prepared(O.random_mps(L, 4, 60 + s)) for s in range(2)]
    layers = [O.generate_bond_d_unitary(p, completion="canonical") for p in psis]
    # second instance: a hand-made segmented layer (one-site segments and one-qubit segment ends)
    one = np.zeros((1, 2, 1), complex)
    one[0, :, 0] = [0.6, 0.8j]
    bell = np.zeros((1, 2, 2), complex)
    bell[0, 0, 0], bell[0, 1, 1] = 0.8, 0.6
    mid = np.zeros((2, 2, 2), complex)
    mid[0, 0, 0], mid[0, 1, 1], mid[1, 1, 0], mid[1, 0, 1] = 0.6, 0.8, 0.8, -0.6
    end = np.zeros((2, 2, 1), complex)
    end[0, 0, 0] = end[1, 1, 0] = 1
    d2 = [one.copy(), bell.copy(), mid.copy(), end.copy(), one.copy(), bell.copy(), end.copy()]
    O.right_canonize(d2, normalize=True)
    layers[1] = O.generate_unitaries(d2, completion="canonical")
    Gk = [layers_to_gates(l, L) for l in layers]
    G, kind = np.stack([g for g, _ in Gk]), np.stack([k for _, k in Gk])
    lc = [O.left_canonize(O.copy_mps(p)) for p in psis]
    for mb in (0, 4, 3):
        a, b = pack_batch(lc, 8 if mb == 0 else 4)
        out, bo, st, amp = drv.apply_inverse_and_compress(a, b, G, kind, max_bond=mb)
        assert st.max() == 0
        for k, (p, l, g) in enumerate(zip(psis, layers, unpack_batch(out, bo))):
            ref = O.apply_inverse_unitary_layer_on_wfn(l, p)
            O.right_canonize(ref, normalize=True)
            O.compress(ref, max_bond=mb or None)
            O.right_canonize(ref, normalize=True)
            assert O.bonds_of(ref) == O.bonds_of(g), (mb, k)
            assert abs(O.inner(g, ref) - 1.0) < 1e-12
            assert abs(amp[k] - O.norm_mps_ovrlap(ref, O.cl_zero_mps(L))) < 1e-12


def test_overlap(drv):
    xs = [O.random_mps(6, 3, 1), O.random_mps(6, 3, 2)]
    ys = [O.random_mps(6, 5, 3), O.random_mps(6, 5, 4)]
    a, ba = pack_batch(xs, 3)
    b, bb = pack_batch(ys, 5)
    res = drv.overlap(a, ba, b, bb)
    for k in range(2):
        ref = O.inner(xs[k], ys[k])
        assert abs(res[k] - ref) < 1e-13 * abs(ref)


def test_prepare_target(drv):
    raw = [O.to_lpr(O.readme_mps(6, 3, s), "lrp") for s in range(2)]
    a, b = pack_batch(raw, 4)
    assert drv.prepare_target(a, b).max() == 0
    for r, g in zip(raw, unpack_batch(a, b)):
        ref = O.prepare_target(r, "lpr")
        O.right_canonize(ref, normalize=True)
        assert O.bonds_of(ref) == O.bonds_of(g)
        assert abs(O.inner(g, ref) - 1.0) < 1e-13


@pytest.mark.parametrize("case", ["compressed", "readme_nocomp", "second_pass", "more_layers_than_half_chain",
                                  "ten_layers_short_chain"])
def test_layer_loop(drv, case):
    if case == "readme_nocomp":
        tg = [_prepared(O.to_lpr(O.readme_mps(6, 2, s), "lrp")) for s in range(2)]
        layers, mb, D, flags = 3, 0, 8, 0
    elif case == "more_layers_than_half_chain":  # encoded state saturates at 2^(L/2)
        tg = [_prepared(O.random_mps(4, 4, 40 + s)) for s in range(1)]
        layers, mb, D, flags = 4, 4, 4, 0
    elif case == "ten_layers_short_chain":  # > 8 layers: the encoded state is compressed between layers
        tg = [_prepared(O.random_mps(4, 4, 5))]
        layers, mb, D, flags = 10, 4, 4, 0
    else:
        tg = [_prepared(O.random_mps(6, 4, 30 + s)) for s in range(2)]
        layers, mb, D, flags = 2, 4, 4, (1 if case == "second_pass" else 0)
    a, b = pack_batch(tg, D)
    r = drv.seq_prepare(a, b, layers, max_bond=mb, flags=flags)
    assert r["status"].max() == 0
    ev = [i for i in range(layers) if i % 4 == 0 or i == layers - 1]
    for k, t in enumerate(tg):
        kw = dict(do_compression=mb > 0, max_bond_dim=mb or None)
        sc = O.sequential_unitary_circuit(t, layers, completion="canonical", record_svals=True, **kw)
        assert np.abs(np.array(sc["disentangled_mps_overlaps"]) - r["dis"][:, k]).max() < 1e-12
        assert np.abs(np.array(sc["overlaps"]) - r["overlaps"][ev, k]).max() < 1e-12
        dev_layers = [gates_to_layers(r["G"][j, k], r["kind"][j, k]) for j in range(layers)]
        tf = O.sequential_unitary_circuit(t, layers, unitaries_override=dev_layers, **kw)
        assert np.abs(np.array(tf["disentangled_mps_overlaps"]) - r["dis"][:, k]).max() < 1e-12
        for j in range(layers):
            for q, s in enumerate(sc["d2_svals"][j]):
                n = min(2, len(s))
                assert np.allclose(r["d2_svals"][j, k, len(t) - 1 - q - 1, :n], s[:n], rtol=1e-11)
        fin = unpack_batch(r["final"], r["final_bonds"])[k]
        assert abs(O.inner(fin, sc["final_mps"]) - 1.0) < 1e-11


def test_status_bits_and_error_codes(drv):
    # bond larger than the padded buffer allows -> QSP_ST_BOND_OVERFLOW, instance skipped
    tg = [_prepared(O.to_lpr(O.readme_mps(8, 4, 0), "lrp"))]
    a, b = pack_batch(tg, 4)
    r = drv.seq_prepare(a, b, 2, max_bond=0)  # no compression: bonds want to double beyond D=4
    assert r["status"][0] & 2
    # no bond is too large any more (the batched large-bond path takes over beyond one SM); a bad shape is refused
    # before anything is launched
    assert drv.lib.qsp_seq_prepare_workspace_bytes(1, 10, 512, 2, 512, 0) > 0
    assert drv.lib.qsp_seq_prepare_workspace_bytes(0, 10, 8, 2, 8, 0) == 0
    assert drv.lib._qsp_left_canonize(None, None, 1, 4, 4, None, None, 0, None) == 1


@pytest.mark.parametrize("name", list(edge_cases()))
def test_edge_cases(drv, name):
    """empty-ish and ragged inputs the reference accepts: product states (every bond 1, all
    one-qubit gates), zero-padded rank-deficient tensors, degenerate Schmidt values, L = 1, 2, 3."""
    raw, layers, mb, D = edge_cases()[name]
    t = _prepared(raw)
    a, b = pack_batch([t], D)
    r = drv.seq_prepare(a, b, layers, max_bond=mb)
    # the GHZ state's Schmidt pair (1/sqrt2, 1/sqrt2) is exactly degenerate: reported, not an error
    assert r["status"][0] == (8 if name == "ghz_degenerate_schmidt" else 0)
    dev_layers = [gates_to_layers(r["G"][j, 0], r["kind"][j, 0]) for j in range(layers)]
    kw = dict(do_compression=mb > 0, max_bond_dim=mb or None)
    tf = O.sequential_unitary_circuit(t, layers, unitaries_override=dev_layers, **kw)
    ev = [i for i in range(layers) if i % 4 == 0 or i == layers - 1]
    assert np.abs(np.array(tf["disentangled_mps_overlaps"]) - r["dis"][:, 0]).max() < 1e-12
    assert np.abs(np.array(tf["overlaps"]) - r["overlaps"][ev, 0]).max() < 1e-12
    sc = O.sequential_unitary_circuit(t, layers, completion="canonical", **kw)
    assert [O.unitaries_specs(l) for l in sc["unitaries"]] == [O.unitaries_specs(l) for l in dev_layers]


@pytest.mark.parametrize("L,chi,layers,cap_log2,fits", [(12, 2, 6, 4, True), (10, 4, 8, 6, True), (14, 6, 6, 4, False)])
def test_encoded_state_beyond_buffer_capacity(drv, L, chi, layers, cap_log2, fits):
    """more layers than the encoded-state buffers can double through (2^layers > cap) on ANY chain: like the
    reference (core.py:139-141) the encoded state is compressed between layers; a true rank beyond the buffers is
    truncated and reported as QSP_ST_BOND_OVERFLOW, never silently.  The test hook shrinks the buffers so that the
    logic of the 256-padded production path runs at emulator-friendly sizes."""
    drv.lib.qsp_debug_flags(cap_log2 << 4)
    try:
        t = _prepared(O.random_mps(L, chi, 5))
        a, b = pack_batch([t], chi)
        r = drv.seq_prepare(a, b, layers, max_bond=chi)
    finally:
        drv.lib.qsp_debug_flags(0)
    dl = [gates_to_layers(r["G"][j, 0], r["kind"][j, 0]) for j in range(layers)]
    tf = O.sequential_unitary_circuit(t, layers, unitaries_override=dl, do_compression=True, max_bond_dim=chi)
    assert np.abs(np.array(tf["disentangled_mps_overlaps"]) - r["dis"][:, 0]).max() < 1e-12
    ev = [i for i in range(layers) if i % 4 == 0 or i == layers - 1]
    if fits:
        assert r["status"][0] == 0
        assert np.abs(np.array(tf["overlaps"]) - r["overlaps"][ev, 0]).max() < 1e-12
    else:
        assert r["status"][0] == 2  # QSP_ST_BOND_OVERFLOW
        assert abs(tf["overlaps"][0] - r["overlaps"][0, 0]) < 1e-12  # the evaluations that fitted are still exact


# ---------------------------------------------------------------------------------------
# large-bond path (qsp_big.cuh / qsp_big_sweeps.cuh): tensor-core GEMM, block Jacobi, batched sweep
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,k,ta,ca,tb,cb", [(70, 45, 33, 0, 0, 0, 0), (16, 130, 9, 1, 1, 0, 0), (65, 64, 64, 0, 0, 1, 1),
                                              (3, 5, 0, 0, 0, 0, 0), (100, 7, 50, 1, 0, 1, 0)])
def test_big_gemm(drv, m, n, k, ta, ca, tb, cb):
    rng = np.random.default_rng(m + n + k)
    nb = 2
    A = rng.standard_normal((nb, m, k)) + 1j * rng.standard_normal((nb, m, k))
    Bm = rng.standard_normal((nb, k, n)) + 1j * rng.standard_normal((nb, k, n))
    As = np.ascontiguousarray((A.conj() if ca else A).transpose(0, 2, 1) if ta else (A.conj() if ca else A))
    Bs = np.ascontiguousarray((Bm.conj() if cb else Bm).transpose(0, 2, 1) if tb else (Bm.conj() if cb else Bm))
    C = np.full((nb, m, n), np.nan + 0j)
    work = np.zeros(nb * 128 + 64, np.uint8)
    drv.lib.qsp_big_gemm(drv.p(As), ta, ca, drv.p(Bs), tb, cb, drv.p(C), m, n, k, nb, drv.p(work), None)
    assert np.abs(C - A @ Bm).max() < 1e-12 * max(k, 1)


@pytest.mark.parametrize("nrows,ndot,extra", [(2, 32, 0), (16, 32, 0), (31, 64, 0), (40, 64, 0), (33, 64, 64)])
def test_big_block_jacobi(drv, nrows, ndot, extra):
    """rows of [M | I-ish] orthogonalised: singular values against LAPACK 1e-12, rows orthogonal; graded and
    rank-deficient inputs included (zero rows / dependent rows hit the Cholesky's zero-pivot branch)."""
    rng = np.random.default_rng(nrows * 7 + ndot)
    B = 3 if nrows <= 32 else 2
    ldw = ndot + extra
    W = np.zeros((B, nrows, ldw), np.complex128)
    live = ndot - 5  # the Gram part is zero padded to a multiple of 32
    for b in range(B):
        M = rng.standard_normal((nrows, live)) + 1j * rng.standard_normal((nrows, live))
        if b == 1:
            M *= (10.0 ** (-6 * rng.random(nrows)))[:, None]  # graded rows
        if b == B - 1 and nrows > 2:
            M[1] = M[0] * (0.3 + 0.1j)  # dependent rows
            M[-1] = 0
        W[b, :, :live] = M
        if extra:
            W[b, :, ndot:ndot + nrows] = np.eye(nrows)
    W0 = W.copy()
    sw, ok = np.zeros(B, np.int32), np.zeros(B, np.int32)
    work = np.zeros(drv.lib.qsp_big_jacobi_work_bytes(nrows, ldw, B), np.uint8)
    drv.lib.qsp_big_jacobi(drv.p(W), nrows, ndot, ldw, B, drv.p(sw), drv.p(ok), drv.p(work), None)
    assert ok.all() and sw.max() <= 20, (ok, sw)
    for b in range(B):
        ref = np.linalg.svd(W0[b, :, :ndot], compute_uv=False)
        nr = np.linalg.norm(W[b, :, :ndot], axis=1)
        got = np.sort(nr)[::-1]
        assert np.abs(got - ref).max() < 1e-13 * ref[0] * 10
        G = W[b, :, :ndot] @ W[b, :, :ndot].conj().T
        d = np.sqrt(np.abs(np.diag(G)))
        off = np.abs(G - np.diag(np.diag(G))) / np.maximum(np.outer(d, d), 1e-300)
        assert off[np.outer(d, d) > 1e-200].max(initial=0) < 1e-13
        if extra:  # accumulated rotation: W_out = J W_in with J unitary
            J = W[b, :, ndot:ndot + nrows]
            assert np.abs(J @ J.conj().T - np.eye(nrows)).max() < 1e-13
            assert np.abs(J @ W0[b, :, :ndot] - W[b, :, :ndot]).max() < 1e-12 * ref[0]


@pytest.mark.parametrize("L,chi,max_bond", [(7, 24, 0), (8, 40, 12)])
def test_big_right_compress_multiblock(big_path, L, chi, max_bond):
    """truncating sweep through the batched large-bond kernels with several 16-row blocks per matrix"""
    drv = big_path
    mpss = []
    for s in range(2):
        m = O.random_mps(L, chi, 20 + s)
        m[0] = m[0] / np.sqrt(O.inner(m, m).real)
        mpss.append(O.left_canonize(m))
    Dout = max_bond if max_bond else chi
    a, b = pack_batch(mpss, chi)
    out, bo, st, sv, amp = drv.right_compress(a, b, Dout, max_bond=max_bond, normalize=1, nsv=3, amp0=True)
    assert st.max() == 0
    for k, (m, g) in enumerate(zip(mpss, unpack_batch(out, bo))):
        svl = []
        ref = O.right_compress(O.copy_mps(m), max_bond=max_bond or None, svals=svl)
        O.right_canonize(ref, normalize=True)
        assert O.bonds_of(ref) == O.bonds_of(g)
        assert abs(abs(O.inner(g, ref)) - 1.0) < 1e-12
        for j, s in enumerate(svl):
            i = L - 1 - j
            n = min(3, len(s))
            assert np.allclose(sv[k, i - 1, :n], s[:n], rtol=1e-12, atol=1e-14 * s[0])
        assert abs(amp[k] - O.norm_mps_ovrlap(ref, O.cl_zero_mps(L)) * np.exp(1j * np.angle(O.inner(g, ref)))) < 1e-12


def test_big_left_canonize_multipanel(big_path):
    """blocked Householder QR of the large-bond path with several 16-column panels, tall and wide site matrices,
    rank-deficient columns and a real-valued input; R >= 0 gauge makes the tensors comparable entry by entry"""
    drv = big_path
    mpss = [O.random_mps(7, 24, 1), O.random_mps(7, 20, 2, dtype=np.float64).copy()]
    mpss[1] = [x.astype(np.complex128) for x in mpss[1]]
    wide = [np.random.default_rng(3).standard_normal((1 if i == 0 else 20, 2, 1 if i == 5 else 20)) + 0j for i in range(6)]
    wide.append(np.zeros((1, 2, 1), complex))
    wide = wide[:6]
    wide[2][:, :, 7] = wide[2][:, :, 3]  # dependent columns
    for m in mpss + [wide]:
        for i in range(len(m)):
            m[i] = m[i] / np.linalg.norm(m[i])
    for m in (mpss[0], mpss[1], wide):
        a, b = pack_batch([m], 24)
        assert drv.left_canonize(a, b).max() == 0
        g = unpack_batch(a, b)[0]
        ref = O.left_canonize(O.copy_mps(m))
        assert O.bonds_of(ref) == O.bonds_of(g)
        assert abs(O.inner(g, ref) / O.inner(ref, ref) - 1.0) < 1e-12
        for x, y in zip(ref[:-1], g[:-1]):
            q = y.reshape(-1, y.shape[2])
            assert np.abs(q.conj().T @ q - np.eye(q.shape[1])).max() < 1e-12
        if m is not wide:  # (with dependent columns the QR is not unique)
            assert max(np.abs(x - y).max() for x, y in zip(ref, g)) < 1e-12


def test_circuit_statevector_kernel(drv):
    """k_sv_gate: the batched statevector of the gate layers equals the oracle's dense simulation (qiskit ordering:
    layers reversed, sites ascending, a 4x4 on [it+1, it], qiskit_circuit.py:38-48), including segmented layers with
    one-qubit gates in the middle of the chain; a shared explicit gate list (qsp_statevector_apply) agrees too."""
    L = 7
    tgs = [_prepared(O.random_mps(L, 4, 41)), _prepared(edge_cases()["product_state"][0] + edge_cases()["product_state"][0][:2])]
    a, b = pack_batch(tgs, 4)
    r = drv.seq_prepare(a, b, 3, max_bond=4)
    state = np.full((2, 1 << L), np.nan + 0j)
    drv.lib.qsp_circuit_statevector(drv.p(r["G"]), drv.p(r["kind"]), 3, 2, L, drv.p(state), None)
    for k in range(2):
        layers = [gates_to_layers(r["G"][j, k], r["kind"][j, k]) for j in range(3)]
        ref = O.circuit_statevector(layers, L).reshape(-1)
        assert np.abs(state[k] - ref).max() < 1e-13
        assert abs(abs(np.vdot(O.mps_to_dense(tgs[k]).reshape(-1), state[k])) - abs(r["overlaps"][2, k])) < 1e-12
    # explicit gates, shared by the batch: instance 0's circuit applied to |0..0> on both rows
    st2 = np.zeros((2, 1 << L), np.complex128)
    st2[:, 0] = 1.0
    for j in reversed(range(3)):
        for it in range(L):
            g = np.ascontiguousarray(r["G"][j, 0, it])
            assert r["kind"][j, 0, it] in (1, 2)
            drv.lib.qsp_statevector_apply(drv.p(st2), 2, L, drv.p(g), 0, None, it, it + 1 if r["kind"][j, 0, it] == 2 else -1, None)
    assert np.abs(st2[0] - state[0]).max() < 1e-13 and np.abs(st2[1] - state[0]).max() < 1e-13


@pytest.mark.parametrize("name", ["product_state", "ghz_degenerate_schmidt", "two_sites"])
def test_edge_cases_through_large_bond_kernels(big_path, name):
    """segment boundaries (one-qubit gates inside the chain), collapsed bonds, degenerate Schmidt pairs and the shortest
    chains through the batched large-bond sweeps (their per-site branches differ from the one-CTA kernels)"""
    test_edge_cases(big_path, name)
