"""CPU arm of bench.py (TEST / MEASUREMENT INFRASTRUCTURE): times the oracle -- the numpy
restatement of the reference's sequential path -- on the host cores, and (parity mode) checks
sampled GPU results against it.

    python oracle/cpu_bench_worker.py time L chi layers n_inst cores reps workload   -> JSON list of wall seconds
    python oracle/cpu_bench_worker.py parity sample.npz cores                        -> JSON dict of differences

bench.py starts it in a fresh interpreter with OPENBLAS/OMP threads pinned to 1, so that
`cores` single-threaded workers do one MPS each."""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
WORKLOAD = "random"


def one(args):
    L, chi, layers, seed = args
    import numpy as np

    from oracle import qsp_oracle as O

    if WORKLOAD == "heisenberg" and L == 32:
        z = np.load(os.path.join(ROOT, "tests", "golden", "heisenberg_L32.npz"))
        b, m = z["bonds"], z["mps"]
        raw = [m[i, : b[i], :, : b[i + 1]].copy() for i in range(len(b) - 1)]
    elif WORKLOAD == "aklt":
        from qsp_b200.misc_states import make_aklt_mps, make_splitted_mps

        raw = O.to_lpr(make_splitted_mps(make_aklt_mps(L // 2)[0]), "lrp")
    else:
        raw = O.random_mps(L, chi, seed)
    # timed like the GPU arm's end-to-end figure: MPSPreparation.__init__ (normalise + compress, tsp.py:63-77) and the call
    t0 = time.time()
    t = O.prepare_target(raw, "lpr")
    O.sequential_unitary_circuit(t, layers, do_compression=True, max_bond_dim=chi, completion="lapack")
    return time.time() - t0


def parity_one(args):
    """teacher-forced check of one sampled instance: the device's unitaries through the oracle's literal
    reference steps (sequential.py:45-70); returns the largest differences of the per-layer numbers."""
    import numpy as np

    from oracle import qsp_oracle as O

    path, k = args
    z = np.load(path)
    bonds, mps = z["bonds"][k], z["mps"][k]
    L = mps.shape[0]
    tgt = [mps[i, : bonds[i], :, : bonds[i + 1]].copy() for i in range(L)]
    G, kind = z["G"][:, k], z["kind"][:, k]
    layers = G.shape[0]
    gs = []
    for j in range(layers):
        out, it = [], 0
        while it < L:
            s, g = it, []
            while True:
                n = 4 if kind[j, it] == 2 else 2
                g.append(np.array(G[j, it, :n, :n]))
                if kind[j, it] == 1:
                    break
                it += 1
            out.append([s, it, g])
            it += 1
        gs.append(out)
    mb = int(z["max_bond"])
    r = O.sequential_unitary_circuit(tgt, layers, do_compression=True, max_bond_dim=mb, unitaries_override=gs,
                                     record_svals=True)
    ev = [i for i in range(layers) if i % 4 == 0 or i == layers - 1]
    d_dis = float(np.abs(np.array(r["disentangled_mps_overlaps"]) - z["dis"][:, k]).max())
    d_ov = float(np.abs(np.array(r["overlaps"]) - z["overlaps"][ev, k]).max())
    d_sv = 0.0  # largest singular value of every bond, relative; bonds with few chained truncations before them
    for j in range(layers):
        for q, s in enumerate(r["d2_svals"][j][:16]):
            i = L - 1 - q
            d_sv = max(d_sv, float(abs(z["d2_svals"][j, k, i - 1, 0] / s[0] - 1.0)))
    return {"dis": d_dis, "overlaps": d_ov, "d2_svals_rel": d_sv}


def main():
    global WORKLOAD
    mode = sys.argv[1]
    ctx = mp.get_context("fork")
    if mode == "parity":
        path, cores = sys.argv[2], int(sys.argv[3])
        import numpy as np

        n = int(np.load(path)["mps"].shape[0])
        with ctx.Pool(min(cores, n)) as pool:
            res = pool.map(parity_one, [(path, k) for k in range(n)], chunksize=1)
        print(json.dumps({"instances": n, "max_abs_diff_dis_overlaps": max(r["dis"] for r in res),
                          "max_abs_diff_overlaps": max(r["overlaps"] for r in res),
                          "max_rel_diff_d2_svals": max(r["d2_svals_rel"] for r in res)}))
        return
    L, chi, layers, n_inst, cores, reps = (int(x) for x in sys.argv[2:8])
    WORKLOAD = sys.argv[8] if len(sys.argv) > 8 else "random"
    jobs = [(L, chi, layers, 1234 + i) for i in range(n_inst)]
    with ctx.Pool(cores) as pool:
        pool.map(one, [(8, 4, 1, 0)] * cores)  # imports / warm-up outside the timed region
        walls = []
        for _ in range(reps):
            t0 = time.time()
            pool.map(one, jobs, chunksize=1)
            walls.append(time.time() - t0)
    print(json.dumps(walls))


if __name__ == "__main__":
    main()
