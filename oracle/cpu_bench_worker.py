"""CPU arm of bench.py (TEST / MEASUREMENT INFRASTRUCTURE): times the oracle -- the numpy
restatement of the reference's sequential path -- on the host cores.

    python oracle/cpu_bench_worker.py L chi layers n_inst cores reps   -> JSON list of wall seconds

bench.py starts it in a fresh interpreter with OPENBLAS/OMP threads pinned to 1, so that
`cores` single-threaded workers do one MPS each."""
import json
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(args):
    L, chi, layers, seed = args
    import numpy as np

    from oracle import qsp_oracle as O

    if WORKLOAD == "heisenberg" and L == 32:
        z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                 "heisenberg_L32.npz"))
        b, m = z["bonds"], z["mps"]
        raw = [m[i, : b[i], :, : b[i + 1]].copy() for i in range(len(b) - 1)]
    elif WORKLOAD == "aklt":
        from qsp_b200.misc_states import make_aklt_mps, make_splitted_mps

        raw = O.to_lpr(make_splitted_mps(make_aklt_mps(L // 2)[0]), "lrp")
    else:
        raw = O.random_mps(L, chi, seed)
    t = O.prepare_target(raw, "lpr")
    t0 = time.time()
    O.sequential_unitary_circuit(t, layers, do_compression=True, max_bond_dim=chi, completion="lapack")
    return time.time() - t0


WORKLOAD = sys.argv[7] if len(sys.argv) > 7 else "random"


def main():
    L, chi, layers, n_inst, cores, reps = (int(x) for x in sys.argv[1:7])
    jobs = [(L, chi, layers, 1234 + i) for i in range(n_inst)]
    ctx = mp.get_context("fork")
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
