"""N > 1 path on the CPU: world_size-2 gloo ranks shard the instances, each runs the layer
loop on its block (emulated kernels stand in for the GPU here) and the all-gather returns
the same results as the single-process run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import HostDriver, pack_batch
from oracle import qsp_oracle as O

from qsp_b200.distributed import gather_results, shard_range

B, L, CHI, LAYERS = 5, 6, 4, 2


def _targets():
    tg = []
    for s in range(B):
        t = O.prepare_target(O.random_mps(L, CHI, 300 + s), "lpr")
        O.right_canonize(t, normalize=True)
        tg.append(t)
    return tg


def _run(tg):
    drv = HostDriver()
    a, b = pack_batch(tg, CHI)
    r = drv.seq_prepare(a, b, LAYERS, max_bond=CHI, want_final=False)
    return {"overlaps": torch.from_numpy(r["overlaps"]), "dis_overlaps": torch.from_numpy(r["dis"]),
            "G": torch.from_numpy(r["G"]), "kind": torch.from_numpy(r["kind"]),
            "status": torch.from_numpy(r["status"])}


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tg = _targets()
    b0, b1 = shard_range(B, rank, world)
    full = gather_results(_run(tg[b0:b1]), B)
    if rank == 0:
        q.put({k: v.numpy() for k, v in full.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    for n in (1, 5, 8, 4096):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_gloo_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = _run(_targets())
    for k, v in ref.items():
        x, y = got[k], v.numpy()
        assert x.shape == y.shape, k
        assert np.array_equal(np.isnan(x.view(np.float64) if np.iscomplexobj(x) else x),
                              np.isnan(y.view(np.float64) if np.iscomplexobj(y) else y))
        assert np.allclose(np.nan_to_num(x), np.nan_to_num(y), rtol=0, atol=0), k
