#!/usr/bin/env python3
"""Benchmark of the sequential MPS-encoding path (BASELINE.json metric: sequential-prep
MPS/sec at L=100, chi=64, 4 layers, do_compression=True, max_bond_dim=chi).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--L 100 --chi 64 --layers 4 --batch 296]

A step = one pass of the hot path (the whole layer loop incl. the encoded-state overlaps)
over one batch of `batch` synthetic random MPS per GPU.  Prints ONE JSON line (rank 0).
Under torchrun every rank owns its own instances (weak scaling, no data-path collective
except the final all-gather of overlaps + unitaries the path's callers need).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from qsp_b200 import flop_model  # noqa: E402

METRIC = "sequential_prep_mps_per_sec"
_T0 = time.time()


def log(msg):
    print(f"[bench {time.time() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)

UNIT = "MPS/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--L", type=int, default=100)
    ap.add_argument("--chi", type=int, default=64)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--batch", type=int, default=592, help="instances per GPU per step (4 per SM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--second-pass", action="store_true")
    ap.add_argument("--workload", default="random", choices=["random", "heisenberg", "aklt"],
                    help="random: SURVEY 8d synthetic MPS; heisenberg: the reference's L=32 ground state "
                         "(tests/golden/heisenberg_L32.npz) replicated over the batch (cfg 3a: --layers 6 --chi 32); "
                         "aklt: make_aklt_mps(L/2) + make_splitted_mps replicated (cfg 3b: --L 100 --layers 6 --chi 32)")
    return ap.parse_args()


def workload_name(a):
    if a.workload == "heisenberg":
        return f"heisenberg_L32_pickle_replicated_layers{a.layers}_compress_maxbond{a.chi}"
    if a.workload == "aklt":
        return f"aklt_L{a.L}_qubits_replicated_layers{a.layers}_compress_maxbond{a.chi}"
    return f"random_mps_L{a.L}_chi{a.chi}_d2_layers{a.layers}_compress_maxbond{a.chi}"


def heisenberg_batch(torch, B, chi, dev):
    """cfg 3a: the Heisenberg L=32 ground state shipped with the reference, replicated B times."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "heisenberg_L32.npz"))
    m, b = z["mps"].astype(np.complex128), z["bonds"].astype(np.int32)
    D = max(int(b.max()), chi)
    pad = np.zeros((m.shape[0], D, 2, D), np.complex128)
    pad[:, : m.shape[1], :, : m.shape[3]] = m
    x = torch.from_numpy(pad).to(dev).unsqueeze(0).repeat(B, 1, 1, 1, 1).contiguous()
    bonds = torch.from_numpy(b).to(dev).unsqueeze(0).repeat(B, 1).contiguous()
    return x, bonds


# --------------------------------------------------------------------------------------
# CPU arm: the oracle (numpy restatement of the reference path; quimb is not installable)
# --------------------------------------------------------------------------------------
def cpu_sample(L, chi, layers_sample, n_inst, cores, reps=1, timeout=420, workload="random"):
    """n_inst instances x layers_sample layers on `cores` single-threaded oracle workers, in a
    separate interpreter (BLAS threads pinned to 1 through the environment) -> list of wall seconds."""
    env = dict(os.environ)
    for k in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        env[k] = "1"
    env["CUDA_VISIBLE_DEVICES"] = ""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "cpu_bench_worker.py"), str(L), str(chi), str(layers_sample),
                          str(n_inst), str(cores), str(reps), workload], env=env, capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError(out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def cpu_cores():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        n = os.cpu_count() or 1
    return max(1, min(n, 16))


def cpu_baseline(a):
    """Bounded sample: one layer of the workload on `cores` instances (one per worker),
    extrapolated linearly to `layers` layers (with compression on every layer costs the same)."""
    cores = cpu_cores()
    log(f"cpu_baseline: {cores} oracle workers, 1 layer each ...")
    try:
        wall = cpu_sample(a.L, a.chi, 1, cores, cores, workload=a.workload)[0]
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"failed: {str(e)[:200]}"}
    v = cores / (wall * a.layers)
    return {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} instances x 1 of {a.layers} layers of {workload_name(a)} "
                      f"(numpy/scipy restatement of the reference path, one single-threaded worker per core, "
                      f"{wall:.1f}s wall; per-layer cost extrapolated x{a.layers})"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_cores()
    steps = max(1, min(a.steps, 3))
    if a.workload == "heisenberg":
        a.L = 32
    walls = cpu_sample(a.L, a.chi, 1, cores, cores, reps=steps, timeout=900, workload=a.workload)
    t = float(np.mean(walls))
    v = cores / (t * a.layers)
    sample = (f"each step = {cores} instances x 1 of {a.layers} layers of {workload_name(a)} on {cores} "
              f"single-threaded workers (numpy/scipy restatement; quimb, the reference's engine, cannot be installed); "
              f"MPS/s = instances / (wall x {a.layers})")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": {"workload": workload_name(a), "L": a.L, "chi": a.chi, "layers": a.layers},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def aklt_batch(torch, B, L, chi, dev):
    """cfg 3b: AKLT chain of L/2 spin-1 sites split into L qubits (qsp_b200.misc_states), replicated."""
    from qsp_b200.engine import pack_mps_list
    from qsp_b200.misc_states import make_aklt_mps, make_splitted_mps
    from qsp_b200.mps import to_lpr

    tens, _ = make_aklt_mps(L // 2)
    arr, b = pack_mps_list([to_lpr(make_splitted_mps(tens), "lrp")], chi)
    x = torch.from_numpy(arr).to(dev).repeat(B, 1, 1, 1, 1).contiguous()
    bonds = torch.from_numpy(b).to(dev).repeat(B, 1).contiguous()
    return x, bonds


def synth_batch(torch, B, L, chi, dev, seed):
    """Random complex MPS with bond profile min(chi, 2^i, 2^(L-i)), entries N(0,1)+iN(0,1)
    (SURVEY.md section 8d), generated on the device; padding zero."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    prof = flop_model.bond_profile(L, chi)
    x = torch.zeros((B, L, chi, 2, chi), dtype=torch.complex128, device=dev)
    for i in range(L):
        l, r = prof[i], prof[i + 1]
        re = torch.randn((B, l, 2, r), generator=g, device=dev, dtype=torch.float64)
        im = torch.randn((B, l, 2, r), generator=g, device=dev, dtype=torch.float64)
        x[:, i, :l, :, :r] = torch.complex(re, im) / (2.0 * max(l, r)) ** 0.5
    bonds = torch.tensor(prof, dtype=torch.int32, device=dev).repeat(B, 1).contiguous()
    return x, bonds


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    import torch
    import torch.distributed as dist

    from qsp_b200 import _runtime
    from qsp_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _runtime.lib()
    eng = Engine(str(dev))
    B, L, chi, layers = a.batch, a.L, a.chi, a.layers

    # ---- inputs resident in HBM (batch * 32 chi^2 L bytes >> 126 MB L2 at the default shape) ----
    if a.workload == "heisenberg":
        raw, bonds0 = heisenberg_batch(torch, B, chi, dev)
        L = a.L = raw.shape[1]
    elif a.workload == "aklt":
        raw, bonds0 = aklt_batch(torch, B, L, chi, dev)
    else:
        raw, bonds0 = synth_batch(torch, B, L, chi, dev, 1234 + rank)
    tgt, tb = raw.clone(), bonds0.clone()
    st = eng.prepare_target(tgt, tb)
    torch.cuda.synchronize()
    assert int(st.max().item()) == 0, "prepare_target status"

    def step():
        res = eng.seq_prepare(tgt, tb, layers, max_bond=chi, second_pass=a.second_pass)
        if world > 1:  # the only collective of the path: overlaps + unitaries of every rank
            ov = [torch.empty_like(res.overlaps) for _ in range(world)]
            dist.all_gather(ov, res.overlaps)
            gg = [torch.empty_like(res.G) for _ in range(world)]
            dist.all_gather(gg, res.G)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    log(f"inputs ready: B={B} L={L} chi={chi} layers={layers} world={world}")
    for _ in range(a.warmup):
        res = step()
    barrier()
    log("warm-up done")
    assert int(res.status.max().item()) == 0, f"device status {res.status.unique().tolist()}"
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        res = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    log(f"timed region done: {ms / a.steps:.1f} ms/step")
    launches = lib.launch_count() - n0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / a.steps
    value = world * B / (ms_per_step * 1e-3)
    ov_last = res.overlaps[layers - 1]
    assert bool(torch.isfinite(ov_last.abs()).all().item())

    # ---- end-to-end through the public batched API with HOST buffers -----------------------
    e2e = None
    if not a.no_e2e:
        from qsp_b200.tsp import BatchedMPSPreparation

        host_raw = torch.empty(raw.shape, dtype=raw.dtype, pin_memory=True)
        host_raw.copy_(raw)
        host_b = torch.empty(bonds0.shape, dtype=bonds0.dtype, pin_memory=True)
        host_b.copy_(bonds0)
        outs = {}

        def e2e_step():
            d_raw = host_raw.to(dev, non_blocking=True)
            d_b = host_b.to(dev, non_blocking=True)
            prep = BatchedMPSPreparation(packed=(d_raw, d_b), device=str(dev))
            _, r = prep.sequential_unitary_circuit(layers, do_compression=True, max_bond_dim=chi)
            for k in ("overlaps", "dis_overlaps", "G", "kind", "status"):
                src = getattr(r, k)
                if k not in outs:
                    outs[k] = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
                outs[k].copy_(src, non_blocking=True)
            torch.cuda.synchronize()

        e2e_step()
        barrier()
        n_e2e = max(1, min(a.steps, 2))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world * B / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(host_raw.numel() * 16 + host_b.numel() * 4),
               "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in outs.values()))}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    log(f"e2e done: {e2e}")
    # ---- roofline of the dominant kernel (k_right_compress, 2chi -> chi), timed live --------
    roof = kernel_roofline(torch, eng, lib, tgt, tb, a, dev)
    log(f"kernel roofline done: {roof['kernel_ms']:.1f} ms, frac {roof['frac']}")
    cpu = None if a.no_cpu_baseline else cpu_baseline(a)
    log("cpu baseline done")
    fm = flop_model.flops_per_mps(L, chi, layers)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "c128", "data": "synthetic",
        "config": {"workload": workload_name(a), "L": L, "chi": chi, "d": 2, "layers": layers,
                   "do_compression": True, "max_bond_dim": chi, "batch_per_gpu": B, "parallelism": f"batch-shard x{world}",
                   "l2": f"inputs {B * L * chi * chi * 32 / 1e9:.2f} GB per GPU > 126 MB L2 (no flush needed)",
                   "F_ref_flops_per_mps": fm["total"], "B_ref_bytes_per_mps": flop_model.bytes_per_mps(L, chi, layers)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _runtime_dev_lib():
    from qsp_b200 import _runtime

    return _runtime.dev_lib()  # probes live in libqsp_b200_dev.so (include/qsp_b200_dev.h), not in the product library


def fp64_peak(torch, lib, dev, sms, tensor=False):
    """FP64 throughput measured in-run (TFLOP/s): MEASURED_PEAKS.json has no FP64 entry.
    tensor=False: DFMA (the pipe the kernels use); tensor=True: warp-level DMMA (mma.sync m8n8k4.f64)."""
    import ctypes

    out = torch.zeros(8, dtype=torch.float64, device=dev)
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    blocks, threads, iters = sms * 8, 256, 20000
    lib = _runtime_dev_lib()
    probe = lib.qsp_dmma_probe if tensor else lib.qsp_dfma_probe
    flops = blocks * (threads // 32) * iters * 8 * 512.0 if tensor else blocks * threads * iters * 32.0
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        probe(ctypes.c_void_p(out.data_ptr()), blocks, threads, iters, s)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def kernel_roofline(torch, eng, lib, tgt, tb, a, dev):
    """One layer replayed through the unit-level entry points, each kernel timed with CUDA
    events on the launching stream; the dominant one is the truncating sweep 2chi -> chi."""
    import ctypes

    B, L, chi = tgt.shape[0], tgt.shape[1], a.chi
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())  # noqa: E731
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    psi, pb = tgt.clone(), tb.clone()
    st = torch.zeros(B, dtype=torch.int32, device=dev)
    d2 = torch.empty((B, L, 2, 2, 2), dtype=torch.complex128, device=dev)
    d2b = torch.empty((B, L + 1), dtype=torch.int32, device=dev)
    G = torch.empty((B, L, 4, 4), dtype=torch.complex128, device=dev)
    kind = torch.empty((B, L), dtype=torch.int32, device=dev)
    out = torch.empty_like(psi)
    ob = torch.empty_like(pb)
    amp = torch.empty(B, dtype=torch.complex128, device=dev)
    nb = max(lib.qsp_right_compress_workspace_bytes(B, L, chi, 2), lib.qsp_apply_inverse_and_compress_workspace_bytes(B, L, chi),
             lib.qsp_left_canonize_workspace_bytes(B, L, chi))
    work = torch.empty(nb, dtype=torch.uint8, device=dev)
    times = {}

    def timed(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times[name] = e0.elapsed_time(e1)

    timed("k_left_canon", lambda: lib.qsp_left_canonize(p(psi), p(pb), B, L, chi, p(st), p(work), nb, s))
    timed("k_right_compress_d2", lambda: lib.qsp_right_compress(p(psi), p(pb), chi, p(d2), p(d2b), 2, B, L, 2, 1e-10, 1,
                                                                 None, 0, None, p(st), p(work), nb, s))
    timed("k_unitaries", lambda: lib.qsp_generate_unitaries(p(d2), p(d2b), B, L, p(G), p(kind), p(st), s))
    reps = []
    for _ in range(2):
        timed("k_right_compress", lambda: lib.qsp_apply_inverse_and_compress(p(psi), p(pb), p(G), p(kind), p(out), p(ob), chi,
                                                                              B, L, chi, 1e-10, 1, p(amp), p(st), p(work), nb, s))
        reps.append(times["k_right_compress"])
    dur_ms = float(np.mean(reps))
    fm = flop_model.flops_per_mps(L, chi, 1)
    # algorithmic flops of what this launch replaces: inverse layer (S2) + re-truncation (S3) + trailing compress (S4)
    alg = B * (fm["S2"] + fm["S3"] + fm["S4"])
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak = fp64_peak(torch, lib, dev, sms)
    achieved = alg / (dur_ms * 1e-3) / 1e12
    traffic, traffic_src, ncu_fp64 = None, None, None
    try:  # dram__bytes_read+write of this kernel from the committed ncu capture, scaled to this launch's batch
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("config") == f"L{L} chi{chi} layers4" or tj.get("config") == f"L{L} chi{chi} layers{a.layers}":
            traffic = tj["dram_bytes_per_instance"] * B
            traffic_src = f"ncu dram__bytes_read.sum+dram__bytes_write.sum at batch {tj['batch']} ({tj['source']}), scaled to batch {B}"
            ncu_fp64 = tj.get("fp64_pipe_busy_pct_ncu")
    except Exception:  # noqa: BLE001
        pass
    layer_ms = sum(v for k, v in times.items())
    return {"bound": "fp64", "kernel": "k_right_compress(fused inverse layer + 2chi->chi)", "achieved": achieved, "peak": peak,
            "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "DFMA probe measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
            "fp64_tensor_peak_dmma_probe": fp64_peak(torch, lib, dev, sms, tensor=True),
            "note": "achieved counts the REFERENCE's flops this launch replaces (S2+S3+S4 of SURVEY 8d: inverse layer with "
                    "SVD splits, two 2chi QR sweeps, two SVD sweeps); the device reaches the same state with one Jacobi sweep "
                    "(DESIGN.md I1-I3), so frac is an algorithmic speed figure, not pipe utilisation",
            "fp64_pipe_busy_pct_ncu": ncu_fp64,
            "algorithmic_flops_per_launch": alg, "kernel_ms": dur_ms,
            "layer_kernel_ms": {k: round(v, 3) for k, v in times.items()},
            "kernel_share_of_layer": dur_ms / layer_ms if layer_ms else None,
            "hbm_view": {"achieved_GBps": B * flop_model.bytes_per_mps(L, chi, 1) / (layer_ms * 1e-3) / 1e9,
                         "peak_GBps": _measured_hbm()}}


def _measured_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f).get("hbm_gbs")
    except Exception:  # noqa: BLE001
        return 6650.0


if __name__ == "__main__":
    main()
