"""bench.py contract pieces that run without a GPU: the reference (CPU) arm prints one JSON line
with the agreed keys, and the CPU worker returns wall times."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--L", "8", "--chi", "4",
                          "--layers", "2", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "MPS/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("random_mps_L8_chi4")


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--L", "8",
                          "--chi", "4", "--layers", "2", "--steps", "1"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_default_workload_is_the_headline_config():
    sys.path.insert(0, ROOT)
    import bench

    sys.argv = ["bench.py"]
    a = bench.parse()
    assert (a.L, a.chi, a.layers, a.gpus, a.impl) == (100, 64, 4, 1, "b200") and a.warmup >= 3
    assert bench.workload_name(a) == "random_mps_L100_chi64_d2_layers4_compress_maxbond64"
    # BASELINE.json configs[3] as named: 4096 instances sharded over the ranks (strong scaling) unless --batch is given
    assert a.batch_total == 4096 and a.batch is None


def test_parity_worker_checks_sampled_instances(tmp_path):
    """the 'parity alongside' leg of bench.py: sampled device results (here: the emulated kernels) teacher-forced
    through the oracle by oracle/cpu_bench_worker.py"""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import HostDriver, pack_batch
    from oracle import qsp_oracle as O

    tg = []
    for s in range(2):
        t = O.prepare_target(O.random_mps(8, 4, 900 + s), "lpr")
        O.right_canonize(t, normalize=True)
        tg.append(t)
    a, b = pack_batch(tg, 4)
    r = HostDriver().seq_prepare(a, b, 3, max_bond=4)
    path = str(tmp_path / "sample.npz")
    np.savez(path, mps=np.nan_to_num(a), bonds=b, G=r["G"], kind=r["kind"], dis=r["dis"], overlaps=r["overlaps"],
             d2_svals=r["d2_svals"], max_bond=4)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "cpu_bench_worker.py"), "parity", path, "2"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["instances"] == 2 and res["max_abs_diff_dis_overlaps"] < 1e-12 and res["max_abs_diff_overlaps"] < 1e-12
    assert res["max_rel_diff_d2_svals"] < 1e-11
