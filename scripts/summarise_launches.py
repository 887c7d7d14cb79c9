"""Turn an ncu launch list (csv with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per
launch) into profiles/<tag>_launches.md and profiles/<tag>_traffic.json.

    python scripts/summarise_launches.py gpurun_out/launches.csv r1 148
"""
import csv
import json
import os
import sys
from collections import OrderedDict, defaultdict

src, tag, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")

rows = []
with open(src) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)

launches = OrderedDict()
for r in rows:
    k = int(r["ID"])
    e = launches.setdefault(k, {"name": r["Kernel Name"].split("(")[0], "block": r["Block Size"], "grid": r["Grid Size"],
                                "ms": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r["Metric Value"].replace(",", ""))
    m, u = r["Metric Name"], r["Metric Unit"]
    if m == "gpu__time_duration.sum":
        e["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u]
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        e["rd" if "read" in m else "wr"] = v

agg = defaultdict(lambda: {"n": 0, "ms": 0.0, "bytes": 0.0})
for e in launches.values():
    a = agg[(e["name"], e["block"])]
    a["n"] += 1
    a["ms"] += e["ms"]
    a["bytes"] += e["rd"] + e["wr"]
total = sum(a["ms"] for a in agg.values())
order = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])

out = [f"# ncu launch list, {tag} (own kernels only)", "",
       "command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
       f"-k regex:^k_ -c 400 --csv python bench.py --batch-total {2 * batch} --steps 1 --warmup 1 --no-cpu-baseline --no-e2e`",
       f"(after the same command exited 0 without ncu; B200, L=100 chi=64 4 layers, two chunks of {batch} instances per step; "
       "per-launch times are cold-cache and serialised - compare SHARES, not absolutes)", "",
       f"{len(launches)} launches captured (warm-up step + timed step + the kernel-roofline replay of one layer), "
       f"total {total:.1f} ms", "",
       "| kernel | block | launches | total ms | share | DRAM bytes / launch |", "|---|---|---|---|---|---|"]
for (name, block), a in order:
    out.append(f"| {name} | {block} | {a['n']} | {a['ms']:.2f} | {100 * a['ms'] / total:.1f}% | "
               f"{a['bytes'] / a['n'] / 1e6:.1f} MB |")

# dominant kernel: the fused gate-apply-and-split launches = k_right_compress at 512 threads with the largest time
dom = [e for e in launches.values() if e["name"] == "k_right_compress" and e["block"].startswith("(512")]
big = max(e["ms"] for e in dom)
fused = [e for e in dom if e["ms"] > 0.6 * big]
mean_ms = sum(e["ms"] for e in fused) / len(fused)
mean_b = sum(e["rd"] + e["wr"] for e in fused) / len(fused)
share = sum(e["ms"] for e in fused) / total
out += ["", f"Dominant kernel `k_right_compress` (512 threads, fused inverse layer + 2chi -> chi): {len(fused)} launches, "
        f"mean {mean_ms:.1f} ms, DRAM traffic {mean_b / 1e9:.3f} GB per launch at B={batch} "
        f"(= {mean_b / batch / 1e6:.1f} MB per MPS-layer); {100 * share:.1f}% of all own-kernel time here "
        "(the remaining 512-thread k_right_compress launches are prepare_target's chi -> chi sweep).", "",
        "First 30 launches in order:", "", "| # | kernel | grid | block | ms | DRAM MB |", "|---|---|---|---|---|---|"]
for k, e in list(launches.items())[:30]:
    out.append(f"| {k} | {e['name']} | {e['grid']} | {e['block']} | {e['ms']:.3f} | {(e['rd'] + e['wr']) / 1e6:.1f} |")
open(os.path.join(root, "profiles", f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")

tj = os.path.join(root, "profiles", "traffic.json")  # read by bench.py (roofline.traffic)
old = json.load(open(tj)) if os.path.exists(tj) else {}
old.update({"kernel": "k_right_compress(fused inverse layer + 2chi->chi)", "config": "L100 chi64", "batch": batch,
            "dram_bytes_per_launch": mean_b, "dram_bytes_per_instance": mean_b / batch,
            "source": f"profiles/{tag}_launches.csv", "share_of_own_kernel_time_ncu": share})
json.dump(old, open(tj, "w"), indent=1)
print("\n".join(out[:24]))
