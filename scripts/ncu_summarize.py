"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/ncu_summarize.py <round-tag> <launches.csv> <full.ncu-rep>

Writes profiles/<tag>_launches.md (every kernel's share of the profiled command, from the
`--metrics gpu__time_duration.sum` pass), profiles/<tag>_kernels.md (the `--set full` metrics of the
hot kernels).  The DRAM bytes bench.py reports as `roofline.traffic` come from scripts/ncu_traffic.py (profiles/r2_traffic.json)."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches_csv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)


def to_us(v, u):
    v = float(v.replace(",", ""))
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(u, v)


with open(launches_csv) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rd:
    if len(r) <= vi:
        continue
    try:
        t = to_us(r[vi], r[ui])
    except ValueError:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
ours = dict(agg)  # the capture is already restricted to the library's kernels (--kernel-name regex)
tot = sum(v[1] for v in ours.values())
with open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: launch list of `bench.py --steps 1 --warmup 1 --streams 1` under `ncu --metrics gpu__time_duration.sum --clock-control none`\n\n")
    f.write("The library's step kernels (--kernel-name regex, scripts/final_profiles.sh); torch kernels of the synthetic-input generation and the "
            "one-time CUB sorts of the graph build are left out. Times are cold-cache and serialised: compare SHARES, not absolutes. The capture "
            "covers the warm-up step, the timed step and the untimed recount pass (sampling + collation + export only).\n\n")
    f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k[:100]}` | {n} | {t:.1f} | {100 * t / tot:.1f}% |\n")
    f.write(f"\nTotal {tot:.1f} us.\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
idx = [(w, hdr.index(w)) for w in want if w in hdr]
ni = hdr.index("Kernel Name")


def gb(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}.get(u, 1.0)


with open(os.path.join(ROOT, "profiles", f"{tag}_kernels.md"), "w") as f:
    f.write(f"# {tag}: `ncu --set full --clock-control none` of the hot kernels inside `bench.py` (products-like, B=65536, fanout [15,10])\n\n")
    f.write("| kernel | " + " | ".join(w for w, _ in idx) + " |\n|---|" + "---:|" * len(idx) + "\n")
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ni]).replace("void ", "")
        f.write(f"| `{name[:60]}` | " + " | ".join(f"{r[i]} {units[i]}" for _, i in idx) + " |\n")
print("wrote", f"profiles/{tag}_launches.md", f"profiles/{tag}_kernels.md")
