#!/usr/bin/env python
"""profiles/r2_traffic.json from an ncu launch list: DRAM bytes per launch of the hot kernels of ONE bench step.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --kernel-name regex:"khop_|tree_rows|rows_|expand_level|roots_assign|lid_clear|level_snapshot|batch_gather|linear_tf32|halo|stage_claim" \
        -c 400 --csv --log-file gpurun_out/r2_traffic.csv \
        python bench.py --steps 1 --warmup 1 --streams 1 --no-e2e --no-cpu-baseline --no-full-graph
    python scripts/ncu_traffic.py gpurun_out/r2_traffic.csv profiles/r2_traffic.json

bench.py reports these numbers as `roofline.traffic` only while the CUDA sources still hash to `source_digest` (bench.source_digest):
a kernel change invalidates the record instead of letting it go stale."""
import csv
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

PHASES = [  # (phase, kernel-name regex); first match wins
    ("sample", r"khop_"),
    ("gather_l1", r"batch_gather_(async_)?kernel<1, 8>|batch_gather_parts|batch_gather_finish"),
    ("gather_deep", r"batch_gather_"),
    ("gemm", r"linear_tf32x3_kernel"),
    ("halo_stage", r"halo_stage|stage_claim"),
    ("collate", r"tree_rows|rows_|expand_level|roots_assign|lid_clear|level_snapshot"),
]


def main(src, dst):
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    launches = {}
    order = []
    for row in csv.DictReader(lines):
        i = int(row["ID"])
        if i not in launches:
            launches[i] = {"name": row["Kernel Name"]}
            order.append(i)
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3,
                 "msecond": 1e3}.get(unit, 1)
        launches[i][row["Metric Name"]] = v * scale
    seq = [launches[i] for i in order]
    # a step starts at every other khop_tile launch (hop 1, hop 2)
    tiles = [k for k, l in enumerate(seq) if "khop_tile_kernel" in l["name"]]
    starts = tiles[0::2]
    if len(starts) < 2:
        raise SystemExit("need at least two steps in the capture")
    step = seq[starts[1]:(starts[2] if len(starts) > 2 else len(seq))]  # the timed step (after one warm-up)
    out = {}
    gemm_seen = 0
    gather_parts_owner = "gather_l1"
    for l in step:
        name = l["name"]
        phase = next((p for p, rx in PHASES if re.search(rx, name)), None)
        if phase is None:
            continue
        if phase == "gemm":
            phase = "gemm_l1" if gemm_seen == 0 else "gemm_deep"
            gemm_seen += 1
        if phase in ("gather_l1", "gather_deep"):
            if "parts" in name or "finish" in name:
                phase = gather_parts_owner
            else:
                phase = "gather_l1" if gemm_seen == 0 else "gather_deep"
                gather_parts_owner = phase
        rec = out.setdefault(phase, {"dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "time_us": 0.0, "launches": 0, "kernels": []})
        rec["dram_read_bytes"] += l.get("dram__bytes_read.sum", 0.0)
        rec["dram_write_bytes"] += l.get("dram__bytes_write.sum", 0.0)
        rec["time_us"] += l.get("gpu__time_duration.sum", 0.0)
        rec["launches"] += 1
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("gigl::", "")
        if short not in rec["kernels"]:
            rec["kernels"].append(short)
    for rec in out.values():
        rec["dram_bytes"] = rec["dram_read_bytes"] + rec["dram_write_bytes"]
    if "gather_l1" in out and "gemm_l1" in out:
        out["gather_spmm_l1"] = {k: out["gather_l1"][k] + out["gemm_l1"][k] for k in ("dram_read_bytes", "dram_write_bytes", "dram_bytes", "time_us")}
    import bench

    out["source_digest"] = bench.source_digest()
    out["how"] = ("ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on "
                  "`bench.py --steps 1 --warmup 1 --streams 1`; the second step of the capture; times are ncu's (cold cache, serialised)")
    out["csv"] = os.path.basename(src)
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: (v if not isinstance(v, dict) else {a: b for a, b in v.items() if a != "kernels"}) for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
