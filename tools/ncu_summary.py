#!/usr/bin/env python3
"""Turns what ncu brought back in gpurun_out/ into the markdown tables committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv              # per-kernel totals and shares of a launch list
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [--json out.json]  # one row per profiled launch (--set full)

`full` shells out to `ncu -i … --page raw --csv` (ncu runs here without a GPU) and keeps the metrics the roofline argument
needs: duration, DRAM bytes, DRAM / SM throughput, issue-active, achieved occupancy, registers, executed instructions, hit
rates. With --json it also writes {kernel: {"dram_bytes": read+write per launch, "duration_ns": …}} averaged over the
profiled launches of each kernel: bench.py reads that file for `roofline.traffic`.
"""
import argparse
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "inst (warp)"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]


def short(name: str) -> str:
    name = re.sub(r"^(void )?dfx::", "", name)
    return re.sub(r"\(.*$", "", name)


def read_csv_rows(text: str):
    lines = text.splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    return list(csv.reader(io.StringIO("\n".join(lines[start:]))))


def cmd_launches(path: str):
    rows = read_csv_rows(open(path).read())
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iu], 1)
        tot[short(r[ik])] += v
        cnt[short(r[ik])] += 1
    total = sum(tot.values())
    print("| kernel | launches | total (ns) | share |\n|---|---:|---:|---:|")
    for k, v in tot.most_common():
        print(f"| {k} | {cnt[k]} | {int(v)} | {v / total:.3f} |")


def to_base(value: str, unit: str) -> float:
    v = float(value.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "usecond": 1e3, "msecond": 1e6, "nsecond": 1, "second": 1e9}
    return v * scale.get(unit, 1)


def cmd_full(path: str, json_out: str | None):
    text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = read_csv_rows(text)
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    cols = [(hdr.index(m), lbl) for m, lbl in KEEP if m in hdr]
    print("| kernel | " + " | ".join(lbl for _, lbl in cols) + " |\n|---|" + "---:|" * len(cols))
    agg = collections.defaultdict(lambda: {"dram_bytes": 0.0, "duration_ns": 0.0, "launches": 0})
    for r in rows[2:]:
        if len(r) <= ik:
            continue
        print(f"| {short(r[ik])} | " + " | ".join(f"{r[i]} {units[i]}".strip() for i, _ in cols) + " |")
        a = agg[short(r[ik])]
        a["dram_bytes"] += to_base(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")]) + \
            to_base(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        a["duration_ns"] += to_base(r[hdr.index("gpu__time_duration.sum")], units[hdr.index("gpu__time_duration.sum")])
        a["launches"] += 1
    if json_out:
        out = {k: {"dram_bytes": v["dram_bytes"] / v["launches"], "duration_ns": v["duration_ns"] / v["launches"], "launches": v["launches"]} for k, v in agg.items()}
        json.dump(out, open(json_out, "w"), indent=1, sort_keys=True)
        print(f"\nwrote {json_out}", file=sys.stderr)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["launches", "full"])
    ap.add_argument("path")
    ap.add_argument("--json")
    a = ap.parse_args()
    cmd_launches(a.path) if a.mode == "launches" else cmd_full(a.path, a.json)
