#!/usr/bin/env python
"""Tables for profiles/*_round_end.md from the artefacts of a GPU session (development aid).

  tools/round_end_report.py launches gpurun_out/launches.csv        -> launch-list table (kernel, launches, total ms, share)
  tools/round_end_report.py full gpurun_out/r2_final.ncu-rep        -> `ncu --set full` table of the captured kernels
  tools/round_end_report.py traffic gpurun_out/r2_final.ncu-rep N_SAMPLES N_IMAGES -> profiles/traffic.json content
"""
import csv
import json
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    if name.startswith("cub::") or "cub::" in name:
        m = re.search(r"(Device\w+Kernel|\w+Kernel)", name)
        return "cub::" + (m.group(1) if m else "kernel")
    return name


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    H = rows[0]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) * 1e-6
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.2f} | {a[1] / tot:.1%} |")
    print(f"\ntotal {tot:.1f} ms over {sum(a[0] for a in agg.values())} launches")


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H = rows[0]
    return [dict(zip(H, r)) for r in rows[2:]]


def fnum(d, k):
    try:
        return float(str(d.get(k, "0")).replace(",", "") or 0)
    except ValueError:
        return 0.0


def full(rep):
    print("| kernel | grid x block | time ms | DRAM read MB | DRAM write MB | SM % | issue active % | warps active % | tensor pipe active % | regs | inst / cycle / SM |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for d in raw(rep):
        nsm = fnum(d, "launch__sm_count") or 148
        print(f"| `{short(d['Kernel Name'])}` | {d['launch__grid_size']} x {d['launch__block_size']} | {fnum(d, 'gpu__time_duration.sum'):.3f} | "
              f"{fnum(d, 'dram__bytes_read.sum'):.1f} | {fnum(d, 'dram__bytes_write.sum'):.1f} | "
              f"{fnum(d, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {fnum(d, 'sm__inst_issued.avg.pct_of_peak_sustained_active') or fnum(d, 'smsp__issue_active.avg.pct'):.1f} | "
              f"{fnum(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {fnum(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{d.get('launch__registers_per_thread')} | {fnum(d, 'sm__inst_executed.sum.per_cycle_active') / nsm:.2f} |")
    print("\n(units as ncu prints them: time in the unit of gpu__time_duration.sum, DRAM in the unit of dram__bytes_*.sum)")


def traffic(rep, n_samples, n_images):
    per = {}
    for d in raw(rep):
        k = short(d["Kernel Name"])
        unit_r = 1.0
        b = fnum(d, "dram__bytes_read.sum") + fnum(d, "dram__bytes_write.sum")
        per.setdefault(k, [0.0, 0])
        per[k][0] += b
        per[k][1] += 1
    print(json.dumps({k: {"sum_as_printed": v[0], "launches": v[1]} for k, v in per.items()}, indent=1))
    print(f"# divide by {n_samples} samples / {n_images} images after converting the printed unit (see --page raw units row)")


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "full":
        full(sys.argv[2])
    elif cmd == "traffic":
        traffic(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
