#!/usr/bin/env python
"""Aggregate an ncu report's source page by CUDA source line: tools/ncu_lines.py report.ncu-rep kernel_regex [topN]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = []
H = None
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        H = r
        continue
    if H and len(r) > 8 and r[0] not in ("", "Line No"):
        try:
            lines.append((int(r[0]), r[1], float(r[H.index("# Samples")] or 0), float(r[H.index("Instructions Executed")] or 0)))
        except ValueError:
            pass
ts = sum(l[2] for l in lines) or 1
ti = sum(l[3] for l in lines) or 1
print(f"kernel {kern}: {len(lines)} source lines, {ts:.0f} samples, {ti:.3e} warp instructions")
for l in sorted(lines, key=lambda x: -x[2])[:top]:
    print(f"{l[2] / ts:6.1%} smp {l[3] / ti:6.1%} inst | L{l[0]:4d} | {l[1].strip()[:120]}")
