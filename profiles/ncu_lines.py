#!/usr/bin/env python
"""Per-source-line totals (instructions executed, stall samples) from an ncu report with -lineinfo / --import-source on.
Usage: python profiles/ncu_lines.py <report.ncu-rep> [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] != "" and hdr:
        try:
            ln = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr[2:], r[2:]))
        def num(k):
            try: return float(d.get(k, "0"))
            except ValueError: return 0.0
        lines.append((cur_file, ln, r[1].strip()[:90], num("Instructions Executed"), num("Warp Stall Sampling (All Samples)"), num("L1 Wavefronts Shared"), num("L1 Tag Requests Global")))
ti = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print(f"total inst {ti:.3e}  total samples {ts:.0f}")
print("by instructions:")
for l in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{l[0]}:{l[1]:4d} inst {100*l[3]/ti:5.1f}%  stall {100*l[4]/ts:5.1f}%  smem_wf {l[5]:.2e} tag {l[6]:.2e} | {l[2]}")
print("by stall samples:")
for l in sorted(lines, key=lambda x: -x[4])[:top // 2]:
    print(f"{l[0]}:{l[1]:4d} inst {100*l[3]/ti:5.1f}%  stall {100*l[4]/ts:5.1f}% | {l[2]}")
