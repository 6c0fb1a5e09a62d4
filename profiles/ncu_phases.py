#!/usr/bin/env python
"""Instruction / stall totals per kernel PHASE from an ncu report (-lineinfo, --import-source on).  Phases are delimited by
marker comments found in the source text embedded in the report itself, so line drift does not matter.
Usage: python profiles/ncu_phases.py <report.ncu-rep> fwd|bwd"""
import csv, io, subprocess, sys
rep, which = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
MARK = {
    "fwd": [("render_fwd_group_kernel(const GridP", "setup"), ("const bool mine = marching", "march+density"),
            ("const unsigned act = __ballot_sync", "publish"), ("if (m < total && role_ok) {", "group loop"),
            ("const float4 raw = ", "composite"), ("if (!alive) return;", "epilogue"), ("struct RayGrad", "other")],
    "bwd": [("render_bwd_coop_kernel(const GridP", "setup"), ("if (alive && i >= s.i_lo", "per-lane chain maths"),
            ("float* Prow = sm.P", "publish"), ("peers = __match_any_sync", "match+leaders"),
            ("while (mm) {", "member sweep"), ("const int* VL = ", "scatter"), ("mark_touched_kernel", "other")],
}[which]
cur = None; hdr = None; lines = []; func = ""
want = "render_fwd" if which == "fwd" else "render_bwd"
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name": func = r[1]; continue
    if want not in func: continue
    if r[0] != "" and hdr:
        try: ln = int(r[0])
        except ValueError: continue
        d = dict(zip(hdr[2:], r[2:]))
        def num(k):
            try: return float(d.get(k, "0"))
            except ValueError: return 0.0
        lines.append((cur, ln, r[1], num("Instructions Executed"), num("Warp Stall Sampling (All Samples)")))
# marker line numbers inside r3d_render.cu (the kernel under test starts at its first marker)
src = sorted([(l[1], l[2]) for l in lines if l[0] == "r3d_render.cu"])
bounds = []
start_ln = None
for text, name in MARK:
    for ln, t in src:
        if text in t and (start_ln is None or ln >= start_ln):
            bounds.append((ln, name)); start_ln = ln if start_ln is None else start_ln; break
acc = {}
for f, ln, t, n, s in lines:
    if f == "r3d_render.cu":
        ph = "setup"
        for b, name in bounds:
            if ln >= b: ph = name
    elif f == "r3d_device.cuh":
        ph = "device.cuh: " + ("depth" if "DepthMarch" in t or "base(" in t or "jitter" in t or "mix32" in t or 120 < ln < 190 else "cell/density/misc")
    else:
        ph = "intrinsics (" + f + ")"
    a = acc.setdefault(ph, [0, 0]); a[0] += n; a[1] += s
ti = sum(v[0] for v in acc.values()); ts = sum(v[1] for v in acc.values())
for k, v in sorted(acc.items(), key=lambda x: -x[1][0]):
    print(f"{k:44s} inst {v[0]:.3e} ({100*v[0]/ti:5.1f}%)   stall samples {100*v[1]/ts:5.1f}%")
