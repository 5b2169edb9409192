#!/usr/bin/env python
"""Instruction / stall-sample shares of an .ncu-rep by source file and by function-sized regions of exec_docs*.cuh."""
import csv, subprocess, io, sys, re
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; agg = {}
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0].isdigit() and len(r) > 9 and r[2] == "-":
        g = lambda x: int(x) if x.isdigit() else 0
        a = agg.setdefault((cur, int(r[0])), [0, 0, 0])
        a[0] += g(r[4]); a[1] += g(r[7]); a[2] += g(r[8])
ti = sum(a[1] for a in agg.values()) or 1; ts = sum(a[0] for a in agg.values()) or 1
# regions = top-level __device__/__global__ functions found in the source files
import pathlib
root = pathlib.Path(__file__).resolve().parent.parent / "trinity_b200" / "csrc"
for f in sorted(set(k[0] for k in agg)):
    i = sum(a[1] for k, a in agg.items() if k[0] == f); s = sum(a[0] for k, a in agg.items() if k[0] == f)
    print(f"{f:32s} inst {100*i/ti:5.1f}%  samples {100*s/ts:5.1f}%")
    p = root / f
    if not p.exists(): continue
    lines = p.read_text().splitlines()
    starts = [(n + 1, m.group(1)) for n, l in enumerate(lines) for m in [re.match(r"^(?:template.*>\s*)?(?:static\s+)?__(?:device|global)__.*?\b(\w+)\s*\(", l)] if m]
    starts.append((len(lines) + 1, "end"))
    for (a0, name), (a1, _) in zip(starts, starts[1:]):
        i = sum(a[1] for k, a in agg.items() if k[0] == f and a0 <= k[1] < a1); s = sum(a[0] for k, a in agg.items() if k[0] == f and a0 <= k[1] < a1)
        t = sum(a[2] for k, a in agg.items() if k[0] == f and a0 <= k[1] < a1)
        if i * 200 > ti:
            print(f"    {name:28s} inst {100*i/ti:5.1f}%  samples {100*s/ts:5.1f}%  lanes {t/max(i,1):4.1f}   (lines {a0}-{a1-1})")
