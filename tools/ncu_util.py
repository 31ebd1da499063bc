#!/usr/bin/env python
"""Lane utilisation per barrier-delimited phase of a kernel in an ncu capture (development aid):
thread instructions / (32 x warp instructions). usage: ncu_util.py <rep> <kernel regex>"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ie, te = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
si = hdr.index("# Samples")
body = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
toti = sum(float(r[ie]) for r in body)
bars = [i for i, r in enumerate(body) if "BAR." in r[1]] + [len(body) - 1]
prev = 0
for b in bars:
    seg = body[prev:b + 1]
    ins = sum(float(r[ie]) for r in seg); th = sum(float(r[te]) for r in seg)
    if ins / toti > 0.004:
        print("sass %4d..%4d  instructions %5.1f%%  lanes/instr %5.1f" % (prev, b, 100 * ins / toti, th / max(ins, 1)))
    prev = b + 1
if len(sys.argv) > 3:   # dump a sass range: lo hi
    lo, hi2 = int(sys.argv[3]), int(sys.argv[4])
    for i in range(lo, hi2 + 1):
        r = body[i]
        print("%5d %10.0f %5.1f %6.0f  %s" % (i, float(r[ie]), float(r[te]) / max(float(r[ie]), 1), float(r[si]), r[1][:90]))
