#!/usr/bin/env python
"""Splits an ncu source-page capture of a kernel at its BAR.SYNC instructions: share of stall samples and of
executed instructions per barrier-delimited phase (development aid). usage: ncu_phases.py <rep> <kernel regex>"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
si, ie, sb = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("stall_barrier")
body = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
tot = sum(float(r[si]) for r in body)
toti = sum(float(r[ie]) for r in body)
print("total samples %d, warp instructions %d" % (tot, toti))
bars = [i for i, r in enumerate(body) if "BAR." in r[1]] + [len(body) - 1]
prev = 0
for b in bars:
    seg = body[prev:b + 1]
    s = sum(float(r[si]) for r in seg); bs = sum(float(r[sb]) for r in seg); ins = sum(float(r[ie]) for r in seg)
    if s / tot > 0.004 or ins / toti > 0.004:
        print("sass %4d..%4d  samples %5.1f%% (barrier wait %5.1f%%)  instructions %5.1f%%" % (prev, b, 100 * s / tot, 100 * bs / tot, 100 * ins / toti))
    prev = b + 1
