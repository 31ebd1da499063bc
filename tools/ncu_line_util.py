#!/usr/bin/env python
"""Per source line of a kernel in an ncu capture: share of the warp instructions and active lanes per instruction
(development aid). usage: ncu_line_util.py <rep> <kernel regex> <cubin stem> [top_n]"""
import csv, io, os, subprocess, sys
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import sass_lines, ROOT

rep, kpat, stem = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kpat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ie, te = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
body = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
sass = sass_lines(stem, os.environ.get("NCU_CUBIN_PAT", kpat))
per = defaultdict(lambda: [0.0, 0.0])
for (addr, loc, text), r in zip(sass, body):
    per[loc][0] += float(r[ie] or 0)
    per[loc][1] += float(r[te] or 0)
tot = sum(v[0] for v in per.values()) or 1
srcs = {}
for loc, (ins, th) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    if loc is None:
        continue
    f = os.path.join(ROOT, "zlib.ts_b200", "csrc", loc[0])
    if f not in srcs:
        srcs[f] = open(f).read().splitlines() if os.path.exists(f) else []
    text = srcs[f][loc[1] - 1].strip() if 0 < loc[1] <= len(srcs[f]) else ""
    print("%5.1f%% inst  %4.1f lanes  %s:%d  %s" % (100 * ins / tot, th / max(ins, 1), loc[0], loc[1], text[:90]))
