#!/usr/bin/env python
"""Per-source-line view of an ncu capture (development aid).

ncu's CSV source page is per SASS instruction; this joins it with `nvdisasm -g` line info of the
matching cubin (extracted from the built .so) and prints the lines that hold the stall samples.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep lz77_chunk zts_lz77 [top_n]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "zlib.ts_b200", "libzlibts_b200.so")


def sass_lines(cubin_stem, kernel_pat):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.startswith(cubin_stem + ".") and f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], check=True, capture_output=True,
                         text=True).stdout
    out, cur, infunc = [], None, False
    for line in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            infunc = re.search(kernel_pat, m.group(1)) is not None
            continue
        if not infunc:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def main():
    rep, kpat, stem = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kpat],
                            capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(csvtxt)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    # only the first captured launch of the kernel
    body = []
    for r in rows[hi + 1:]:
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        body.append(r)
    sass = sass_lines(stem, os.environ.get("NCU_CUBIN_PAT", kpat))  # mangled-name pattern when kpat is ambiguous there
    if len(sass) != len(body):
        print("warning: %d SASS instructions in the cubin vs %d in the report (rebuilt since the capture?)" %
              (len(sass), len(body)), file=sys.stderr)
    col = {n: hdr.index(n) for n in hdr}
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    per = defaultdict(lambda: defaultdict(float))
    for (addr, loc, text), r in zip(sass, body):
        d = per[loc]
        d["samples"] += float(r[col["# Samples"]] or 0)
        d["inst"] += float(r[col["Instructions Executed"]] or 0)
        d["conf"] += float(r[col["L1 Wavefronts Shared Excessive"]] or 0)
        for s in stall_cols:
            d[s] += float(r[col[s]] or 0)
    tot = sum(d["samples"] for d in per.values()) or 1
    toti = sum(d["inst"] for d in per.values()) or 1
    srcs = {}
    print("total samples %d, warp instructions %d" % (tot, toti))
    for loc, d in sorted(per.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        if loc is None:
            continue
        f = os.path.join(ROOT, "zlib.ts_b200", "csrc", loc[0])
        if f not in srcs:
            srcs[f] = open(f).read().splitlines() if os.path.exists(f) else []
        text = srcs[f][loc[1] - 1].strip() if 0 < loc[1] <= len(srcs[f]) else ""
        stalls = sorted(((d[s], s[6:]) for s in stall_cols), reverse=True)[:3]
        st = " ".join("%s=%.0f%%" % (n, 100 * v / max(d["samples"], 1)) for v, n in stalls if v)
        print("%5.1f%% smp %5.1f%% inst %7.0f xs-wave  %s:%d  %-70s [%s]" %
              (100 * d["samples"] / tot, 100 * d["inst"] / toti, d["conf"], loc[0], loc[1], text[:70], st))


if __name__ == "__main__":
    main()
