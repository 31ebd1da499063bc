#!/usr/bin/env python
"""Summarises ncu captures into profiles/: key metrics per kernel (JSON), per-line hot spots (text), traffic.json.

    python tools/profile_summary.py <tag>      # reads gpurun_out/prof_deflate_<tag>.ncu-rep, prof_inflate_<tag>.ncu-rep
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static"]
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def rows_of(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2:]


def main():
    tag = sys.argv[1]
    units_per_launch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096  # chunks (streams) one captured launch processed
    out = {}
    traffic = {}
    inst = {}
    for leg in ("deflate", "fast", "inflate", "frame"):
        rep = os.path.join(ROOT, "gpurun_out", "prof_%s_%s.ncu-rep" % (leg, tag))
        if not os.path.exists(rep):
            continue
        hdr, units, body = rows_of(rep)
        for r in body:
            name = r[hdr.index("Kernel Name")].split("(")[0]
            d = {}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    try:
                        v = float(r[i].replace(",", ""))
                    except ValueError:
                        continue
                    u = units[i]
                    d[k] = {"value": v, "unit": u}
            out.setdefault(name, []).append(d)
            if leg == "frame":   # captured on another workload (C4 sample): not part of the bench step's traffic table
                continue
            if "smsp__inst_executed.sum" in d and name not in inst:
                inst[name] = d["smsp__inst_executed.sum"]["value"]
            try:
                rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
                traffic[name] = rd["value"] * UNIT_SCALE.get(rd["unit"], 1) + wr["value"] * UNIT_SCALE.get(wr["unit"], 1)
            except KeyError:
                pass
    json.dump(out, open(os.path.join(ROOT, "profiles", "%s_ncu_summary.json" % tag), "w"), indent=1)
    json.dump({"_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes), ncu --set full, tag " + tag,
               "_units_per_launch": units_per_launch, "_inst_executed": inst, **traffic}, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", "%s_source_hotspots.txt" % tag), "w") as f:
        for leg, pat, stem in (("deflate", "lz77_chunk", "zts_lz77"), ("deflate", "huffman_build", "zts_huffman"),
                               ("deflate", "bitpack", "zts_deflate"),
                               ("inflate", "inflate_warp", "zts_inflate")):
            rep = os.path.join(ROOT, "gpurun_out", "prof_%s_%s.ncu-rep" % (leg, tag))
            if os.path.exists(rep):
                f.write("==== %s ====\n" % pat)
                f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, pat, stem, "30"],
                                       capture_output=True, text=True).stdout)
    for k, v in traffic.items():
        print(k, "%.1f MB/launch" % (v / 1e6))


if __name__ == "__main__":
    main()
