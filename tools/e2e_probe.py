"""End-to-end (pinned host buffers) timing of the C2 step, development aid. usage: e2e_probe.py [mib]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zlibts_b200 as z
from zlibts_b200 import synth
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = mib << 20
data = synth.mixed(n, 2)
s = torch.cuda.Stream()
eng = z.Engine(0, s.cuda_stream)
cap = z.deflate_bound(n)
h_in = torch.from_numpy(data).pin_memory(); h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
items = z.make_items(1); items["in_len"], items["out_cap"] = n, cap
with torch.cuda.stream(s):
    for mode, name in ((z.MODE_COMPAT, "compat"), (z.MODE_FAST, "fast")):
        for _ in range(2):
            eng.deflate_batch_host(h_in, h_out, items, mode=mode)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(s)
        for _ in range(5):
            r = eng.deflate_batch_host(h_in, h_out, items, mode=mode)
        e1.record(s); e1.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("%s e2e %.2f ms  %.2f GB/s  waves=%s" % (name, ms, n / ms / 1e6, os.environ.get("ZTS_HOST_WAVES", "8")), flush=True)
