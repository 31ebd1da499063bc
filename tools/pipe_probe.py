"""Device-resident C2 step with and without the per-kernel timers (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zlibts_b200 as z
from zlibts_b200 import synth
n = 256 << 20
data = synth.mixed(n, 2)
s = torch.cuda.Stream()
eng = z.Engine(0, s.cuda_stream)
cap = z.deflate_bound(n)
items = z.make_items(1); items["in_len"], items["out_cap"] = n, cap
with torch.cuda.stream(s):
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    for prof in (False, True, False, True):
        eng.profile_enable(prof)
        for _ in range(3):
            eng.deflate_batch(d_in, d_out, items)
        eng.profile_reset()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(s)
        for _ in range(5):
            r = eng.deflate_batch(d_in, d_out, items)
        e1.record(s); e1.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("timers %s: %.2f ms per step  %.2f GB/s" % (prof, ms, n / ms / 1e6), flush=True)
        if prof:
            for k, v in eng.profile_read().items():
                if v["launches"]: print("   %-28s %8.3f ms  x%d" % (k, v["ms"] / 5, v["launches"]))
