"""One compat deflate of 64 MiB of a single kind of data (development aid for ncu captures per kind).
usage: kind_case.py random|text|runs|records [mib]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zlibts_b200 as z
from zlibts_b200 import synth

kind = sys.argv[1]
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = mib << 20
rng = np.random.default_rng(7)
if kind == "random":
    data = rng.integers(0, 256, n, dtype=np.uint8)
elif kind == "text":
    data = synth.text(n, 1)
elif kind == "mixed":
    data = synth.mixed(n, 2)
elif kind == "runs":
    data = np.repeat(rng.integers(0, 256, n // 4096, dtype=np.uint8), 4096)
else:
    r = np.zeros((n // 8, 8), dtype=np.uint8)
    v = (7 * np.arange(n // 8, dtype=np.uint64)).astype(np.uint32)
    for k in range(4):
        r[:, k] = (v >> (8 * k)) & 0xFF
    r[:, 4] = rng.integers(0, 16, n // 8)
    data = r.reshape(-1)
data = np.ascontiguousarray(data)
s = torch.cuda.Stream()
eng = z.Engine(0, s.cuda_stream)
with torch.cuda.stream(s):
    d_in = torch.from_numpy(data).cuda()
    cap = z.deflate_bound(n)
    d_z = torch.empty(cap, dtype=torch.uint8, device="cuda")
    items = z.make_items(1); items["in_len"], items["out_cap"] = n, cap
    for it in range(2):
        r = eng.deflate_batch(d_in, d_z, items)
    print(kind, int(r["out_len"][0]) / n)
