"""Batched inflate probe (development aid): N independent 64 KiB streams, even = text, odd = mixed (the C3 shape),
made by the GPU encoder (compat bytes = the reference's per chunk). usage: probe_inflate.py [streams]
ZLB_INFLATE_GROUP=32|16|8 selects the lanes per stream."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zlibts_b200 as z
from zlibts_b200 import synth

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n = ns * 65536
half = n // 2
t = synth.text(half, 1).reshape(-1, 65536)
m = synth.mixed(half, 2).reshape(-1, 65536)
data = np.empty((ns, 65536), dtype=np.uint8)
data[0::2] = t
data[1::2] = m
data = data.reshape(-1)
s = torch.cuda.Stream()
eng = z.Engine(0, s.cuda_stream)
with torch.cuda.stream(s):
    d_in = torch.from_numpy(data).cuda()
    caps = z.deflate_bound(65536)
    it2 = z.make_items(ns)
    it2["in_off"] = np.arange(ns, dtype=np.uint64) * 65536; it2["in_len"] = 65536
    it2["out_off"] = np.arange(ns, dtype=np.uint64) * caps; it2["out_cap"] = caps
    d_zz = torch.empty(ns * caps, dtype=torch.uint8, device="cuda")
    r = eng.deflate_batch(d_in, d_zz, it2)
    it3 = z.make_items(ns)
    it3["in_off"] = it2["out_off"]; it3["in_len"] = r["out_len"]
    it3["out_off"] = np.arange(ns, dtype=np.uint64) * 65536; it3["out_cap"] = 65536
    d_o = torch.zeros(n, dtype=torch.uint8, device="cuda")
    eng.profile_enable(True)
    for it in range(4):
        eng.profile_reset()
        r3 = eng.inflate_batch(d_zz, d_o, it3)
        ms = sum(v["ms"] for k, v in eng.profile_read().items() if "inflate" in k)
    ok = bool(torch.equal(d_o, d_in)) and int(r3["status"].max()) == 0
    print("group %s: inflate %d streams: %.3f ms  %.2f GB/s out  ok=%s" % (os.environ.get("ZLB_INFLATE_GROUP", "default"), ns, ms, n / ms / 1e6, ok))
