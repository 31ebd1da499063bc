"""Quick per-kernel timing probe (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zlibts_b200 as z
from zlibts_b200 import synth

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kind = sys.argv[2] if len(sys.argv) > 2 else "mixed"
n = mib << 20
t0 = time.time()
data = synth.mixed(n, 2) if kind == "mixed" else synth.text(n, 1)
print("gen %.2fs" % (time.time() - t0))
s = torch.cuda.Stream()
eng = z.Engine(0, s.cuda_stream)
with torch.cuda.stream(s):
    d_in = torch.from_numpy(data).cuda()
    cap = z.deflate_bound(n)
    d_z = torch.empty(cap, dtype=torch.uint8, device="cuda")
    items = z.make_items(1); items["in_len"], items["out_cap"] = n, cap
    eng.profile_enable(True)
    for it in range(3):
        eng.profile_reset()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(s)
        r = eng.deflate_batch(d_in, d_z, items)
        e1.record(s); e1.synchronize()
        ms = e0.elapsed_time(e1)
        print("deflate %d MiB: %.2f ms  %.2f GB/s  ratio %.4f" % (mib, ms, n / ms / 1e6, int(r["out_len"][0]) / n))
        for k, v in eng.profile_read().items():
            if v["launches"]: print("   %-28s %8.3f ms  x%d" % (k, v["ms"], v["launches"]))
    # inflate: independent 64 KiB streams
    nchunk = n // 65536
    caps = z.deflate_bound(65536)
    it2 = z.make_items(nchunk)
    it2["in_off"] = np.arange(nchunk) * 65536; it2["in_len"] = 65536
    it2["out_off"] = np.arange(nchunk) * caps; it2["out_cap"] = caps
    d_zz = torch.empty(nchunk * caps, dtype=torch.uint8, device="cuda")
    r = eng.deflate_batch(d_in, d_zz, it2)
    it3 = z.make_items(nchunk)
    it3["in_off"] = it2["out_off"]; it3["in_len"] = r["out_len"]
    it3["out_off"] = np.arange(nchunk) * 65536; it3["out_cap"] = 65536
    d_o = torch.empty(n, dtype=torch.uint8, device="cuda")
    for it in range(3):
        eng.profile_reset()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(s)
        r3 = eng.inflate_batch(d_zz, d_o, it3)
        e1.record(s); e1.synchronize()
        ms = e0.elapsed_time(e1)
        print("inflate %d streams: %.2f ms  %.2f GB/s out" % (nchunk, ms, n / ms / 1e6))
    assert torch.equal(d_o, d_in) and int(r3["status"].max()) == 0
    whole = z.make_items(1); whole["in_len"] = n
    for it in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(s)
        rc = eng.checksum_batch(d_in, whole)
        e1.record(s); e1.synchronize()
        print("checksum both: %.2f ms %.1f GB/s" % (e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e6))
