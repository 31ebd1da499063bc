"""End-to-end batched inflate probe (development aid): N independent 64 KiB streams through zlb_inflate_batch_host on
page-locked buffers. usage: probe_inflate_host.py [streams]   (ZTS_TRACE_WAVES=1 prints when each wave is decoded)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zlibts_b200 as z
from zlibts_b200 import synth

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
CH = 65536
n = ns * CH
data = np.empty((ns, CH), dtype=np.uint8)
data[0::2] = synth.text(n // 2, 1).reshape(-1, CH)
data[1::2] = synth.mixed(n // 2, 2).reshape(-1, CH)
data = data.reshape(-1)
s = torch.cuda.Stream()
eng = z.Engine(0, s.cuda_stream)
with torch.cuda.stream(s):
    d_in = torch.from_numpy(data).cuda()
    slot = z.deflate_bound(CH)
    it = z.make_items(ns)
    it["in_off"] = np.arange(ns, dtype=np.uint64) * CH; it["in_len"] = CH
    it["out_off"] = np.arange(ns, dtype=np.uint64) * slot; it["out_cap"] = slot
    d_z = torch.empty(ns * slot, dtype=torch.uint8, device="cuda")
    r = eng.deflate_batch(d_in, d_z, it)
    hz = d_z.cpu().numpy()
lens = r["out_len"].astype(np.uint64)
offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
h_c = torch.empty(int(lens.sum()), dtype=torch.uint8).pin_memory()
hc = h_c.numpy()
for k in range(ns):
    hc[int(offs[k]):int(offs[k] + lens[k])] = hz[k * slot:k * slot + int(lens[k])]
h_o = torch.empty(n, dtype=torch.uint8).pin_memory()
it2 = z.make_items(ns)
it2["in_off"], it2["in_len"] = offs, lens
it2["out_off"] = np.arange(ns, dtype=np.uint64) * CH; it2["out_cap"] = CH
with torch.cuda.stream(s):
    d_c = h_c.cuda(); d_o = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        eng.inflate_batch(d_c, d_o, it2)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(s); eng.inflate_batch(d_c, d_o, it2); e1.record(s); e1.synchronize()
    print("device resident: %.2f ms (%.1f GB/s)" % (e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e6))
    eng.inflate_batch_host(h_c, h_o, it2)
    for _ in range(2):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(s); rr = eng.inflate_batch_host(h_c, h_o, it2); e1.record(s); e1.synchronize()
        print("end to end: %.2f ms (%.1f GB/s)  ok=%s" % (e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e6,
              bool(np.array_equal(h_o.numpy()[:1 << 20], data[:1 << 20])) and int(rr["status"].max()) == 0), flush=True)
