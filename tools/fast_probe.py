"""fast-mode sweep: speed and ratio vs candidate depth (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, zlib
import zlibts_b200 as z
from zlibts_b200 import synth
s = torch.cuda.Stream(); eng = z.Engine(0, s.cuda_stream)
for kind, n in (("mixed", 256 << 20), ("text", 64 << 20)):
    data = synth.mixed(n, 2) if kind == "mixed" else synth.text(n, 1)
    with torch.cuda.stream(s):
        d_in = torch.from_numpy(data).cuda(); cap = z.deflate_bound(n)
        d_z = torch.empty(cap, dtype=torch.uint8, device="cuda"); d_o = torch.empty(n, dtype=torch.uint8, device="cuda")
        it = z.make_items(1); it["in_len"], it["out_cap"] = n, cap
        base = None
        for mode, name in [(z.MODE_COMPAT, "compat")] + [(z.mode_fast(d), "fast%d" % d) for d in (4, 8, 16, 32, 64)]:
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record(s); r = eng.deflate_batch(d_in, d_z, it, mode=mode); e1.record(s); e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            clen = int(r["out_len"][0]); base = base or clen
            it2 = z.make_items(1); it2["in_len"], it2["out_cap"] = clen, n
            r2 = eng.inflate_batch(d_z, d_o, it2, z.INFLATE_SPLIT)
            ok = int(r2["status"][0]) == 0 and torch.equal(d_o, d_in)
            print("%-6s %-8s %7.2f ms %6.2f GB/s ratio %.4f (%+.2f%% vs compat) roundtrip %s" % (kind, name, best, n / best / 1e6, clen / n, 100.0 * (clen - base) / base, ok))
