"""Runs the C2 deflate repeatedly (compat, primed, fast) and checks that every run writes the same bytes (development
aid: the LZ77 kernel hands tiles from warp to warp through flags; a race would show up as a run that differs)."""
import sys, os, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zlibts_b200 as z
from zlibts_b200 import synth
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = 256 << 20
data = synth.mixed(n, 2)
s = torch.cuda.Stream()
eng = z.Engine(0, s.cuda_stream)
with torch.cuda.stream(s):
    d_in = torch.from_numpy(data).cuda()
    for name, mode in (("compat", z.MODE_COMPAT), ("primed", z.MODE_PRIMED), ("fast", z.MODE_FAST)):
        cap = z.deflate_bound(n, 0, z.DYNAMIC, mode)
        d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        items = z.make_items(1); items["in_len"], items["out_cap"] = n, cap
        ref = None
        for k in range(reps):
            d_out.zero_()
            r = eng.deflate_batch(d_in, d_out, items, mode=mode)
            m = int(r["out_len"][0])
            sig = (m, int(d_out[:m].to(torch.int64).sum().item()), int((d_out[:m].to(torch.int64) * torch.arange(m, device="cuda") % 1000003).sum().item()))
            if ref is None:
                ref = sig
                first = d_out[:m].clone()
            assert sig == ref and torch.equal(d_out[:m], first), (name, k, sig, ref)
        print(name, "x%d identical" % reps, ref[0], flush=True)
        if name == "compat":
            assert zlib.decompress(first.cpu().numpy().tobytes(), -15) == data.tobytes()
            print("   decodes to the input", flush=True)
