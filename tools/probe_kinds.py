"""Where the LZ77 time goes per kind of data (development aid): the four segment kinds of mixed(), each alone."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zlibts_b200 as z
from zlibts_b200 import synth

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = mib << 20
rng = np.random.default_rng(7)


def records(n):
    r = np.zeros((n // 8, 8), dtype=np.uint8)
    v = (7 * np.arange(n // 8, dtype=np.uint64)).astype(np.uint32)
    for k in range(4):
        r[:, k] = (v >> (8 * k)) & 0xFF
    r[:, 4] = rng.integers(0, 16, n // 8)
    return r.reshape(-1)


kinds = {
    "random": lambda: rng.integers(0, 256, n, dtype=np.uint8),
    "text": lambda: synth.text(n, 1),
    "runs": lambda: np.repeat(rng.integers(0, 256, n // 4096, dtype=np.uint8), 4096),
    "records": lambda: records(n),
    "mixed": lambda: synth.mixed(n, 2),
}
s = torch.cuda.Stream()
eng = z.Engine(0, s.cuda_stream)
with torch.cuda.stream(s):
    for name, gen in kinds.items():
        data = np.ascontiguousarray(gen())
        d_in = torch.from_numpy(data).cuda()
        cap = z.deflate_bound(n)
        d_z = torch.empty(cap, dtype=torch.uint8, device="cuda")
        items = z.make_items(1); items["in_len"], items["out_cap"] = n, cap
        eng.profile_enable(True)
        for mode, label in ((z.MODE_COMPAT, "compat"), (z.MODE_FAST, "fast")):
            for it in range(3):
                eng.profile_reset()
                r = eng.deflate_batch(d_in, d_z, items, mode=mode)
                prof = eng.profile_read()
            lz = sum(v["ms"] for k, v in prof.items() if "lz77" in k)
            tot = sum(v["ms"] for k, v in prof.items())
            print("%-8s %-6s %4d MiB: lz77 %7.3f ms (%.2f GB/s)  all kernels %7.3f ms  ratio %.4f" % (
                name, label, mib, lz, n / lz / 1e6, tot, int(r["out_len"][0]) / n), flush=True)
