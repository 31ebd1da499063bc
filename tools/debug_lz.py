"""Development aid: find chunks whose compat bytes differ from the oracle's and show the first differing token."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import zlibts_b200 as z
from zlibts_b200 import synth
import oracle
from test_deflate_gpu import oracle_tokens

kind = sys.argv[1] if len(sys.argv) > 1 else "runs"
mib = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n = mib << 20
rng = np.random.default_rng(7)
if kind == "runs":
    data = np.repeat(rng.integers(0, 256, n // 4096, dtype=np.uint8), 4096)
elif kind == "mixed":
    data = synth.mixed(n, 2)
else:
    data = synth.text(n, 1)
data = np.ascontiguousarray(data)
eng = z.Engine(0)
nch = n // 65536
items = z.make_items(nch)
cap = z.deflate_bound(65536)
items["in_off"] = np.arange(nch, dtype=np.uint64) * 65536
items["in_len"] = 65536
items["out_off"] = np.arange(nch, dtype=np.uint64) * cap
items["out_cap"] = cap
d_in = torch.from_numpy(data).cuda()
d_out = torch.zeros(nch * cap, dtype=torch.uint8, device="cuda")
res = eng.deflate_batch(d_in, d_out, items)
h = d_out.cpu().numpy()
bad = []
for c in range(nch):
    want = oracle.raw_deflate(data[c * 65536:(c + 1) * 65536])
    got = h[c * cap:c * cap + int(res["out_len"][c])].tobytes()
    if got != want:
        bad.append(c)
print("chunks", nch, "bad", len(bad), bad[:10])
for c in bad[:3]:
    chunk = data[c * 65536:(c + 1) * 65536]
    tok, hist = eng.debug_lz77(torch.from_numpy(chunk.copy()).cuda(), 65536)
    wt, wh = oracle_tokens(chunk.tobytes())
    print("chunk", c, "single-chunk tokens", len(tok), "oracle", len(wt))
    m = min(len(tok), len(wt))
    d = np.nonzero(tok[:m] != wt[:m])[0]
    if d.size:
        i = int(d[0])
        pos = 0
        for k in range(i):
            t = int(wt[k]); pos += ((t >> 16) & 0xFF) + 3 if t & 0x80000000 else 1
        print("  first diff at token", i, "position", pos, "tile", pos // 256, "got", [hex(int(x)) for x in tok[i:i+4]], "want", [hex(int(x)) for x in wt[i:i+4]])
    else:
        print("  single-chunk tokens equal up to", m)
