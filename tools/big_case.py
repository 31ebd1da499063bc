import sys, time, zlib
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import zlibts_b200 as z
from zlibts_b200 import synth
n = (5 << 30) + 12345
t = time.time()
part = synth.mixed(1 << 28, 99)
h = np.empty(n, dtype=np.uint8)
for k in range(0, n, 1 << 28):
    m = min(1 << 28, n - k)
    h[k:k + m] = part[:m]
    h[k:k + 64] = np.frombuffer(k.to_bytes(8, 'little') * 8, dtype=np.uint8)  # make the copies differ
print('gen', time.time() - t)
s = torch.cuda.Stream(); e = z.Engine(0, s.cuda_stream)
cap = z.deflate_bound(n)
it = z.make_items(1); it['in_len'], it['out_cap'] = n, cap
with torch.cuda.stream(s):
    d_in = torch.from_numpy(h).cuda(); d_z = torch.empty(cap, dtype=torch.uint8, device='cuda')
    t = time.time(); r = e.deflate_batch(d_in, d_z, it, flags=z.DEFLATE_WANT_CRC32 | z.DEFLATE_WANT_ADLER32); print('deflate s', time.time() - t, r)
    clen = int(r['out_len'][0]); assert int(r['status'][0]) == 0
    d_o = torch.zeros(n, dtype=torch.uint8, device='cuda')
    it2 = z.make_items(1); it2['in_len'], it2['out_cap'] = clen, n
    t = time.time(); r2 = e.inflate_batch(d_z, d_o, it2, z.INFLATE_SPLIT | z.INFLATE_WANT_CRC32); print('inflate s', time.time() - t, r2)
    assert int(r2['status'][0]) == 0 and int(r2['out_len'][0]) == n and int(r2['in_used'][0]) == clen
    assert torch.equal(d_o, d_in)
crc = zlib.crc32(h); ad = zlib.adler32(h)
assert int(r['crc32'][0]) == crc == int(r2['crc32'][0]) and int(r['adler32'][0]) == ad, (r, crc, ad)
print('5 GiB ok', clen / n)
