"""Small end-to-end case for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import zlibts_b200 as z
from zlibts_b200 import synth

eng = z.Engine(0)
z.api.set_engine(eng)
datas = [synth.mixed(3 * 65536 + 1234, 5).tobytes(), synth.text(70000, 6).tobytes(), b"x" * 5000, b"ab", bytes(range(256)) * 9]
for d in datas:
    for ctype in (2, 1, 0):
        c = z.Deflate(d, {"compressionType": ctype}).compress().tobytes()
        assert zlib.decompress(c) == d
        assert z.Inflate(c, {"verify": True}).decompress().tobytes() == d
g = z.GZip(datas[0], {"filename": "a"}).compress()
assert z.GUnzip(g).decompress().tobytes() == datas[0]
zp = z.Zip()
for i, d in enumerate(datas):
    zp.addFile(d, "f%d" % i)
arc = zp.compress()
out = z.Unzip(arc, {"verify": True}).decompressAll()
assert [v.tobytes() for v in out.values()] == datas
print("sanitize case ok, launches", eng.launch_count)
