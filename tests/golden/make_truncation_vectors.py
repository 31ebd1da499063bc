"""Golden vectors for truncated input, produced by THE REFERENCE ITSELF (dist/Zlib-main.js under oracle/minijs):
what `new RawInflate(stream[:cut]).decompress()` does for EVERY cut of a few streams -- 'input buffer is broken'
(readBits, src/RawInflate.ts:188), 'invalid code length: N' (readCodeByTable, :238), another message, or a result.
Run here (the GPU box has no /root/reference):   python tests/golden/make_truncation_vectors.py
Writes tests/golden/truncation_vectors.json.
"""
import json
import os
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def streams():
    import oracle
    from zlibts_b200 import synth
    text = synth.text(700, 77).tobytes()
    out = [("fixed, zlib level 6", zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_FIXED)),
           ("dynamic, zlib level 9", zlib.compressobj(9, zlib.DEFLATED, -15))]
    res = []
    for name, co in out:
        s = co.compress(text) + co.flush()
        res.append((name, s))
    res.append(("dynamic, reference encoder", oracle.raw_deflate(text[:300], oracle.DYNAMIC)))
    res.append(("stored", zlib.compressobj(0, zlib.DEFLATED, -15).compress(text[:60]) + zlib.compressobj(0, zlib.DEFLATED, -15).flush()))
    co = zlib.compressobj(0, zlib.DEFLATED, -15)
    res[-1] = ("stored", co.compress(text[:60]) + co.flush())
    return res


def main():
    from oracle import refjs
    vec = []
    for name, s in streams():
        b = refjs.Batch()
        for cut in range(len(s)):
            b.add("rawinflate", s[:cut], 0, 1)
        rs = b.run()
        cuts = []
        for r in rs:
            if isinstance(r, refjs.RefError):
                cuts.append("ERR " + str(r))
            else:
                info, out = r
                cuts.append("OK %d %s" % (len(out), info["ip"]))
        vec.append({"name": name, "stream": s.hex(), "cuts": cuts})
        kinds = {}
        for c in cuts:
            k = c.split(":")[0] if c.startswith("ERR") else "OK"
            kinds[k] = kinds.get(k, 0) + 1
        print(name, len(s), kinds)
    with open(os.path.join(HERE, "truncation_vectors.json"), "w") as f:
        json.dump({"made_by": "tests/golden/make_truncation_vectors.py: dist/Zlib-main.js under oracle/minijs", "vectors": vec}, f, indent=0)


if __name__ == "__main__":
    main()
