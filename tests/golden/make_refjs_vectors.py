"""Golden vectors produced by THE REFERENCE ITSELF: /root/reference/dist/Zlib-main.js, unmodified, executed by
oracle/minijs (a JavaScript interpreter, test infrastructure). Run here (the GPU box has no /root/reference):

    python tests/golden/make_refjs_vectors.py

Writes tests/golden/refjs_vectors.json:
  appendix_c   -- SURVEY.md Appendix C regenerated from the interpreter (not typed in): raw DYNAMIC / FIXED bytes of the
                  small inputs, sizes + SHA-256 of bytes 0..255, text(65536, 1), mixed(65536, 2), the zlib stream of "a"
  fuzz         -- 240 seeded inputs <= 4 KiB (small alphabets, runs, periodic, text, mixed): (length, sha256[:16]) of
                  the reference's DYNAMIC and FIXED output, the inputs being reproducible from (seed, recipe)
  chunks       -- 64 KiB chunks of the benchmark generators: the same for the reference's DYNAMIC output
  containers   -- zlib / gzip / zip bytes (time fields masked / fixed date) of a few inputs
  inflate      -- the reference's RawInflate on zlib-made streams: output sha + .ip
  lengths      -- getLengths (Heap + reversePackageMerge) on tie-heavy histograms
"""
import hashlib
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()[:16]


def fuzz_inputs():
    """(recipe, bytes): deterministic, reproducible from the recipe alone (tests rebuild them with this function)."""
    from zlibts_b200 import synth
    out = []
    rng = np.random.default_rng(20261019)
    for i in range(240):
        kind = i % 8
        n = int(rng.integers(1, 4097))
        if kind == 0:
            d = rng.integers(0, 2, n, dtype=np.uint8)
        elif kind == 1:
            d = rng.integers(0, 3, n, dtype=np.uint8) + 97
        elif kind == 2:
            d = rng.integers(0, 16, n, dtype=np.uint8)
        elif kind == 3:
            d = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == 4:
            per = int(rng.integers(1, 40))
            d = np.resize(rng.integers(0, 256, per, dtype=np.uint8), n)
        elif kind == 5:
            d = np.repeat(rng.integers(0, 4, (n + 15) // 16, dtype=np.uint8), 16)[:n]
        elif kind == 6:
            d = synth.text(n, 3000 + i)
        else:
            d = synth.mixed(n, 3000 + i, 64)
        out.append(({"i": i, "kind": kind, "n": n}, np.ascontiguousarray(d, dtype=np.uint8).tobytes()))
    return out


def chunk_inputs():
    from zlibts_b200 import synth
    rng = np.random.default_rng(20261020)
    blk = rng.integers(0, 256, 1000, dtype=np.uint8).tobytes()
    return [("text_65536_1", synth.text(65536, 1).tobytes()), ("mixed_65536_2", synth.mixed(65536, 2).tobytes()),
            ("mixed_65536_3_512", synth.mixed(65536, 3, 512).tobytes()), ("text_30000_5", synth.text(30000, 5).tobytes()),
            ("rand2_65536", rng.integers(0, 2, 65536, dtype=np.uint8).tobytes()),
            ("rand256_65536", rng.integers(0, 256, 65536, dtype=np.uint8).tobytes()),
            ("zeros_65536", b"\0" * 65536), ("xy_60000", b"xy" * 30000), ("block1000_x66", (blk * 66)[:65536]),
            ("text_1048576_1", synth.text(1 << 20, 1).tobytes())]   # C1: one block for the whole MiB, Uint16 heap wrap and all


def length_cases():
    rng = np.random.default_rng(20261021)
    cases = []
    for k in range(60):
        nsym = [286, 30, 19][k % 3]
        limit = [15, 7, 7][k % 3]
        mode = k % 4
        if mode == 0:
            f = rng.integers(0, 3, nsym)
        elif mode == 1:
            f = rng.integers(0, 1000, nsym) * (rng.random(nsym) < 0.3)
        elif mode == 2:
            f = np.ones(nsym, dtype=np.int64) * int(rng.integers(1, 5))
        else:
            f = np.floor(2.0 ** (np.arange(nsym) % 24) * rng.random()).astype(np.int64) % 65536
        if nsym == 19:
            f = f % 256
        cases.append((np.asarray(f, dtype=np.uint32), limit))
    return cases


def main():
    from oracle import refjs
    assert refjs.available(), "needs /root/reference and g++"
    small = [("a", b"a"), ("abc", b"abc"), ("a*10", b"a" * 10), ("abc*4", b"abc" * 4), ("hello*4", b"hello hello hello hello")]
    b = refjs.Batch()
    idx = {}
    for name, d in small:
        idx["s", name, 2] = b.add("rawdeflate", d, 2, 0)
        idx["s", name, 1] = b.add("rawdeflate", d, 1, 0)
    b255 = bytes(range(256))
    idx["b255", 2] = b.add("rawdeflate", b255, 2, 0)
    idx["b255", 1] = b.add("rawdeflate", b255, 1, 0)
    idx["zlib_a"] = b.add("deflate", b"a", 2)
    fz = fuzz_inputs()
    for rec, d in fz:
        idx["f", rec["i"], 2] = b.add("rawdeflate", d, 2, 0)
        idx["f", rec["i"], 1] = b.add("rawdeflate", d, 1, 0)
    ch = chunk_inputs()
    for name, d in ch:
        idx["c", name] = b.add("rawdeflate", d, 2, 0)
        idx["crc", name] = b.add_value("crc32", d)
        idx["adl", name] = b.add_value("adler32", d)
    # containers
    cont_in = [("hello*4", b"hello hello hello hello"), ("text_5000_9", None)]
    from zlibts_b200 import synth
    cont_in[1] = ("text_5000_9", synth.text(5000, 9).tobytes())
    for name, d in cont_in:
        idx["z", name] = b.add("deflate", d, 2)
        idx["g", name] = b.add("gzip", d)
    idx["zip"] = b.add_zip([("a.txt", cont_in[0][1]), ("dir/b.bin", cont_in[1][1]), ("empty", b"")], 1700000000000)
    # inflate direction: streams made by CPython zlib (all block types), decoded by the reference
    inf_cases = []
    rng = np.random.default_rng(20261022)
    for k in range(40):
        n = int(rng.integers(1, 20000))
        d = synth.mixed(n, 7000 + k, 256).tobytes() if k % 2 else synth.text(n, 7000 + k).tobytes()
        level = [0, 1, 6, 9][k % 4]
        strat = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE][(k // 4) % 4]
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strat)
        s = co.compress(d) + co.flush()
        inf_cases.append({"k": k, "n": n, "level": level, "strategy": int(strat), "stream_sha": sha(s), "data_sha": sha(d)})
        idx["i", k] = b.add("rawinflate", s + b"\0\0\0\0", 0, 1)   # trailing bytes: SURVEY B-7
    lc = length_cases()
    for k, (f, limit) in enumerate(lc):
        idx["l", k] = b.add("lengths", f.astype("<u4").tobytes(), limit)
    res = b.run()

    def get(key):
        r = res[idx[key]]
        if isinstance(r, refjs.RefError):
            raise r
        return r

    out = {"source": "/root/reference/dist/Zlib-main.js executed by oracle/minijs (tests/golden/make_refjs_vectors.py)",
           "bundle_sha256": hashlib.sha256(open(refjs.BUNDLE, "rb").read()).hexdigest()}
    ac = {"raw": []}
    for name, d in small:
        ac["raw"].append({"name": name, "input_hex": d.hex(), "dynamic_hex": get(("s", name, 2))[1].hex(),
                          "fixed_hex": get(("s", name, 1))[1].hex()})
    ac["bytes_0_255"] = {"dynamic": [len(get(("b255", 2))[1]), sha(get(("b255", 2))[1])],
                         "fixed": [len(get(("b255", 1))[1]), sha(get(("b255", 1))[1])]}
    for name in ("text_65536_1", "mixed_65536_2"):
        d = dict(ch)[name]
        o = get(("c", name))[1]
        ac[name] = {"data_sha": sha(d), "crc32": "%08x" % int(get(("crc", name))[0]["value"]),
                    "adler32": "%08x" % int(get(("adl", name))[0]["value"]), "dynamic_len": len(o), "dynamic_sha": sha(o)}
    ac["zlib_a_hex"] = get("zlib_a")[1].hex()
    out["appendix_c"] = ac
    out["fuzz"] = [{"i": rec["i"], "kind": rec["kind"], "n": rec["n"], "data_sha": sha(d),
                    "dynamic": [len(get(("f", rec["i"], 2))[1]), sha(get(("f", rec["i"], 2))[1])],
                    "fixed": [len(get(("f", rec["i"], 1))[1]), sha(get(("f", rec["i"], 1))[1])]} for rec, d in fz]
    out["chunks"] = [{"name": name, "n": len(d), "data_sha": sha(d), "crc32": int(get(("crc", name))[0]["value"]),
                      "adler32": int(get(("adl", name))[0]["value"]),
                      "dynamic": [len(get(("c", name))[1]), sha(get(("c", name))[1])]} for name, d in ch]
    cont = {}
    for name, d in cont_in:
        z = get(("z", name))[1]
        g = bytearray(get(("g", name))[1])
        g[4:8] = b"\0\0\0\0"  # MTIME = Date.now()
        cont[name] = {"zlib": [len(z), sha(z)], "gzip_mtime_masked": [len(g), sha(g)]}
    cont["zip_3_files_date_1700000000000"] = {"len": len(get("zip")[1]), "sha": sha(get("zip")[1]),
                                              "hex_head": get("zip")[1][:64].hex()}
    out["containers"] = cont
    for c in inf_cases:
        info, o = get(("i", c["k"]))
        assert sha(o) == c["data_sha"], c
        c["ip"] = int(info["ip"])
    out["inflate"] = inf_cases
    out["lengths"] = [{"k": k, "limit": limit, "freqs": f.tolist(), "lengths": list(get(("l", k))[1])}
                      for k, (f, limit) in enumerate(lc)]
    json.dump(out, open(os.path.join(HERE, "refjs_vectors.json"), "w"), indent=0, separators=(",", ":"))
    print("wrote refjs_vectors.json:", len(out["fuzz"]), "fuzz,", len(out["chunks"]), "chunks,", len(inf_cases),
          "inflate,", len(lc), "length cases")


if __name__ == "__main__":
    main()
