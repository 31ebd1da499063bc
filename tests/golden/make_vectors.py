"""Generates tests/golden/*.json (run from the repo root: `python tests/golden/make_vectors.py`).

appendix_c.json    -- the provisional known-answer vectors of SURVEY.md Appendix C, typed in from the
                      survey (an independent transliteration made before this repo existed); the script
                      only re-serialises them, it does NOT derive them from the oracle.
oracle_vectors.json -- sizes and SHA-256 prefixes of the C oracle's output on a fixed input set, frozen
                      AFTER the oracle agreed with oracle/js_model.py and Appendix C; guards against drift.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

APPENDIX_C = {
    "source": "SURVEY.md Appendix C (provisional, [model]); not yet replayed under a real JS engine",
    "raw": [
        {"name": "a", "input_hex": b"a".hex(), "dynamic_hex": "05c081080000000020d6fd254e", "fixed_hex": "4b0400"},
        {"name": "abc", "input_hex": b"abc".hex(), "dynamic_hex": "05c081100000000231d63e7f87dd3a",
         "fixed_hex": "4b4c4a0600"},
        {"name": "a*10", "input_hex": (b"a" * 10).hex(), "dynamic_hex": "3dc0b1000000008030d6fc25f6b506",
         "fixed_hex": "4b840300"},
        {"name": "abc*4", "input_hex": (b"abc" * 4).hex(), "dynamic_hex": "3dc2310d00000002a0ac6aff0e7c6ea43b",
         "fixed_hex": "4b4c4a862300"},
        {"name": "hello*4", "input_hex": b"hello hello hello hello".hex(),
         "dynamic_hex": "65c43109000010c3402b6faed02150ff5b04fc720dec9e02", "fixed_hex": "cb48cdc9c957c02001"},
    ],
    "bytes_0_255": {"dynamic": [281, "a86c6b617557ad12"], "fixed": [272, "6fc4dd7e84f5a59f"]},
    "text_65536_1": {"data_sha": "f319e122f2a75c84", "crc32": "bcca4285", "adler32": "41f03696",
                     "dynamic_len": 26670, "dynamic_sha": "ac2a2aebf982e18b"},
    "mixed_65536_2": {"data_sha": "4c866b5ccc6f56b6", "crc32": "6db092d7", "adler32": "b89c85cd",
                      "dynamic_len": 38120, "dynamic_sha": "ea16379d11fe90ef"},
    "zlib_a_hex": "789c05c081080000000020d6fd254e00620062",
}


def inputs():
    """(name, bytes) pairs of the frozen-vector set; deterministic."""
    import numpy as np
    from zlibts_b200 import synth
    rng = np.random.default_rng(20261018)
    out = [("empty-ish-1", b"\x00"), ("two", b"ab"), ("three", b"abc"), ("four", b"abcd")]
    for n, a in [(64, 2), (1000, 2), (5000, 3), (20000, 4), (65536, 2), (65536, 16), (65536, 256), (65535, 64),
                 (32769, 5), (32768, 5)]:
        out.append(("rand_%d_%d" % (n, a), rng.integers(0, a, n, dtype=np.uint8).tobytes()))
    out += [("zeros_65536", b"\0" * 65536), ("xy_60000", b"xy" * 30000),
            ("text_65536_1", synth.text(65536, 1).tobytes()), ("text_30000_5", synth.text(30000, 5).tobytes()),
            ("mixed_65536_2", synth.mixed(65536, 2).tobytes()), ("mixed_65536_3_512", synth.mixed(65536, 3, 512).tobytes()),
            ("mixed_50001_6", synth.mixed(50001, 6).tobytes()), ("text_262144_7", synth.text(262144, 7).tobytes())]
    blk = rng.integers(0, 256, 1000, dtype=np.uint8).tobytes()
    out.append(("block1000_x66", (blk * 66)[:65536]))
    return out


def main():
    import oracle
    json.dump(APPENDIX_C, open(os.path.join(HERE, "appendix_c.json"), "w"), indent=1)
    cases = []
    for name, d in inputs():
        c = {"name": name, "n": len(d), "data_sha": hashlib.sha256(d).hexdigest()[:16],
             "crc32": oracle.crc32(d), "adler32": oracle.adler32(d)}
        for key, ctype in (("dynamic", oracle.DYNAMIC), ("fixed", oracle.FIXED)):
            o = oracle.raw_deflate(d, ctype)
            c[key] = [len(o), hashlib.sha256(o).hexdigest()[:16]]
        cases.append(c)
    json.dump({"source": "oracle/zts_oracle.c via tests/golden/make_vectors.py", "cases": cases},
              open(os.path.join(HERE, "oracle_vectors.json"), "w"), indent=1)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
