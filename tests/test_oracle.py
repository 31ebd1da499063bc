"""CPU: pins the C oracle (oracle/zts_oracle.c) -- the checker every GPU parity test leans on.

The executed reference pins the oracle in tests/test_refjs.py (dist/Zlib-main.js under oracle/minijs). This file
holds the C restatement against three further, independent witnesses:
  1. oracle/js_model.py, a separately written statement-by-statement Python model of the TS sources;
  2. the provisional known-answer vectors of SURVEY.md Appendix C (tests/golden/appendix_c.json);
  3. CPython zlib as the RFC 1950/1951 cross-oracle (decodes every stream, produces streams to decode,
     crc32 / adler32).
tests/golden/oracle_vectors.json (made by tests/golden/make_vectors.py) freezes oracle outputs so that a
later edit of the oracle cannot drift silently.
"""
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

import oracle
from oracle import js_model as js
from helpers import rand_bytes, zlib_raw

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def small_inputs():
    rng = np.random.default_rng(101)
    ins = [b"a", b"ab", b"abc", b"abcd", b"aaaaaaaaaa", b"abcabcabcabc", b"hello hello hello hello", bytes(range(256)),
           b"\0" * 700, b"ab" * 400, b"bbbbb"]
    for n, a in [(1, 2), (2, 2), (3, 2), (4, 2), (7, 2), (50, 2), (300, 2), (1500, 2), (1500, 3), (1200, 4), (900, 16),
                 (2000, 26), (1000, 256), (259, 1), (260, 1), (517, 1), (3000, 5)]:
        ins.append(rand_bytes(rng, n, a).tobytes())
    blk = rand_bytes(rng, 300, 256).tobytes()
    ins.append(blk * 5)                                  # equal-length candidates: nearest must win
    ins.append(blk[:100] + b"x" + blk[:100] + b"y" + blk[:99])
    return ins


def test_checksums_match_model_and_zlib():
    rng = np.random.default_rng(1)
    for n in [0, 1, 7, 8, 9, 1023, 1024, 1025, 5552, 5553, 70000]:
        d = rand_bytes(rng, n).tobytes()
        assert oracle.crc32(d) == zlib.crc32(d) == js.crc32(d)
        assert oracle.adler32(d) == zlib.adler32(d) == js.adler32(d)
    d = b"\xff" * 200000  # worst case for the deferred modulo (src/Adler32.ts:38-45)
    assert oracle.adler32(d) == zlib.adler32(d)
    # update() chaining (src/CRC32.ts:25, src/Adler32.ts:28)
    a, b = rand_bytes(rng, 3000).tobytes(), rand_bytes(rng, 5000).tobytes()
    assert oracle.crc32(b, oracle.crc32(a)) == zlib.crc32(a + b)
    assert oracle.adler32(b, oracle.adler32(a)) == zlib.adler32(a + b)


def test_lz77_tokens_match_model():
    for d in small_inputs():
        m = js.LZ77(d, 0)
        want = m.encode()
        tok, fl, fd = oracle.lz77(d)
        assert tok.tolist() == want, len(d)
        assert fl.tolist() == m.freqs_litlen and fd.tolist() == m.freqs_dist
        assert fl[256] == 2                               # src/LZ77.ts:127 + :279


def test_lz77_closed_form_spec():
    """longest match over ALL earlier equal-key positions within 32768, nearest on ties, greedy."""
    rng = np.random.default_rng(5)
    for n, a in [(400, 2), (600, 3), (800, 4), (500, 8)]:
        d = rand_bytes(rng, n, a).tobytes()
        toks, p = [], 0
        while p < n:
            if p + 3 >= n:
                toks += list(d[p:])
                break
            best_len, best_q = 0, -1
            for q in range(p - 1, max(-1, p - 32769), -1):
                if d[q:q + 3] != d[p:p + 3]:
                    continue
                length = 0
                while length < min(258, n - p) and d[q + length] == d[p + length]:
                    length += 1
                if length > best_len:
                    best_len, best_q = length, q
            if best_len >= 3:
                toks.append((best_len, p - best_q))
                p += best_len
            else:
                toks.append(d[p])
                p += 1
        tok, _, _ = oracle.lz77(d)
        t, got, i = tok.tolist(), [], 0
        while t[i] != 256:
            if t[i] < 256:
                got.append(t[i])
                i += 1
            else:
                got.append((js._LEN_BASE[t[i] - 257] + t[i + 1], js._DIST_BASE[t[i + 3]] + t[i + 4]))
                i += 6
        assert got == toks, (n, a)


def test_code_lengths_match_model_heap_and_package_merge():
    rng = np.random.default_rng(2)
    diag = {}
    for trial in range(400):
        nsym, limit = [(286, 15), (30, 7), (19, 7)][trial % 3]
        kind = trial % 7
        if kind == 0:
            f = rng.integers(0, 4, nsym)
        elif kind == 1:
            f = rng.integers(0, 60000, nsym)
        elif kind == 2:
            f = rng.geometric(0.02, nsym) * (rng.random(nsym) < 0.5)
        elif kind == 3:
            f = np.zeros(nsym, dtype=np.int64)
            f[rng.integers(0, nsym, int(rng.integers(1, 4)))] = rng.integers(1, 100)
        elif kind == 4:
            f = (2.0 ** (np.arange(nsym) % 24) * rng.random()).astype(np.int64) % 65536   # limit binds
        elif kind == 5:
            f = np.full(nsym, int(rng.integers(1, 9)))                                    # all ties: heap order decides
        else:
            f = rng.integers(65530, 65545, nsym)                                          # Uint16 wrap (src/Heap.ts:22)
        f = np.asarray(f, dtype=np.uint32)
        want = js.get_lengths(f.tolist(), limit, diag)
        got = oracle.get_lengths(f, limit)
        assert got.tolist() == want, (trial, nsym, limit)
        used = [l for l in want if l]
        if len(used) > 1:
            assert max(used) <= limit
            assert sum(2.0 ** -l for l in used) == 1.0                                    # Kraft-complete
    assert diag.get("oob_type", 0) == 0  # type[] never names a symbol >= symbols


def test_raw_deflate_bytes_match_model():
    for d in small_inputs():
        for ctype in (oracle.DYNAMIC, oracle.FIXED, oracle.NONE):
            want = js.raw_deflate(d, ctype)
            got = oracle.raw_deflate(d, ctype)
            assert got == want, (len(d), ctype, got.hex()[:60], want.hex()[:60])
            assert zlib.decompress(got, -15) == d
    # outputIndex / outputBuffer prefix is preserved (src/RawDeflate.ts:74-80)
    d = b"prefix test prefix test prefix"
    assert oracle.raw_deflate(d, oracle.DYNAMIC, prefix=b"\x78\x9c") == b"\x78\x9c" + oracle.raw_deflate(d)
    assert js.raw_deflate(d, 2, prefix=b"\x78\x9c") == b"\x78\x9c" + js.raw_deflate(d)


def test_appendix_c_known_answers():
    vec = json.load(open(os.path.join(GOLDEN, "appendix_c.json")))
    for v in vec["raw"]:
        d = bytes.fromhex(v["input_hex"])
        assert oracle.raw_deflate(d, oracle.DYNAMIC).hex() == v["dynamic_hex"], v["name"]
        assert oracle.raw_deflate(d, oracle.FIXED).hex() == v["fixed_hex"], v["name"]
        assert js.raw_deflate(d, 2).hex() == v["dynamic_hex"] and js.raw_deflate(d, 1).hex() == v["fixed_hex"]
    d = bytes(range(256))
    for ctype, (ln, sha) in ((oracle.DYNAMIC, vec["bytes_0_255"]["dynamic"]), (oracle.FIXED, vec["bytes_0_255"]["fixed"])):
        o = oracle.raw_deflate(d, ctype)
        assert len(o) == ln and hashlib.sha256(o).hexdigest()[:16] == sha
    from zlibts_b200 import synth
    for name, gen in (("text_65536_1", lambda: synth.text(65536, 1)), ("mixed_65536_2", lambda: synth.mixed(65536, 2))):
        v = vec[name]
        data = gen().tobytes()
        assert hashlib.sha256(data).hexdigest()[:16] == v["data_sha"]
        assert "%08x" % oracle.crc32(data) == v["crc32"] and "%08x" % oracle.adler32(data) == v["adler32"]
        o = oracle.raw_deflate(data)
        assert len(o) == v["dynamic_len"] and hashlib.sha256(o).hexdigest()[:16] == v["dynamic_sha"]
    # zlib-wrapped "a" (Zlib.Deflate): 78 9c + raw + Adler-32 big-endian
    z = b"\x78\x9c" + oracle.raw_deflate(b"a") + oracle.adler32(b"a").to_bytes(4, "big")
    assert z.hex() == vec["zlib_a_hex"] and zlib.decompress(z) == b"a"


def test_frozen_oracle_vectors():
    vec = json.load(open(os.path.join(GOLDEN, "oracle_vectors.json")))
    from make_vectors import inputs
    ins = inputs()
    assert len(ins) == len(vec["cases"])
    for (name, d), v in zip(ins, vec["cases"]):
        assert v["name"] == name and hashlib.sha256(d).hexdigest()[:16] == v["data_sha"]
        for key, ctype in (("dynamic", oracle.DYNAMIC), ("fixed", oracle.FIXED)):
            o = oracle.raw_deflate(d, ctype)
            assert [len(o), hashlib.sha256(o).hexdigest()[:16]] == v[key], (name, key)
        assert v["crc32"] == oracle.crc32(d) and v["adler32"] == oracle.adler32(d)


def test_inflate_matches_model_and_zlib():
    rng = np.random.default_rng(3)
    datas = small_inputs() + [rand_bytes(rng, 70000, 200).tobytes(), b"q" * 100000]
    for d in datas:
        for level in (0, 1, 6, 9):                      # stored, fixed and dynamic blocks from stock zlib
            s = zlib_raw(d, level)
            out, ip = oracle.raw_inflate(s + b"\0\0\0\0", 0, out_cap=len(d) + 16)
            assert out == d and ip == len(s)
            if len(d) <= 3000:
                mo, mip = js.raw_inflate(s + b"\0\0\0\0")
                assert mo == d and mip == ip
        if len(d) <= 3000:
            own = oracle.raw_deflate(d)
            out, ip = oracle.raw_inflate(b"\x78\x9c" + own + b"\0\0\0\0", 2, out_cap=len(d) + 16)  # `index` option
            assert out == d and ip == 2 + len(own)


def test_inflate_error_statuses():
    with pytest.raises(oracle.OracleError) as e:
        oracle.raw_inflate(b"\x07\x00\x00\x00\x00")    # BTYPE 3 (src/RawInflate.ts:168)
    assert e.value.code == 2
    s = zlib_raw(b"hello world hello world", 6)
    with pytest.raises(oracle.OracleError):
        oracle.raw_inflate(s[:5])                       # truncated
    with pytest.raises(js.InflateError):
        js.raw_inflate(b"\x07\x00\x00\x00\x00")


def test_reference_readbits_end_quirk_is_modelled():
    """readBits throws when ip + needed >= length even if the stream just fits."""
    d = b"abcabcabcabc"
    s = oracle.raw_deflate(d)
    out, ip = oracle.raw_inflate(s + b"\0", 0, out_cap=64, mirror_readbits_quirk=True)
    assert out == d and ip == len(s)
    mo, mip = js.raw_inflate(s + b"\0")
    assert mo == d and mip == len(s)


def test_join_marker_stream_decodes_in_reference_decoder():
    """chunks joined by an empty stored block form one stream the reference's RawInflate accepts."""
    rng = np.random.default_rng(4)
    for trial in range(20):
        parts = [rand_bytes(rng, int(rng.integers(1, 900)), int(rng.choice([2, 4, 26]))).tobytes()
                 for _ in range(int(rng.integers(2, 5)))]
        stream = bytearray()
        for k, p in enumerate(parts):
            c = bytearray(oracle.raw_deflate(p))
            last = k + 1 == len(parts)
            if not last:
                c[0] &= 0xFE
            stream += c
            if not last:
                # total bits of the block are not known from the bytes alone; re-derive by decoding
                _, used_bits = _decode_one_block_bits(bytes(c))
                pad = len(c) * 8 - used_bits
                stream += (b"" if pad >= 3 else b"\x00") + b"\x00\x00\xff\xff"
        whole = b"".join(parts)
        assert zlib.decompress(bytes(stream), -15) == whole
        out, ip = oracle.raw_inflate(bytes(stream) + b"\0\0\0\0", 0, out_cap=len(whole) + 8)
        assert out == whole and ip == len(stream)
        mo, mip = js.raw_inflate(bytes(stream) + b"\0\0\0\0")
        assert mo == whole and mip == len(stream)


def _decode_one_block_bits(block):
    """bits used by the single (non-final) block in `block`, via the model's bit reader."""
    r = js.RawInflate(block + b"\0\0\0\0\0\0\0\0")
    r.parse_block()
    return bytes(r.out), r.ip * 8 - r.bitsbuflen


def test_container_restatement_is_read_by_cpython():
    """oracle/containers.py (Deflate / GZip / Zip framing) against independent readers: zlib, gzip, zipfile."""
    import datetime
    import gzip
    import io
    import zipfile
    from oracle import containers as oc
    rng = np.random.default_rng(12)
    date = datetime.datetime(2026, 10, 18, 12, 34, 56)
    datas = [b"", b"a", rand_bytes(rng, 3000, 4).tobytes(), rand_bytes(rng, 9000, 256).tobytes()]
    for d in datas:
        for ctype in (oracle.DYNAMIC, oracle.FIXED, oracle.NONE):
            if ctype != oracle.DYNAMIC and not d:
                # upstream quirk: NONE of nothing writes no block at all (src/RawDeflate.ts:122-153) and FIXED of
                # nothing loses its end-of-block symbol (LZ77 output array of length 0, src/LZ77.ts:122,278)
                assert oc.zlib_stream(d, ctype).hex() == ("780100000001" if ctype == oracle.NONE else "785e0300000001")
                continue
            assert zlib.decompress(oc.zlib_stream(d, ctype)) == d
        g = oc.gzip_member(d, "name.bin", "note", True, 1234567890)
        assert gzip.decompress(g) == d and g[4:8] == (1234567890).to_bytes(4, "little")
    assert oc.zlib_stream(b"a").hex() == "789c05c081080000000020d6fd254e00620062"
    assert gzip.decompress(b"".join(oc.gzip_member(d) for d in datas)) == b"".join(datas)
    files = [{"name": "f%d" % i, "data": d, "date": date, "method": 8 if i % 2 == 0 else 0, "comment": "c%d" % i}
             for i, d in enumerate(datas)]
    arc = oc.zip_archive(files, b"zc")
    with zipfile.ZipFile(io.BytesIO(arc)) as zf:
        assert zf.testzip() is None and zf.comment == b"zc"
        for f in files:
            assert zf.read(f["name"]) == f["data"]
            assert zf.getinfo(f["name"]).date_time == (2026, 10, 18, 12, 34, 56)
            assert zf.getinfo(f["name"]).comment == f["comment"].encode()


def test_truncated_input_errors_equal_the_executed_reference():
    """tests/golden/truncation_vectors.json: what the reference itself (dist/Zlib-main.js under oracle/minijs) throws for
    EVERY cut of four streams -- 'input buffer is broken', 'invalid code length: N', '... header: LEN / NLEN'. The
    oracle, with the reference's readBits end check mirrored, gives the same message at every cut."""
    import json
    import os
    vec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "truncation_vectors.json")))["vectors"]
    n = 0
    for v in vec:
        s = bytes.fromhex(v["stream"])
        for cut, want in enumerate(v["cuts"]):
            try:
                out, ip = oracle.raw_inflate(s[:cut], 0, out_cap=4096, mirror_readbits_quirk=True)
                got = "OK %d %d" % (len(out), ip)
            except oracle.OracleError as e:
                got = "ERR " + str(e)
            assert got == want, (v["name"], cut, got, want)
            n += 1
    assert n > 1200


def test_inflate_of_hand_built_48_bit_symbols():
    """the stream builder of tests/test_inflate_gpu.py::test_matches_of_48_bits_each, CPU side: CPython zlib and the
    oracle's RawInflate read it alike"""
    import zlib
    import numpy as np
    from helpers import long_code_match_stream
    hist = np.random.default_rng(2).integers(0, 256, 40000, dtype=np.uint8).tobytes()
    s, expect = long_code_match_stream(hist, 500)
    assert zlib.decompressobj(-15).decompress(s) == expect
    out, ip = oracle.raw_inflate(s + b"\0\0\0\0", 0, out_cap=len(expect))
    assert out == expect and ip == len(s)
