"""CPU: the C oracle against THE REFERENCE ITSELF.

tests/golden/refjs_vectors.json holds outputs of /root/reference/dist/Zlib-main.js, unmodified, executed by
oracle/minijs (a JavaScript interpreter written as test infrastructure; tests/golden/make_refjs_vectors.py is the
generating script). The first group of tests compares the oracle with those committed vectors and runs everywhere
(the GPU box has no /root/reference). The second group runs the interpreter live, where the reference sources are
present, on fresh inputs -- a differential fuzz of oracle.raw_deflate / raw_inflate / get_lengths / containers
against the reference's own RawDeflate / RawInflate / getLengths / Deflate / GZip / Zip."""
import datetime
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

import oracle
from oracle import containers, refjs
from helpers import rand_bytes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VEC = json.load(open(os.path.join(GOLDEN, "refjs_vectors.json")))
live = pytest.mark.skipif(not refjs.available(), reason="the reference sources (/root/reference) are not on this machine")


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()[:16]


# ---- against the committed vectors (made by the reference under the interpreter) ---------------------------------
def test_appendix_c_regenerated_by_the_reference_equals_the_survey_model():
    """SURVEY.md Appendix C was a model's prediction; the reference, executed, gives exactly those bytes."""
    survey = json.load(open(os.path.join(GOLDEN, "appendix_c.json")))
    ref = VEC["appendix_c"]
    assert [(r["name"], r["dynamic_hex"], r["fixed_hex"]) for r in ref["raw"]] == \
           [(r["name"], r["dynamic_hex"], r["fixed_hex"]) for r in survey["raw"]]
    assert ref["bytes_0_255"] == survey["bytes_0_255"]
    assert ref["text_65536_1"] == survey["text_65536_1"] and ref["mixed_65536_2"] == survey["mixed_65536_2"]
    assert ref["zlib_a_hex"] == survey["zlib_a_hex"]


def test_oracle_equals_reference_on_the_small_known_answers():
    for r in VEC["appendix_c"]["raw"]:
        d = bytes.fromhex(r["input_hex"])
        assert oracle.raw_deflate(d, oracle.DYNAMIC).hex() == r["dynamic_hex"], r["name"]
        assert oracle.raw_deflate(d, oracle.FIXED).hex() == r["fixed_hex"], r["name"]
    b = bytes(range(256))
    for key, ct in (("dynamic", oracle.DYNAMIC), ("fixed", oracle.FIXED)):
        o = oracle.raw_deflate(b, ct)
        assert [len(o), sha(o)] == VEC["appendix_c"]["bytes_0_255"][key]


def test_oracle_equals_reference_on_240_fuzz_inputs():
    import make_refjs_vectors as mk
    ins = mk.fuzz_inputs()
    assert len(ins) == len(VEC["fuzz"]) >= 200
    for (rec, d), v in zip(ins, VEC["fuzz"]):
        assert rec["i"] == v["i"] and len(d) == v["n"] and sha(d) == v["data_sha"], "input recipe drifted"
        for key, ct in (("dynamic", oracle.DYNAMIC), ("fixed", oracle.FIXED)):
            o = oracle.raw_deflate(d, ct)
            assert [len(o), sha(o)] == v[key], (v["i"], v["kind"], v["n"], key)


def test_oracle_equals_reference_on_benchmark_chunks_and_one_mib():
    """64 KiB chunks of the BASELINE generators, adversarial chunks, and the 1 MiB text of config C1 as ONE block
    (frequencies beyond 65535 wrap in the reference's Uint16 heap, SURVEY B-3: the oracle wraps identically)."""
    import make_refjs_vectors as mk
    for (name, d), v in zip(mk.chunk_inputs(), VEC["chunks"]):
        assert name == v["name"] and sha(d) == v["data_sha"]
        assert oracle.crc32(d) == v["crc32"] == zlib.crc32(d)
        assert oracle.adler32(d) == v["adler32"] == zlib.adler32(d)
        o = oracle.raw_deflate(d, oracle.DYNAMIC)
        assert [len(o), sha(o)] == v["dynamic"], name
        assert zlib.decompress(o, -15) == d


def test_oracle_code_lengths_equal_reference_getLengths():
    for c in VEC["lengths"]:
        got = oracle.get_lengths(np.array(c["freqs"], dtype=np.uint32), c["limit"])
        assert got.tolist() == c["lengths"], c["k"]


def test_oracle_inflate_ip_equals_reference_rawinflate():
    import make_refjs_vectors  # noqa: F401  (keeps the generator importable)
    from zlibts_b200 import synth
    for c in VEC["inflate"]:
        k, n = c["k"], c["n"]
        d = synth.mixed(n, 7000 + k, 256).tobytes() if k % 2 else synth.text(n, 7000 + k).tobytes()
        co = zlib.compressobj(c["level"], zlib.DEFLATED, -15, 9, c["strategy"])
        s = co.compress(d) + co.flush()
        assert sha(s) == c["stream_sha"] and sha(d) == c["data_sha"]
        out, ip = oracle.raw_inflate(s + b"\0\0\0\0")
        assert out == d and ip == c["ip"], k


def test_container_restatement_equals_reference_containers():
    from zlibts_b200 import synth
    ins = {"hello*4": b"hello hello hello hello", "text_5000_9": synth.text(5000, 9).tobytes()}
    for name, d in ins.items():
        v = VEC["containers"][name]
        z = containers.zlib_stream(d)
        assert [len(z), sha(z)] == v["zlib"]
        g = containers.gzip_member(d, mtime=0)
        assert [len(g), sha(g)] == v["gzip_mtime_masked"]
    date = datetime.datetime.fromtimestamp(1700000000, datetime.timezone.utc)
    files = [{"name": "a.txt", "data": ins["hello*4"], "date": date}, {"name": "dir/b.bin", "data": ins["text_5000_9"], "date": date},
             {"name": "empty", "data": b"", "date": date}]
    za = containers.zip_archive(files)
    v = VEC["containers"]["zip_3_files_date_1700000000000"]
    assert len(za) == v["len"] and sha(za) == v["sha"], za[:64].hex() + " vs " + v["hex_head"]


# ---- live: the interpreter runs the bundle on fresh inputs -----------------------------------------------------
@live
def test_live_differential_fuzz_raw_deflate():
    rng = np.random.default_rng(int.from_bytes(os.urandom(4), "little"))
    ins = []
    for k in range(120):
        n = int(rng.integers(1, 3000))
        a = [2, 3, 4, 16, 256][k % 5]
        ins.append(rand_bytes(rng, n, a).tobytes())
    ins += [b"a" * 300, b"ab" * 700, bytes(range(256)) * 3, b"\0" * 5000 + b"\1" * 5000]
    b = refjs.Batch()
    for d in ins:
        b.add("rawdeflate", d, 2, 0)
        b.add("rawdeflate", d, 1, 0)
        b.add("rawdeflate", d, 0, 0)
    res = b.run()
    for i, d in enumerate(ins):
        for j, ct in enumerate((oracle.DYNAMIC, oracle.FIXED, oracle.NONE)):
            r = res[3 * i + j]
            assert not isinstance(r, refjs.RefError), r
            assert r[1] == oracle.raw_deflate(d, ct), (len(d), ct, d[:40].hex())


@live
def test_live_reference_decodes_oracle_and_joined_streams():
    """Round trip through the reference's own RawInflate / Inflate / GUnzip / Unzip: oracle-made streams and a
    sync-joined multi-chunk stream (what the engine writes for inputs > one chunk)."""
    from zlibts_b200 import synth
    d1, d2, d3 = synth.text(3000, 21).tobytes(), synth.mixed(2500, 22, 128).tobytes(), b"z" * 777
    joined = b""
    for k, d in enumerate((d1, d2, d3)):
        blk = bytearray(oracle.raw_deflate(d))
        if k < 2:
            blk[0] &= 0xFE                                   # BFINAL = 0
            pad = (8 - (oracle_bits(d) % 8)) % 8              # zero padding bits of the last byte
            joined += bytes(blk) + (b"" if pad >= 3 else b"\0") + b"\0\0\xff\xff"
        else:
            joined += bytes(blk)
    whole = d1 + d2 + d3
    b = refjs.Batch()
    b.add("rawinflate", joined + b"\0\0\0\0", 0, 1)
    b.add("inflate", containers.zlib_stream(d1), 1)
    b.add("gunzip", containers.gzip_member(d2, mtime=5) + containers.gzip_member(d3, mtime=6))
    date = datetime.datetime(2024, 5, 6, 7, 8, 10)
    b.add_unzip(containers.zip_archive([{"name": "x", "data": d1, "date": date}, {"name": "y", "data": d3, "date": date}]))
    b.add("inflate", containers.zlib_stream(d1)[:-1] + b"\0", 1)   # broken Adler-32
    r = b.run()
    assert r[0][1] == whole and int(r[0][0]["ip"]) == len(joined)
    assert r[1][1] == d1
    assert r[2][1] == d2 + d3 and r[2][0]["members"] == "2"
    assert r[3][1] == [d1, d3]
    assert isinstance(r[4], refjs.RefError) and "invalid adler-32 checksum" in str(r[4])


def oracle_bits(d):
    """Bits of the oracle's DYNAMIC block for d: decode it with zlib and ask how many bits were left unused."""
    blk = oracle.raw_deflate(d)
    do = zlib.decompressobj(-15)
    do.decompress(blk)
    # the stream ends inside the last byte; its length in bits is found by flipping to a non-final block + probing the
    # padding: simpler and exact -- count from the oracle's own bit writer via the difference of two encodings
    # (a block followed by one more zero byte decodes iff the padding is >= 3 bits of an empty stored block header).
    for pad in range(8):
        test = bytearray(blk)
        test[0] &= 0xFE
        tail = (b"" if pad >= 3 else b"\0") + b"\0\0\xff\xff" + bytes(oracle.raw_deflate(b"q"))
        try:
            if zlib.decompress(bytes(test) + tail, -15) == d + b"q":
                # ambiguous between "pad >= 3" and "< 3" only through the extra zero byte: accept the first that decodes
                return len(blk) * 8 - pad
        except zlib.error:
            continue
    raise AssertionError("no padding assumption decodes")


@live
def test_live_code_lengths_and_checksums():
    rng = np.random.default_rng(99)
    b = refjs.Batch()
    cases = []
    for k in range(40):
        nsym, limit = [(286, 15), (30, 7), (19, 7)][k % 3]
        f = (rng.integers(0, 60000, nsym) * (rng.random(nsym) < 0.5)).astype(np.uint32)
        if nsym == 19:
            f %= 256
        cases.append((f, limit))
        b.add("lengths", f.astype("<u4").tobytes(), limit)
    datas = [rand_bytes(rng, n).tobytes() for n in (1, 7, 8, 9, 5552, 5553, 70000)]
    for d in datas:
        b.add_value("crc32", d)
        b.add_value("adler32", d)
    res = b.run()
    for k, (f, limit) in enumerate(cases):
        assert list(res[k][1]) == oracle.get_lengths(f, limit).tolist(), k
    for i, d in enumerate(datas):
        assert int(res[40 + 2 * i][0]["value"]) == oracle.crc32(d) == zlib.crc32(d)
        assert int(res[41 + 2 * i][0]["value"]) == oracle.adler32(d) == zlib.adler32(d)


@live
def test_live_reference_bug_b1_is_real():
    """SURVEY B-1: Zlib.Deflate.compress throws a RangeError once the stream outgrows 32 KiB. The host mirror
    (zlib.ts_b200/api.py) implements the evident intent instead; this pins that the deviation is the reference's bug."""
    from zlibts_b200 import synth
    b = refjs.Batch()
    b.add("deflate", synth.text(200, 1).tobytes(), 2)
    b.add("deflate", rand_bytes(np.random.default_rng(5), 40000).tobytes(), 2)
    r = b.run()
    assert not isinstance(r[0], refjs.RefError)
    assert isinstance(r[1], refjs.RefError)
