"""GPU: the zlib.ts-shaped host API (zlibts_b200.api) end to end, cross-checked with CPython's zlib / gzip /
zipfile (independent RFC 1950/1952/PKZIP implementations) and the oracle's RawDeflate bytes.

Mirrors what the reference's manual pages do (test/Deflate.html:45-63, test/GZip.html:45-66, test/Zip.html:45-69:
compress -> decompress -> compare) and adds byte-level checks those pages lack."""
import datetime
import gzip
import io
import struct
import zipfile
import zlib

import numpy as np
import pytest

import oracle
from helpers import rand_bytes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Z(engine):
    import zlibts_b200 as z
    z.api.set_engine(engine)
    return z


def samples():
    from zlibts_b200 import synth
    rng = np.random.default_rng(31)
    return [b"a", b"hello hello hello hello", bytes(range(256)), rand_bytes(rng, 3000, 4).tobytes(),
            synth.text(65536, 1).tobytes(), synth.mixed(65536, 2).tobytes(), synth.text(200000, 3).tobytes(),
            synth.mixed(300001, 4).tobytes()]


def test_deflate_inflate_zlib_container(Z):
    for d in samples():
        for ctype, hdr in ((Z.CompressionType.DYNAMIC, b"\x78\x9c"), (Z.CompressionType.FIXED, b"\x78\x5e"),
                           (Z.CompressionType.NONE, b"\x78\x01")):
            df = Z.Deflate(d, {"compressionType": ctype})
            c = df.compress().tobytes()
            assert c[:2] == hdr and df.adler32 == zlib.adler32(d)
            assert zlib.decompress(c) == d                                   # RFC 1950 valid
            if len(d) <= 65536:                                              # == reference bytes (src/Deflate.ts:60-99)
                assert c == hdr + oracle.raw_deflate(d, ctype) + struct.pack(">I", oracle.adler32(d))
            inf = Z.Inflate(np.frombuffer(c, dtype=np.uint8), {"verify": True})
            assert inf.decompress().tobytes() == d
            assert inf.ip == len(c) - 4 and inf.adler32 == zlib.adler32(d)
    # Appendix C: Zlib.Deflate of "a"
    assert Z.Deflate(b"a").compress().tobytes().hex() == "789c05c081080000000020d6fd254e00620062"
    assert Z.Deflate.compress_static(b"a", {}).tobytes().hex() == "789c05c081080000000020d6fd254e00620062"


def test_inflate_reads_stock_zlib_and_reports_errors(Z):
    d = samples()[4]
    for level in (0, 1, 6, 9):
        c = zlib.compress(d, level)
        assert Z.Inflate(c, {"verify": True}).decompress().tobytes() == d
    c = bytearray(zlib.compress(d))
    c[-1] ^= 1
    with pytest.raises(Z.ZlibError, match="invalid adler-32 checksum"):
        Z.Inflate(bytes(c), {"verify": True}).decompress()
    assert Z.Inflate(bytes(c)).decompress().tobytes() == d                   # verify is opt-in (src/Inflate.ts:39,81)
    with pytest.raises(Z.ZlibError, match="unsupported compression method"):
        Z.Inflate(b"\x77\x9c\x00")
    with pytest.raises(Z.ZlibError, match="invalid fcheck flag"):
        Z.Inflate(b"\x78\x9d\x00")
    with pytest.raises(Z.ZlibError, match="fdict flag is not supported"):
        Z.Inflate(bytes([0x78, 0x20 | (31 - (0x7820 % 31))]) + b"\0\0\0\0")
    with pytest.raises(Z.ZlibError, match="unknown BTYPE"):
        Z.Inflate(b"\x78\x9c\x07\x00\x00\x00\x00\x00").decompress()
    # `index` option: container embedded at an offset
    blob = b"junk!" + zlib.compress(b"payload payload payload")
    assert Z.Inflate(blob, {"index": 5, "verify": True}).decompress().tobytes() == b"payload payload payload"


def test_raw_deflate_output_buffer_contract(Z):
    d = b"prefix contract " * 20
    r = Z.RawDeflate(d, {"outputBuffer": b"HEAD....", "outputIndex": 4})
    out = r.compress().tobytes()
    assert out[:4] == b"HEAD" and out[4:] == oracle.raw_deflate(d) and r.op == len(out)
    ri = Z.RawInflate(out + b"\0\0\0\0", {"index": 4})
    assert ri.decompress().tobytes() == d and ri.ip == len(out)
    with pytest.raises(Z.ZlibError):
        Z.RawDeflate(d, {"lazy": 8})


def test_gzip_gunzip(Z):
    for d in samples():
        g = Z.GZip(d, {"filename": "f.bin", "comment": "made on a B200", "hcrc": True, "b200": {"mtime": 1234567890}})
        c = g.compress().tobytes()
        assert gzip.decompress(c) == d and g.crc32 == zlib.crc32(d)
        assert c[:4] == b"\x1f\x8b\x08\x1a" and c[4:8] == struct.pack("<I", 1234567890) and c[8:10] == b"\x00\x03"
        assert c[-8:] == struct.pack("<II", zlib.crc32(d), len(d))
        gu = Z.GUnzip(np.frombuffer(c, dtype=np.uint8))
        assert gu.decompress().tobytes() == d
        m = gu.getMembers()[0]
        assert m["name"] == "f.bin" and m["comment"] == "made on a B200" and m["mtime"] == 1234567890
    # stock gzip output, multi-member (src/GUnzip.ts:56-58)
    a, b = samples()[4], samples()[1]
    c = gzip.compress(a, mtime=0) + gzip.compress(b, mtime=0)
    gu = Z.GUnzip(c)
    assert gu.decompress().tobytes() == a + b and len(gu.getMembers()) == 2
    bad = bytearray(gzip.compress(a))
    bad[-6] ^= 0x10
    with pytest.raises(Z.ZlibError, match="invalid CRC-32 checksum"):
        Z.GUnzip(bytes(bad)).decompress()
    with pytest.raises(Z.ZlibError, match="invalid file signature"):
        Z.GUnzip(b"\x1f\x8c" + bytes(20)).decompress()


def test_zip_unzip_roundtrip_and_stdlib_interop(Z):
    rng = np.random.default_rng(32)
    from zlibts_b200 import synth
    files = {}
    for i in range(60):
        n = int(rng.integers(0, 40000))
        files["dir/f%03d.bin" % i] = (synth.text(n, 100 + i) if i % 2 else synth.mixed(n, 100 + i, 512)).tobytes()
    files["big.bin"] = synth.mixed(200000, 7).tobytes()
    date = datetime.datetime(2026, 10, 18, 12, 34, 56)
    zp = Z.Zip(b"archive comment")
    for k, (name, data) in enumerate(files.items()):
        opts = {"date": date}
        if k % 7 == 3:
            opts["compressionMethod"] = Z.ZipCompressionMethod.STORE
        if k % 5 == 1:
            opts["comment"] = "entry %d" % k
        zp.addFile(data, name, opts)
    arc = zp.compress().tobytes()
    # CPython's zipfile reads it: names, data, CRCs, timestamps
    with zipfile.ZipFile(io.BytesIO(arc)) as zf:
        assert zf.testzip() is None and zf.comment == b"archive comment"
        assert zf.namelist() == list(files)
        for name, data in files.items():
            assert zf.read(name) == data
            assert zf.getinfo(name).date_time == (2026, 10, 18, 12, 34, 56)
    # our Unzip reads it, entry by entry and as one batch, with CRC verification
    uz = Z.Unzip(np.frombuffer(arc, dtype=np.uint8), {"verify": True})
    assert uz.getFilenames() == list(files)
    assert uz.decompress("big.bin").tobytes() == files["big.bin"]
    allf = uz.decompressAll()
    assert {k: v.tobytes() for k, v in allf.items()} == files
    with pytest.raises(Z.ZlibError, match="not found"):
        uz.decompress("nope")
    # an archive written by zipfile (deflated + stored entries)
    bio = io.BytesIO()
    with zipfile.ZipFile(bio, "w") as zf:
        for k, (name, data) in enumerate(list(files.items())[:20]):
            zf.writestr(zipfile.ZipInfo(name, (2020, 1, 2, 3, 4, 6)), data,
                        zipfile.ZIP_DEFLATED if k % 3 else zipfile.ZIP_STORED)
    uz = Z.Unzip(bio.getvalue(), {"verify": True})
    for name in uz.getFilenames():
        assert uz.decompress(name).tobytes() == files[name]
    # CRC mismatch is reported only with verify (src/Unzip.ts:293-301)
    broken = bytearray(arc)
    lh_crc = arc.index(b"PK\x03\x04") + 14
    broken[lh_crc] ^= 0xFF
    with pytest.raises(Z.ZlibError, match="Incorrect crc"):
        Z.Unzip(bytes(broken), {"verify": True}).getFileData(0)
    Z.Unzip(bytes(broken)).getFileData(0)


def test_checksum_classes(Z):
    rng = np.random.default_rng(33)
    d = rand_bytes(rng, 100000).tobytes()
    assert Z.CRC32.create(d) == zlib.crc32(d) and Z.Adler32.create(d) == zlib.adler32(d)
    assert Z.CRC32.create(d, 10, 500) == zlib.crc32(d[10:510])
    assert Z.CRC32.update(d[5000:], Z.CRC32.create(d[:5000])) == zlib.crc32(d)
    assert Z.Adler32.update(Z.Adler32.create(d[:777]), d[777:]) == zlib.adler32(d)
    assert Z.Adler32.create("abc") == zlib.adler32(b"abc")


def test_config_c1_text_1mib_roundtrip(Z):
    """BASELINE config 1: Zlib.Deflate / Inflate round trip of 1 MiB synthetic text (dynamic Huffman)."""
    from zlibts_b200 import synth
    d = synth.text(1 << 20, 1).tobytes()
    c = Z.Deflate(d).compress().tobytes()
    assert zlib.decompress(c) == d
    out = Z.Inflate(c, {"verify": True}).decompress().tobytes()
    assert out == d
    # 16 chunks: every chunk's bytes are the reference's for that chunk, so the size is the sum + join markers
    ref = sum(len(oracle.raw_deflate(d[k << 16:(k + 1) << 16])) for k in range(16))
    assert ref + 15 * 4 <= len(c) - 6 <= ref + 15 * 5
    # the reference's own decoder (oracle restatement) accepts the joined stream
    o2, ip = oracle.raw_inflate(c, 2, out_cap=len(d))
    assert o2 == d and ip == len(c) - 4


def test_fast_mode_option_through_the_api(Z):
    from zlibts_b200 import synth
    d = synth.text(300000, 77).tobytes()
    slow = Z.Deflate(d).compress().tobytes()
    fast = Z.Deflate(d, {"b200": {"mode": "fast"}}).compress().tobytes()
    assert zlib.decompress(fast) == d and fast != slow and len(fast) <= 1.03 * len(slow)
    assert Z.Inflate(fast, {"verify": True}).decompress().tobytes() == d
    g = Z.GZip(d, {"deflateOptions": {"b200": {"mode": "fast", "depth": 8}}}).compress().tobytes()
    assert gzip.decompress(g) == d


def test_primed_mode_option_through_the_api(Z):
    from zlibts_b200 import synth
    d = synth.text(400000, 78).tobytes()
    plain = Z.Deflate(d).compress().tobytes()
    for name in ("primed", "fast-primed"):
        c = Z.Deflate(d, {"b200": {"mode": name}}).compress().tobytes()
        assert zlib.decompress(c) == d and Z.Inflate(c, {"verify": True}).decompress().tobytes() == d
        if name == "primed":
            assert len(c) < 0.97 * len(plain)
    g = Z.GZip(d, {"deflateOptions": {"b200": {"mode": "primed"}}}).compress().tobytes()
    assert gzip.decompress(g) == d and Z.GUnzip(g).decompress().tobytes() == d
    zp = Z.Zip()
    zp.addFile(d, "t.txt", {"deflateOptions": {"b200": {"mode": "primed"}}})
    with zipfile.ZipFile(io.BytesIO(zp.compress().tobytes())) as zf:
        assert zf.read("t.txt") == d
    with pytest.raises(Z.ZlibError, match="unknown b200 mode"):
        Z.Deflate(d, {"b200": {"mode": "turbo"}}).compress()


def test_gunzip_of_1024_members_is_one_batch(Z):
    """src/GUnzip.ts:56-58 loops over the members; here every `1F 8B 08` in the buffer is a candidate, all candidates are
    inflated in ONE call and the chain of true members is walked on the host. 1024 members (stock gzip + this engine's
    own, signature bytes inside the data as decoys) cost a handful of launches, not a thousand."""
    rng = np.random.default_rng(77)
    parts, blob = [], []
    for k in range(1024):
        n = int(rng.integers(1, 3000))
        d = rng.integers(0, 6, n, dtype=np.uint8).tobytes() + b"\x1f\x8b\x08\x00decoy" * int(k % 3 == 0)
        parts.append(d)
        blob.append(gzip.compress(d, mtime=k) if k % 2 else Z.GZip(d).compress().tobytes())
    arc = b"".join(blob)
    eng = Z.api.engine()
    l0 = eng.launch_count
    gu = Z.GUnzip(arc)
    out = gu.decompress().tobytes()
    launches = eng.launch_count - l0
    assert out == b"".join(parts)
    ms = gu.getMembers()
    assert len(ms) == 1024 and all(m["data"].tobytes() == d for m, d in zip(ms, parts))
    assert ms[3]["mtime"] == 3 and ms[5]["isize"] == len(parts[5]) and ms[7]["crc32"] == zlib.crc32(parts[7])
    assert launches < 64, launches
    # a member whose trailer is wrong is reported like the reference's loop would: at that member, same message
    bad = bytearray(arc)
    off = sum(len(b) for b in blob[:500]) + len(blob[500]) - 7
    bad[off] ^= 0x55
    with pytest.raises(Z.ZlibError, match="invalid CRC-32 checksum"):
        Z.GUnzip(bytes(bad)).decompress()
