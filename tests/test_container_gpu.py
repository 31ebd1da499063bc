"""GPU: container assembly on the device (zlb_archive / zlb_archive_host) against the oracle's
restatement of Deflate.compress / GZip.compress / Zip.compress (oracle/containers.py), byte for byte for entries of
at most one chunk, and through CPython's zlib / gzip / zipfile readers for everything."""
import datetime
import gzip
import io
import struct
import zipfile
import zlib

import numpy as np
import pytest

import oracle
from oracle import containers as oc
from helpers import rand_bytes

pytestmark = pytest.mark.gpu

DATE = datetime.datetime(2026, 10, 18, 12, 34, 56)


@pytest.fixture(scope="module")
def Z(engine):
    import zlibts_b200 as z
    z.api.set_engine(engine)
    return z


def small_inputs():
    from zlibts_b200 import synth
    rng = np.random.default_rng(41)
    return [b"", b"a", b"abcabcabcabc" * 9, bytes(range(256)) * 3, rand_bytes(rng, 777, 3).tobytes(),
            rand_bytes(rng, 4099, 256).tobytes(), synth.text(65536, 5).tobytes(), synth.mixed(65536, 6, 512).tobytes(),
            synth.text(1, 7).tobytes(), synth.mixed(31111, 8, 256).tobytes()]


def test_zlib_streams_equal_the_oracle(Z):
    data = small_inputs()
    for ctype in (Z.CompressionType.DYNAMIC, Z.CompressionType.FIXED, Z.CompressionType.NONE):
        arc, res = Z.zlib_many(data, ctype)
        pos = 0
        for d, r in zip(data, res):
            assert int(r["in_used"]) == pos and int(r["status"]) == 0
            got = arc[pos:pos + int(r["out_len"])].tobytes()
            if not d and ctype == Z.CompressionType.FIXED:
                # deliberate deviation: upstream loses the end-of-block symbol of an empty FIXED block ("03", an
                # undecodable stream); the engine writes it ("03 00")
                assert got == b"\x78\x5e\x03\x00" + struct.pack(">I", 1)
            else:
                assert got == oc.zlib_stream(d, ctype)             # src/Deflate.ts:60-99
            assert int(r["adler32"]) == zlib.adler32(d)
            if d or ctype != Z.CompressionType.NONE:               # NONE of nothing is no block at all, as upstream
                assert zlib.decompress(got) == d
            pos += int(r["out_len"])
        assert pos == arc.size


def test_gzip_members_equal_the_oracle_and_concatenate(Z):
    data = small_inputs()
    arc, res = Z.gzip_many(data, mtime=1234567890)
    pos = 0
    for d, r in zip(data, res):
        got = arc[pos:pos + int(r["out_len"])].tobytes()
        assert got == oc.gzip_member(d, mtime=1234567890)          # src/GZip.ts:96-194
        assert int(r["crc32"]) == zlib.crc32(d)
        pos += int(r["out_len"])
    assert pos == arc.size
    assert gzip.decompress(arc.tobytes()) == b"".join(data)        # RFC 1952 multi-member
    gu = Z.GUnzip(arc)
    assert gu.decompress().tobytes() == b"".join(data) and len(gu.getMembers()) == len(data)
    # the class, with every header field (src/GZip.ts:108-156)
    for d in data[:6]:
        g = Z.GZip(d, {"filename": "nŁ.bin", "comment": "c", "hcrc": True, "b200": {"mtime": 77}})
        assert g.compress().tobytes() == oc.gzip_member(d, "nŁ.bin", "c", True, 77)


def _zip_case(rng):
    from zlibts_b200 import synth
    files = []
    for i in range(40):
        n = int(rng.integers(0, 50000))
        d = (synth.text(n, 200 + i) if i % 2 else synth.mixed(n, 200 + i, 512)).tobytes()
        f = {"name": "d/f%03d.bin" % i, "data": d, "date": DATE, "method": 0 if i % 7 == 3 else 8}
        if i % 5 == 1:
            f["comment"] = "entry %d" % i
        if i % 11 == 2:
            f["os"] = 3
        files.append(f)
    files.append({"name": "", "data": b"", "date": DATE, "method": 8})      # empty name, empty data
    files.append({"name": "stored-empty", "data": b"", "date": DATE, "method": 0})
    return files


def _add_all(zp, files):
    for f in files:
        opts = {"date": f["date"], "compressionMethod": f["method"]}
        if "comment" in f:
            opts["comment"] = f["comment"]
        if "os" in f:
            opts["os"] = f["os"]
        zp.addFile(f["data"], f["name"], opts)


def test_zip_archive_equals_the_oracle(Z):
    files = _zip_case(np.random.default_rng(42))
    zp = Z.Zip(b"archive comment")
    _add_all(zp, files)
    arc = zp.compress()
    want = oc.zip_archive(files, b"archive comment")                       # src/Zip.ts:80-372
    assert arc.tobytes() == want
    with zipfile.ZipFile(io.BytesIO(arc.tobytes())) as zf:
        assert zf.testzip() is None and zf.comment == b"archive comment"
        for f in files:
            if f["name"]:
                assert zf.read(f["name"]) == f["data"]
    # a second compress() finds every entry compressed already (src/Zip.ts:142) and writes the same archive
    assert zp.compress().tobytes() == want
    # entries with different deflate settings take the entry-by-entry route: same layout
    zp2 = Z.Zip()
    zp2.addFile(files[0]["data"], "a", {"date": DATE})
    zp2.addFile(files[1]["data"], "b", {"date": DATE, "deflateOptions": {"compressionType": Z.CompressionType.FIXED}})
    a2 = zp2.compress().tobytes()
    with zipfile.ZipFile(io.BytesIO(a2)) as zf:
        assert zf.read("a") == files[0]["data"] and zf.read("b") == files[1]["data"]
    # no entries: the end record alone
    assert Z.Zip(b"xy").compress().tobytes() == oc.zip_archive([], b"xy")


def test_zip_with_entries_larger_than_a_chunk(Z):
    from zlibts_b200 import synth
    files = [{"name": "big%d" % i, "data": synth.mixed(70000 + 150001 * i, 300 + i).tobytes(), "date": DATE,
              "method": 8 if i != 2 else 0} for i in range(5)]
    zp = Z.Zip()
    _add_all(zp, files)
    arc = zp.compress().tobytes()
    with zipfile.ZipFile(io.BytesIO(arc)) as zf:
        assert zf.testzip() is None
        for f in files:
            assert zf.read(f["name"]) == f["data"]
    uz = Z.Unzip(np.frombuffer(arc, dtype=np.uint8), {"verify": True})
    got = uz.decompressAll()
    assert {k: v.tobytes() for k, v in got.items()} == {f["name"]: f["data"] for f in files}


def test_archive_on_device_buffers_and_capacity_check(Z, engine):
    import torch
    data = small_inputs()
    lens = np.array([len(d) for d in data], dtype=np.uint64)
    blob = np.frombuffer(b"".join(data), dtype=np.uint8)
    head = b"\x78\x9c"
    ent = Z.make_entries(len(data))
    ent["in_off"] = np.concatenate([[0], np.cumsum(lens)[:-1]])
    ent["in_len"] = lens
    ent["head_off"], ent["head_len"] = 0, 2
    d_in = torch.from_numpy(blob.copy()).cuda()
    d_meta = torch.from_numpy(np.frombuffer(head, dtype=np.uint8).copy()).cuda()
    bound = Z.archive_bound(Z.FRAME_ZLIB, ent)
    d_out = torch.zeros(bound, dtype=torch.uint8, device="cuda")
    total, res = engine.archive(Z.FRAME_ZLIB, d_in, d_meta, ent, d_out)
    want = b"".join(oc.zlib_stream(d) for d in data)
    assert total == len(want) <= bound and d_out[:total].cpu().numpy().tobytes() == want
    assert int(res["out_len"].sum()) == total
    # too small an output: refused, nothing written, the size needed is reported in the error text
    d_small = torch.zeros(total - 1, dtype=torch.uint8, device="cuda")
    with pytest.raises(Z.EngineError, match="archive needs %d bytes" % total):
        engine.archive(Z.FRAME_ZLIB, d_in, d_meta, ent, d_small)
    assert int(d_small.max()) == 0
    with pytest.raises(Z.EngineError, match="unknown container kind"):
        engine.archive(9, d_in, d_meta, ent, d_out)


def test_fast_mode_archives_are_valid(Z):
    from zlibts_b200 import synth
    data = [synth.text(100000, 9).tobytes(), synth.mixed(65536, 10).tobytes()]
    arc, res = Z.gzip_many(data, mode=Z.mode_fast())
    assert gzip.decompress(arc.tobytes()) == b"".join(data)
    assert [int(c) for c in res["crc32"]] == [zlib.crc32(d) for d in data]


def test_zip32_limits_are_refused_not_wrapped(Z, engine):
    import torch
    n = 65536   # one more than the 16-bit entry count of the end record holds (src/Zip.ts:351-354)
    ent = Z.make_entries(n)
    ent["head_len"], ent["cdir_len"], ent["method"] = 30, 46, 0
    ent["cdir_off"] = 30
    meta = torch.zeros(30 + 46 + 22, dtype=torch.uint8, device="cuda")
    d_in = torch.zeros(16, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(Z.archive_bound(Z.FRAME_ZIP, ent, 22), dtype=torch.uint8, device="cuda")
    with pytest.raises(Z.EngineError, match="does not fit ZIP32"):
        engine.archive(Z.FRAME_ZIP, d_in, meta, ent, d_out, tail=(76, 22))
    total, _ = engine.archive(Z.FRAME_ZIP, d_in, meta, ent[:65535], d_out, tail=(76, 22))
    assert total == 65535 * 76 + 22
