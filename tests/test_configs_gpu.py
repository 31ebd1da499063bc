"""GPU: BASELINE.json's configs at reduced (fast) sizes, with the size-independent properties the full-size runs
use (tools/bench_configs.py runs them at full size; results under profiles/).

C2  chunk-parallel deflate of a mixed buffer: every chunk's bytes are the reference's; the whole is one stream.
C3  batched inflate of independent 64 KiB zlib streams produced by the reference-compatible encoder.
C4  Zip of many files + Unzip round trip with per-entry CRC-32.
C5  gzip member of a sharded buffer: deflate + CRC-32, marker-split inflate + CRC-32, checksum combine."""
import gzip
import os
import struct
import zlib

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
CH = 65536


def test_c2_chunks_are_reference_bytes(engine):
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import synth
    n_chunks = 96
    data = synth.mixed(n_chunks * CH, 2)
    slot = z.deflate_bound(CH)
    items = z.make_items(n_chunks)
    items["in_off"] = np.arange(n_chunks, dtype=np.uint64) * CH
    items["in_len"] = CH
    items["out_off"] = np.arange(n_chunks, dtype=np.uint64) * slot
    items["out_cap"] = slot
    d_in = torch.from_numpy(data).cuda()
    d_out = torch.zeros(n_chunks * slot, dtype=torch.uint8, device="cuda")
    r = engine.deflate_batch(d_in, d_out, items)
    h = d_out.cpu().numpy()
    assert int(r["status"].max()) == 0
    total_ref = 0
    for k in range(n_chunks):                                     # compat: byte identity per chunk
        want = oracle.raw_deflate(data[k * CH:(k + 1) * CH])
        got = h[k * slot:k * slot + int(r["out_len"][k])].tobytes()
        assert got == want, k
        total_ref += len(want)
    # the same buffer as ONE item: ratio equals the reference's per-chunk ratio up to the join markers
    one = z.make_items(1)
    one["in_len"], one["out_cap"] = data.size, z.deflate_bound(data.size)
    d_one = torch.zeros(int(one["out_cap"][0]), dtype=torch.uint8, device="cuda")
    r1 = engine.deflate_batch(d_in, d_one, one)
    clen = int(r1["out_len"][0])
    assert total_ref + 4 * (n_chunks - 1) <= clen <= total_ref + 5 * (n_chunks - 1)
    assert zlib.decompress(d_one[:clen].cpu().numpy().tobytes(), -15) == data.tobytes()


def test_c3_batched_inflate_of_zlib_streams(engine):
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import synth
    n = 512
    plain = np.empty(n * CH, dtype=np.uint8)
    streams = []
    for i in range(n):
        buf = plain[i * CH:(i + 1) * CH]
        if i % 2 == 0:
            synth.text(CH, 1000 + i, out=buf)
        else:
            synth.mixed(CH, 1000 + i, 4096, out=buf)
    # BASELINE: "zlib streams produced by the reference" -- every stream comes from the oracle's encoder (the reference's
    # algorithm, pinned to the executed reference); the GPU encoder must write the same bytes for every one of them
    slots, slot, ln = oracle.zlib_chunks_keep_mt(plain, CH, threads=os.cpu_count() or 1)
    blobs = [slots[i * slot:i * slot + int(ln[i])].tobytes() for i in range(n)]
    z.api.set_engine(engine)
    outs, res = z.deflate_many([plain[i * CH:(i + 1) * CH] for i in range(n)], want_adler32=True)
    for i in range(n):
        assert blobs[i] == b"\x78\x9c" + outs[i].tobytes() + struct.pack(">I", int(res["adler32"][i])), i
    lens = np.array([len(b) for b in blobs], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    blob = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    items = z.make_items(n)
    items["in_off"], items["in_len"] = offs + 2, lens - 2          # `index: 2` of src/Inflate.ts:61-66
    items["out_off"] = np.arange(n, dtype=np.uint64) * CH
    items["out_cap"] = CH
    d_o = torch.zeros(n * CH, dtype=torch.uint8, device="cuda")
    r = engine.inflate_batch(torch.from_numpy(blob.copy()).cuda(), d_o, items, z.INFLATE_WANT_ADLER32)
    assert int(r["status"].max()) == 0 and int(r["out_len"].min()) == CH
    assert np.array_equal(d_o.cpu().numpy(), plain)
    assert np.array_equal(r["in_used"].astype(np.uint64), lens - 6)   # `.ip` lands on the Adler-32
    assert np.array_equal(r["adler32"], res["adler32"])
    assert zlib.decompress(blobs[7]) == plain[7 * CH:8 * CH].tobytes()


def test_c4_zip_many_files_roundtrip(engine):
    import datetime
    import io
    import zipfile
    import zlibts_b200 as z
    from zlibts_b200 import synth
    z.api.set_engine(engine)
    x, files = 4, {}
    for i in range(600):
        x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
        n = int(256 * 2 ** (((x >> 16) % 81) / 8))
        files["f%05d.bin" % i] = (synth.text(n, 4000 + i) if i % 2 == 0 else synth.mixed(n, 4000 + i)).tobytes()
    zp = z.Zip()
    for name, d in files.items():
        zp.addFile(d, name, {"date": datetime.datetime(2026, 10, 18)})
    arc = zp.compress()
    out = z.Unzip(arc, {"verify": True}).decompressAll()
    assert {k: v.tobytes() for k, v in out.items()} == files
    with zipfile.ZipFile(io.BytesIO(arc.tobytes())) as zf:
        assert zf.testzip() is None and len(zf.namelist()) == 600


def test_c5_gzip_member_sharded(engine):
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import shard, synth
    shards = 24
    n = shards << 20
    h = np.empty(n, dtype=np.uint8)
    for s in range(shards):
        synth.mixed(1 << 20, 5000 + s, 4096, out=h[s << 20:(s + 1) << 20])
    d_in = torch.from_numpy(h).cuda()
    # three "ranks": contiguous chunk ranges, every rank but the last leaves the stream open
    pieces, parts = [], []
    for r, (lo, hi) in enumerate(shard.chunk_ranges(n, CH, 3)):
        it = z.make_items(1)
        it["in_off"], it["in_len"], it["out_cap"] = lo, hi - lo, z.deflate_bound(hi - lo)
        d_z = torch.zeros(int(it["out_cap"][0]), dtype=torch.uint8, device="cuda")
        res = engine.deflate_batch(d_in, d_z, it, flags=z.DEFLATE_WANT_CRC32 | (0 if r == 2 else z.DEFLATE_NOT_FINAL))
        assert int(res["status"][0]) == 0
        pieces.append(d_z[:int(res["out_len"][0])].cpu().numpy().tobytes())
        parts.append((int(res["crc32"][0]), 1, hi - lo))
    crc, _, total = shard.combine_checksums(parts, z.crc32_combine, z.adler32_combine)
    assert crc == zlib.crc32(h) and total == n
    member = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" + b"".join(pieces) + struct.pack("<II", crc, n & 0xFFFFFFFF)
    assert gzip.decompress(member) == h.tobytes()                  # CPython reads the stitched member
    z.api.set_engine(engine)
    gu = z.GUnzip(member)                                          # our reader: marker-split inflate + CRC-32 + ISIZE
    assert gu.decompress().tobytes() == h.tobytes() and gu.crc32 == crc
