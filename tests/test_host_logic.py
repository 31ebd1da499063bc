"""CPU: host-side logic that needs no GPU -- container glue of the API mirror, sharding plan, and the N>1 exchange
(world_size-2 gloo processes)."""
import datetime
import io
import os
import socket
import struct
import sys
import zipfile
import zlib

import numpy as np
import pytest


def test_zlib_and_gzip_header_bytes():
    import zlibts_b200 as z
    assert z.api._zlib_header(z.CompressionType.DYNAMIC) == b"\x78\x9c"
    assert z.api._zlib_header(z.CompressionType.FIXED) == b"\x78\x5e"
    assert z.api._zlib_header(z.CompressionType.NONE) == b"\x78\x01"
    for t in (0, 1, 2):
        h = z.api._zlib_header(t)
        assert ((h[0] << 8) + h[1]) % 31 == 0
    assert z.api._gzip_string("ałb") == b"a" + struct.pack("<H", 0x142) + b"b\0"   # src/GZip.ts:135-138
    assert z.api._dos_time(datetime.datetime(2026, 10, 18, 12, 34, 56)) == struct.pack(
        "<HH", (12 << 11) | (34 << 5) | 28, ((2026 - 1980) << 9) | (10 << 5) | 18)


def test_unzip_parses_central_directory_without_gpu():
    import zlibts_b200 as z
    bio = io.BytesIO()
    with zipfile.ZipFile(bio, "w", zipfile.ZIP_DEFLATED) as zf:
        zf.writestr("a.txt", b"hello" * 100)
        zf.writestr("dir/b.bin", bytes(range(256)))
        zf.comment = b"cmt"
    uz = z.Unzip(bio.getvalue())
    assert uz.getFilenames() == ["a.txt", "dir/b.bin"]
    assert uz.EOCD["totalEntries"] == 2 and uz.EOCD["comment"].tobytes() == b"cmt"
    assert uz.fileHeaderList[1]["plainSize"] == 256 and uz.fileHeaderList[0]["crc32"] == zlib.crc32(b"hello" * 100)
    with pytest.raises(z.ZlibError, match="End of Central Directory Record not found"):
        z.Unzip(b"\0" * 64).getFilenames()
    with pytest.raises(z.ZlibError, match="wrong index"):
        uz._local(5)


def test_option_validation_without_gpu():
    import zlibts_b200 as z
    with pytest.raises(z.ZlibError, match="lazy"):
        z.Deflate(b"abc", {"lazy": 4})
    with pytest.raises(z.ZlibError, match="unsupported compression method"):
        z.Inflate(b"\x77\x9c\x00")
    zp = z.Zip()
    zp.addFile(b"x", "x", {"password": "pw"})
    with pytest.raises(NotImplementedError):
        zp.compress()


def test_shard_plan_balances_and_covers():
    from zlibts_b200 import shard
    rng = np.random.default_rng(1)
    for world in (1, 2, 3, 4, 8):
        sizes = rng.integers(0, 300000, 1000)
        b = shard.plan_ranges(sizes, world)
        assert b[0] == 0 and b[-1] == 1000 and all(b[i] <= b[i + 1] for i in range(world))
        per = [int(sizes[b[i]:b[i + 1]].sum()) for i in range(world)]
        assert sum(per) == int(sizes.sum())
        assert max(per) - min(per) <= 2 * 300000
        r = shard.chunk_ranges(1_000_003, 65536, world)
        assert r[0][0] == 0 and r[-1][1] == 1_000_003 and all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        assert all(lo % 65536 == 0 for lo, _ in r)
    parts = []
    data = rng.integers(0, 256, 500000, dtype=np.uint8).tobytes()
    import zlibts_b200 as z
    for lo, hi in shard.chunk_ranges(len(data), 65536, 4):
        parts.append((zlib.crc32(data[lo:hi]), zlib.adler32(data[lo:hi]), hi - lo))
    assert shard.combine_checksums(parts, z.crc32_combine, z.adler32_combine) == (zlib.crc32(data), zlib.adler32(data), len(data))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


CHUNK_T = 4096  # chunk size of the two-rank test (the engine's chunk_bytes parameter; small keeps the CPU test fast)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import zlibts_b200 as z
    from zlibts_b200 import shard
    import oracle
    from oracle import js_model
    data = z.synth.mixed(CHUNK_T * 5 + 321, 77, 512)
    lo, hi = shard.chunk_ranges(data.size, CHUNK_T, world)[rank]
    # every rank "compresses" its own chunk range (the CPU oracle stands in for the GPU engine in this CPU test)
    # and joins its chunks exactly as the engine does
    mine = b""
    for c0 in range(lo, hi, CHUNK_T):
        blk = bytearray(oracle.raw_deflate(data[c0:min(c0 + CHUNK_T, hi)]))
        last = min(c0 + CHUNK_T, hi) == data.size
        if not last:
            blk[0] &= 0xFE
            r = js_model.RawInflate(bytes(blk) + bytes(8))
            r.parse_block()
            pad = len(blk) * 8 - (r.ip * 8 - r.bitsbuflen)
            blk += (b"" if pad >= 3 else b"\x00") + b"\x00\x00\xff\xff"
        mine += bytes(blk)
    off, total = shard.exclusive_scan(len(mine), rank, world, dist)
    parts = shard.gather_parts((zlib.crc32(data[lo:hi].tobytes()), zlib.adler32(data[lo:hi].tobytes()), hi - lo), rank,
                               world, dist)
    q.put((rank, off, total, mine, parts))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_gloo():
    """N = 2: shard by chunk range, exclusive scan of output sizes over gloo, stitch, decode as ONE stream."""
    import torch.multiprocessing as mp
    import zlibts_b200 as z
    from zlibts_b200 import shard
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    data = z.synth.mixed(CHUNK_T * 5 + 321, 77, 512).tobytes()
    total = got[0][2]
    out = bytearray(total)
    for rank, off, tot, mine, parts in got:
        assert tot == total
        out[off:off + len(mine)] = mine
    assert got[0][1] == 0 and got[1][1] == len(got[0][3])
    assert zlib.decompress(bytes(out), -15) == data
    crc, adler, n = shard.combine_checksums(got[0][4], z.crc32_combine, z.adler32_combine)
    assert (crc, adler, n) == (zlib.crc32(data), zlib.adler32(data), len(data))
