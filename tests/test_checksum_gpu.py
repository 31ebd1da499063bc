"""GPU: slice-and-combine CRC-32 / Adler-32 kernels vs the oracle (src/CRC32.ts, src/Adler32.ts) and CPython zlib."""
import zlib

import numpy as np
import pytest

import oracle
from helpers import pack

pytestmark = pytest.mark.gpu


def _run(engine, buffers, align=1, host=False):
    import torch
    import zlibts_b200 as z
    blob, offs, lens = pack(buffers, align)
    items = z.make_items(len(buffers))
    items["in_off"] = offs
    items["in_len"] = lens
    if host:
        return engine.checksum_batch_host(blob, items)
    return engine.checksum_batch(torch.from_numpy(blob).cuda(), items)


def test_edge_lengths(engine):
    rng = np.random.default_rng(1)
    sizes = [0, 1, 2, 3, 15, 16, 17, 31, 32, 33, 255, 256, 257, 4095, 4096, 4097, 65535, 65536, 65537,
             262143, 262144, 262145, 524288 + 5, 1048576 + 77]
    bufs = [rng.integers(0, 256, n, dtype=np.uint8).tobytes() for n in sizes]
    for align in (1, 16):
        res = _run(engine, bufs, align)
        for b, r in zip(bufs, res):
            assert int(r["crc32"]) == zlib.crc32(b) == oracle.crc32(b), len(b)
            assert int(r["adler32"]) == zlib.adler32(b) == oracle.adler32(b), len(b)


def test_extreme_bytes(engine):
    bufs = [b"\xff" * n for n in (1, 5551, 5552, 5553, 70000, 300000)] + [b"\x00" * 100000, b"a"]
    res = _run(engine, bufs)
    for b, r in zip(bufs, res):
        assert int(r["crc32"]) == zlib.crc32(b)
        assert int(r["adler32"]) == zlib.adler32(b)


def test_many_small_items_and_host_path(engine):
    rng = np.random.default_rng(2)
    bufs = [rng.integers(0, 256, int(n), dtype=np.uint8).tobytes() for n in rng.integers(0, 3000, 2000)]
    for host in (False, True):
        res = _run(engine, bufs, host=host)
        for b, r in zip(bufs, res):
            assert int(r["crc32"]) == zlib.crc32(b)
            assert int(r["adler32"]) == zlib.adler32(b)


def test_large_buffer_property(engine):
    """64 MiB: checksum of the whole equals the combine of the checksums of its shards."""
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import synth
    n = 64 << 20
    data = synth.mixed(n, 77)
    d = torch.from_numpy(data).cuda()
    whole = z.make_items(1)
    whole["in_len"] = n
    r = engine.checksum_batch(d, whole)
    assert int(r["crc32"][0]) == zlib.crc32(data)
    assert int(r["adler32"][0]) == zlib.adler32(data)
    shards = z.make_items(8)
    shards["in_off"] = np.arange(8) * (n // 8)
    shards["in_len"] = n // 8
    rs = engine.checksum_batch(d, shards)
    crc, ad = int(rs["crc32"][0]), int(rs["adler32"][0])
    for k in range(1, 8):
        crc = z.crc32_combine(crc, int(rs["crc32"][k]), n // 8)
        ad = z.adler32_combine(ad, int(rs["adler32"][k]), n // 8)
    assert crc == int(r["crc32"][0]) and ad == int(r["adler32"][0])
