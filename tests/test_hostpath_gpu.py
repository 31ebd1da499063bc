"""GPU: the "_host" entry points of the C ABI -- page-locked and pageable caller buffers, exact copy-back, the bounded
marker split. The reference has no such seam (it works on JS heap arrays, src/RawDeflate.ts:52-58); what is held
here is that whatever memory a binding passes, the bytes are those of the device path."""
import zlib

import numpy as np
import pytest

from helpers import rand_bytes, zlib_raw

pytestmark = pytest.mark.gpu


def _items(z, in_offs, in_lens, out_offs, out_caps):
    it = z.make_items(len(in_lens))
    it["in_off"], it["in_len"], it["out_off"], it["out_cap"] = in_offs, in_lens, out_offs, out_caps
    return it


def test_host_alloc_is_page_locked_and_usable(engine):
    import zlibts_b200 as z
    a = z.host_alloc(1 << 20)
    assert a.size == 1 << 20 and z.host_is_pinned(a)
    assert not z.host_is_pinned(np.zeros(1 << 20, dtype=np.uint8))
    a[:] = 7
    it = _items(z, [0], [a.size], [0], [0])
    r = engine.checksum_batch_host(a, it)
    assert int(r["adler32"][0]) == zlib.adler32(a.tobytes())
    z.host_free(a)


@pytest.mark.parametrize("mib", [3, 80])
def test_pageable_and_pinned_buffers_give_the_same_bytes(engine, mib):
    """3 MiB: one wave; 80 MiB: the multi-wave pipelines (input staged by the copy threads ahead of the waves, output
    drained behind them). Deflate then inflate, each through pinned and through pageable buffers."""
    import zlibts_b200 as z
    from zlibts_b200 import synth
    n = mib << 20
    src = synth.mixed(n, 77)
    cap = z.deflate_bound(n)
    it = _items(z, [0], [n], [0], [cap])
    pin_in, pin_out = z.host_alloc(n), z.host_alloc(cap)
    pin_in[:] = src
    pg_in, pg_out = src.copy(), np.full(cap, 0xAB, dtype=np.uint8)
    assert z.host_is_pinned(pin_in) and not z.host_is_pinned(pg_in)
    r1 = engine.deflate_batch_host(pin_in, pin_out, it, flags=z.DEFLATE_WANT_CRC32)
    r2 = engine.deflate_batch_host(pg_in, pg_out, it, flags=z.DEFLATE_WANT_CRC32)
    m = int(r1["out_len"][0])
    assert int(r1["status"][0]) == 0 and int(r2["status"][0]) == 0 and int(r2["out_len"][0]) == m
    assert np.array_equal(pin_out[:m], pg_out[:m]) and int(r1["crc32"][0]) == int(r2["crc32"][0]) == zlib.crc32(src)
    assert (pg_out[m:] == 0xAB).all(), "bytes behind the stream were touched"
    assert zlib.decompress(pg_out[:m].tobytes(), -15) == src.tobytes()
    # inflate: the chunks of the stream as independent items would need their offsets; one big item with SPLIT instead
    it2 = _items(z, [0], [m], [0], [n])
    for h_in, h_out in ((pin_out[:m], z.host_alloc(n)), (pg_out[:m].copy(), np.full(n + 5, 0xCD, dtype=np.uint8))):
        r = engine.inflate_batch_host(h_in, h_out, it2, flags=z.INFLATE_SPLIT | z.INFLATE_WANT_ADLER32)
        assert int(r["status"][0]) == 0 and int(r["out_len"][0]) == n and int(r["adler32"][0]) == zlib.adler32(src)
        assert np.array_equal(h_out[:n], src)
        if h_out.size > n:
            assert (h_out[n:] == 0xCD).all()


def test_batched_inflate_host_waves_copy_back_only_what_was_written(engine):
    """4096 streams (several waves), pageable buffers, slots with slack: every output equals its input, the slack
    behind every item keeps the caller's bytes (the staging buffer of earlier calls must not leak into it)."""
    import zlibts_b200 as z
    rng = np.random.default_rng(3)
    n_items = 4096
    datas = [rand_bytes(rng, int(rng.integers(1, 3000)), 16).tobytes() for _ in range(n_items)]
    streams = [zlib_raw(d, 6) for d in datas]
    in_lens = np.array([len(s) for s in streams], dtype=np.uint64)
    in_offs = np.concatenate([[0], np.cumsum(in_lens)[:-1]]).astype(np.uint64)
    caps = np.array([len(d) + 9 for d in datas], dtype=np.uint64)
    out_offs = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64)
    blob = np.frombuffer(b"".join(streams), dtype=np.uint8).copy()
    it = _items(z, in_offs, in_lens, out_offs, caps)
    # a first call leaves other bytes in the library's staging buffers
    junk = np.full(int(caps.sum()), 0x5A, dtype=np.uint8)
    engine.inflate_batch_host(blob, junk, it)
    h_out = np.full(int(caps.sum()), 0xEE, dtype=np.uint8)
    r = engine.inflate_batch_host(blob, h_out, it)
    assert int(r["status"].max()) == 0
    for k in (0, 1, 17, 1023, 1024, 2047, 2048, 4095):
        o, d = int(out_offs[k]), datas[k]
        assert h_out[o:o + len(d)].tobytes() == d and (h_out[o + len(d):o + len(d) + 9] == 0xEE).all(), k
    flat = np.concatenate([np.frombuffer(d, dtype=np.uint8) for d in datas])
    mask = np.ones(h_out.size, dtype=bool)
    for o, d in zip(out_offs, datas):
        mask[int(o):int(o) + len(d)] = False
    assert (h_out[mask] == 0xEE).all() and np.array_equal(h_out[~mask], flat)


def test_many_deflate_items_copy_back_only_what_was_written(engine):
    import zlibts_b200 as z
    from zlibts_b200 import synth
    n_items = 600   # above the per-item read-back limit of the wave pipeline: the merged-range path
    datas = [synth.text(2000 + 7 * i, 50 + i).tobytes() for i in range(n_items)]
    lens = np.array([len(d) for d in datas], dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    caps = np.array([z.deflate_bound(int(l)) for l in lens], dtype=np.uint64)
    ooffs = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64)
    blob = np.frombuffer(b"".join(datas), dtype=np.uint8).copy()
    it = _items(z, offs, lens, ooffs, caps)
    engine.deflate_batch_host(blob, np.full(int(caps.sum()), 0x11, dtype=np.uint8), it)
    h_out = np.full(int(caps.sum()), 0x77, dtype=np.uint8)
    r = engine.deflate_batch_host(blob, h_out, it)
    assert int(r["status"].max()) == 0
    for k in range(n_items):
        o, m = int(ooffs[k]), int(r["out_len"][k])
        assert (h_out[o + m:o + int(caps[k])] == 0x77).all(), k
        if k % 37 == 0:
            assert zlib.decompress(h_out[o:o + m].tobytes(), -15) == datas[k]


def test_marker_split_scratch_is_bounded(engine):
    """A valid stream of 250 000 empty stored blocks + one with data is full of `00 00 FF FF`: the split must not ask
    for a scratch slot per marker (that was 32 GiB); it declines and the one-warp decoder gives the result."""
    import torch
    import zlibts_b200 as z
    payload = b"payload after a quarter of a million empty blocks"
    s = b"\x00\x00\x00\xff\xff" * 250000 + b"\x01" + len(payload).to_bytes(2, "little") + \
        (len(payload) ^ 0xFFFF).to_bytes(2, "little") + payload
    assert zlib.decompress(s, -15) == payload
    free0, _ = torch.cuda.mem_get_info()
    it = _items(z, [0], [len(s)], [0], [4096])
    d_in = torch.from_numpy(np.frombuffer(s, dtype=np.uint8).copy()).cuda()
    d_out = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    r = engine.inflate_batch(d_in, d_out, it, z.INFLATE_SPLIT)
    assert int(r["status"][0]) == 0 and d_out[:len(payload)].cpu().numpy().tobytes() == payload
    assert int(r["in_used"][0]) == len(s)
    free1, _ = torch.cuda.mem_get_info()
    # 1.25 MB of input: the bound is what such an input could inflate to (2048 x = 2.6 GB), not a slot per marker
    assert free0 - free1 < (3 << 30), "the marker split reserved more scratch than the input could ever need"
