"""GPU: chunk-parallel deflate vs the oracle (reference RawDeflate restatement).

Stage parity: LZ77 tokens + histograms (src/LZ77.ts), code lengths (src/RawDeflate.ts:440-571).
End to end:   compat bytes == oracle RawDeflate(chunk) per chunk; joined multi-chunk streams decode
              with the oracle's RawInflate and with CPython zlib."""
import zlib

import numpy as np
import pytest

import oracle
from helpers import pack, rand_bytes

pytestmark = pytest.mark.gpu

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163,
            195, 227, 258]
DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
             4097, 6145, 8193, 12289, 16385, 24577]


def oracle_tokens(data):
    """oracle LZ77 Uint16 stream -> engine token words (literal byte | 0x80000000 | (len-3)<<16 | (dist-1))."""
    tok, fl, fd = oracle.lz77(data)
    out, i = [], 0
    t = tok.tolist()
    while i < len(t):
        s = t[i]
        if s < 256:
            out.append(s)
            i += 1
        elif s == 256:
            break
        else:
            ln = LEN_BASE[s - 257] + t[i + 1]
            ds = DIST_BASE[t[i + 3]] + t[i + 4]
            out.append(0x80000000 | ((ln - 3) << 16) | (ds - 1))
            i += 6
    return np.array(out, dtype=np.uint32), np.concatenate([fl, fd])


def sample_inputs():
    from zlibts_b200 import synth
    rng = np.random.default_rng(11)
    ins = [b"a", b"ab", b"abc", b"abcd", b"aaaaaaaaaa", b"abcabcabcabc", b"hello hello hello hello", bytes(range(256))]
    ins += [rand_bytes(rng, n, a).tobytes() for n, a in
            [(5, 2), (300, 2), (5000, 2), (65536, 2), (65536, 3), (40000, 4), (65536, 16), (65536, 256), (2049, 5),
             (2048, 7), (4097, 3), (65535, 64)]]
    ins += [b"\0" * 65536, b"xy" * 30000, (b"0123456789abcdefghijklmnopqrstuvwxyz" * 2000)[:65536]]
    ins += [synth.text(65536, 1).tobytes(), synth.mixed(65536, 2).tobytes(), synth.mixed(65536, 3, 512).tobytes(),
            synth.text(30000, 5).tobytes(), synth.mixed(50001, 6).tobytes()]
    # a 1000-byte block repeated: many equally long candidates (ties -> nearest)
    blk = rand_bytes(rng, 1000, 256).tobytes()
    ins.append((blk * 66)[:65536])
    return ins


def test_lz77_tokens_and_histograms(engine):
    import torch
    for data in sample_inputs():
        d = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
        tok, hist = engine.debug_lz77(d, len(data))
        want_tok, want_hist = oracle_tokens(data)
        assert len(tok) == len(want_tok), (len(data), len(tok), len(want_tok))
        bad = np.nonzero(tok != want_tok)[0]
        assert bad.size == 0, (len(data), int(bad[0]), hex(int(tok[bad[0]])), hex(int(want_tok[bad[0]])))
        assert np.array_equal(hist, want_hist), len(data)


def test_code_lengths_match_reference_heap_and_package_merge(engine):
    rng = np.random.default_rng(12)
    cases = []
    for _ in range(60):
        nsym, limit = [(286, 15), (30, 7), (19, 7)][int(rng.integers(0, 3))]
        kind = int(rng.integers(0, 5))
        if kind == 0:
            f = rng.integers(0, 4, nsym)
        elif kind == 1:
            f = rng.integers(0, 60000, nsym)
        elif kind == 2:
            f = (rng.geometric(0.02, nsym) * (rng.random(nsym) < 0.5)).astype(np.int64)
        elif kind == 3:
            f = np.zeros(nsym, dtype=np.int64)
            f[rng.integers(0, nsym, int(rng.integers(1, 4)))] = rng.integers(1, 100)
        else:
            f = np.floor(2.0 ** (np.arange(nsym) % 24) * rng.random()).astype(np.int64) % 65536  # skew: limit binds
        cases.append((f.astype(np.uint32), limit))
    for f, limit in cases:
        got = engine.debug_code_lengths(f, limit)
        want = oracle.get_lengths(f, limit)
        assert np.array_equal(got, want), (f.tolist(), limit, got.tolist(), want.tolist())


def _deflate_items(engine, datas, block_type=2, chunk_bytes=0, flags=0, align=1, host=False):
    import torch
    import zlibts_b200 as z
    blob, offs, lens = pack(datas, align)
    caps = [z.deflate_bound(n, chunk_bytes, block_type) for n in lens]
    ooffs = np.concatenate([[0], np.cumsum(caps)]).astype(np.uint64)
    items = z.make_items(len(datas))
    items["in_off"], items["in_len"], items["out_off"], items["out_cap"] = offs, lens, ooffs[:-1], caps
    if host:
        h_out = np.zeros(int(ooffs[-1]), dtype=np.uint8)
        res = engine.deflate_batch_host(blob, h_out, items, block_type, chunk_bytes, flags)
    else:
        d_out = torch.zeros(int(ooffs[-1]), dtype=torch.uint8, device="cuda")
        res = engine.deflate_batch(torch.from_numpy(blob).cuda(), d_out, items, block_type, chunk_bytes, flags)
        h_out = d_out.cpu().numpy()
    outs = [h_out[int(o):int(o) + int(r["out_len"])].tobytes() for o, r in zip(ooffs[:-1], res)]
    return outs, res


def test_single_chunk_bytes_identical_to_reference(engine):
    datas = sample_inputs()
    for btype in (oracle.DYNAMIC, oracle.FIXED, oracle.NONE):
        outs, res = _deflate_items(engine, datas, btype, align=1)
        for d, o, r in zip(datas, outs, res):
            assert int(r["status"]) == 0
            want = oracle.raw_deflate(d, btype)
            assert o == want, (btype, len(d), len(o), len(want))
            assert zlib.decompress(o, -15) == d


def test_multi_chunk_join_roundtrip_and_per_chunk_identity(engine):
    from zlibts_b200 import synth
    rng = np.random.default_rng(13)
    datas = [synth.mixed(300000, 21).tobytes(), synth.text(200001, 22).tobytes(), rand_bytes(rng, 70000, 3).tobytes(),
             b"z" * 131072, rand_bytes(rng, 131073, 256).tobytes()]
    for chunk in (0, 4096, 1000):
        outs, res = _deflate_items(engine, datas, oracle.DYNAMIC, chunk)
        cb = chunk or 65536
        for d, o, r in zip(datas, outs, res):
            assert int(r["status"]) == 0 and int(r["blocks"]) == -(-len(d) // cb)
            assert zlib.decompress(o, -15) == d                           # RFC-valid single stream
            ref, ip = oracle.raw_inflate(o + b"\0\0\0\0", 0, out_cap=len(d))  # reference decoder accepts the joins
            assert ref == d and ip == len(o)
            # compat: chunk k's bytes == reference RawDeflate(chunk k) with BFINAL masked
            pos = 0
            nchunks = -(-len(d) // cb)
            for k in range(nchunks):
                want = bytearray(oracle.raw_deflate(d[k * cb:(k + 1) * cb]))
                if k + 1 < nchunks:
                    want[0] &= 0xFE
                got = o[pos:pos + len(want)]
                assert got == bytes(want), (chunk, k)
                pos += len(want)
                if k + 1 < nchunks:  # join marker: [00] 00 00 FF FF
                    if o[pos:pos + 4] == b"\x00\x00\xff\xff":
                        pos += 4
                    else:
                        assert o[pos:pos + 5] == b"\x00\x00\x00\xff\xff", (chunk, k)
                        pos += 5
            assert pos == len(o)


def test_batch_of_entries_host_path_and_checksums(engine):
    """zip-like batch: many small independent items through the host-buffer entry point."""
    import zlibts_b200 as z
    rng = np.random.default_rng(14)
    datas = [rand_bytes(rng, int(n), int(a)).tobytes()
             for n, a in zip(rng.integers(1, 9000, 300), rng.choice([2, 4, 26, 256], 300))]
    datas += [b"", b"x"]
    outs, res = _deflate_items(engine, datas, oracle.DYNAMIC, flags=z.DEFLATE_WANT_CRC32 | z.DEFLATE_WANT_ADLER32,
                               host=True)
    for d, o, r in zip(datas, outs, res):
        assert int(r["status"]) == 0
        assert zlib.decompress(o, -15) == d
        if d:
            assert o == oracle.raw_deflate(d)
        assert int(r["crc32"]) == zlib.crc32(d) and int(r["adler32"]) == zlib.adler32(d)


def test_output_overflow_status(engine):
    import torch
    import zlibts_b200 as z
    rng = np.random.default_rng(15)
    data = rand_bytes(rng, 100000, 256)
    items = z.make_items(1)
    items["in_len"], items["out_cap"] = data.size, 50000
    d_out = torch.zeros(60000, dtype=torch.uint8, device="cuda")
    res = engine.deflate_batch(torch.from_numpy(data).cuda(), d_out, items)
    assert int(res["status"][0]) == z.ST_OUT_OVERFLOW
    assert int(d_out[50000:].max()) == 0  # nothing written past the slot


def test_gpu_roundtrip_large_property(engine):
    """C2-shaped property test at 64 MiB: GPU deflate -> GPU inflate reproduces the input bit-exactly."""
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import synth
    n = 64 << 20
    data = synth.mixed(n, 2)
    d_in = torch.from_numpy(data).cuda()
    cap = z.deflate_bound(n)
    d_z = torch.empty(cap, dtype=torch.uint8, device="cuda")
    items = z.make_items(1)
    items["in_len"], items["out_cap"] = n, cap
    r = engine.deflate_batch(d_in, d_z, items, flags=z.DEFLATE_WANT_CRC32)
    assert int(r["status"][0]) == 0 and int(r["blocks"][0]) == n // 65536
    clen = int(r["out_len"][0])
    d_o = torch.empty(n, dtype=torch.uint8, device="cuda")
    it2 = z.make_items(1)
    it2["in_len"], it2["out_cap"] = clen, n
    r2 = engine.inflate_batch(d_z, d_o, it2, flags=z.INFLATE_WANT_CRC32)
    assert int(r2["status"][0]) == 0 and int(r2["out_len"][0]) == n and int(r2["in_used"][0]) == clen
    assert torch.equal(d_o, d_in)
    assert int(r2["crc32"][0]) == int(r["crc32"][0]) == zlib.crc32(data)
    # first chunk equals the reference's bytes for that chunk (BFINAL masked)
    want = bytearray(oracle.raw_deflate(data[:65536]))
    want[0] &= 0xFE
    assert bytes(d_z[:len(want)].cpu().numpy()) == bytes(want)


def test_not_final_shards_concatenate_into_one_stream(engine):
    """Multi-GPU layout: every rank but the last deflates its shard with NOT_FINAL; the shards'
    outputs, laid end to end at the scanned offsets, are ONE stream for zlib and for the reference's decoder."""
    import zlibts_b200 as z
    from zlibts_b200 import shard, synth
    data = synth.mixed(9 * 65536 + 777, 41).tobytes()
    ranges = shard.chunk_ranges(len(data), 65536, 3)
    pieces, parts = [], []
    for r, (lo, hi) in enumerate(ranges):
        fl = z.DEFLATE_WANT_CRC32 | z.DEFLATE_WANT_ADLER32 | (0 if r == 2 else z.DEFLATE_NOT_FINAL)
        outs, res = _deflate_items(engine, [data[lo:hi]], oracle.DYNAMIC, flags=fl)
        pieces.append(outs[0])
        parts.append((int(res["crc32"][0]), int(res["adler32"][0]), hi - lo))
    whole = b"".join(pieces)
    assert zlib.decompress(whole, -15) == data
    ref, ip = oracle.raw_inflate(whole + b"\0\0\0\0", 0, out_cap=len(data))
    assert ref == data and ip == len(whole)
    assert shard.combine_checksums(parts, z.crc32_combine, z.adler32_combine) == (zlib.crc32(data), zlib.adler32(data), len(data))


def test_fast_mode_valid_streams_and_ratio_within_tolerance(engine):
    """ZLB_MODE_FAST (bounded search depth, optionally with lazy evaluation): a valid stream for zlib and for the
    reference's decoder; size within 3 % of the reference-compatible mode on the benchmark generators (north_star
    tolerance)."""
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import synth
    rng = np.random.default_rng(53)
    cases = [("mixed", synth.mixed(40 * 65536 + 99, 51).tobytes(), 1.03), ("text", synth.text(24 * 65536, 52).tobytes(), 1.03),
             ("random", rand_bytes(rng, 5 * 65536 + 7, 256).tobytes(), 1.03),
             ("periodic", (b"0123456789abcdefghijklmnopqrstuvwxyz" * 40000)[:20 * 65536], None),
             ("zeros", b"\0" * 300000, None), ("tiny", b"abcabcabcabc", None), ("one", b"x", None)]
    for name, data, tol in cases:
        n = len(data)
        d_in = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
        cap = z.deflate_bound(n)
        it = z.make_items(1)
        it["in_len"], it["out_cap"] = n, cap
        sizes = {}
        for mname, mode in (("compat", z.MODE_COMPAT), ("fast", z.MODE_FAST), ("fast8", z.mode_fast(8)), ("fast64", z.mode_fast(64)),
                            ("lazy", z.MODE_FAST | z.MODE_LAZY), ("lazy64", z.mode_fast(64) | z.MODE_LAZY)):
            d_z = torch.zeros(cap, dtype=torch.uint8, device="cuda")
            r = engine.deflate_batch(d_in, d_z, it, mode=mode)
            assert int(r["status"][0]) == 0
            out = d_z[:int(r["out_len"][0])].cpu().numpy().tobytes()
            assert zlib.decompress(out, -15) == data, (name, mname)
            ref, ip = oracle.raw_inflate(out + b"\0\0\0\0", 0, out_cap=n)
            assert ref == data and ip == len(out), (name, mname)
            sizes[mname] = len(out)
        if tol:
            assert sizes["fast"] <= sizes["compat"] * tol, (name, sizes)
            assert sizes["fast64"] <= sizes["compat"] * tol, (name, sizes)
            assert sizes["lazy"] <= sizes["fast"], (name, sizes)  # lazy evaluation never loses to the greedy parse here
        if name == "text":  # ... and beats the reference's (exhaustive, greedy) parse where matches are short
            assert sizes["lazy64"] < sizes["compat"], sizes
        else:
            assert sizes["fast"] <= max(3 * sizes["compat"], sizes["compat"] + n // 50), (name, sizes)
    # a batch of small items through the fast kernel (ragged sizes around the tile size)
    datas = [rand_bytes(rng, int(k), 4).tobytes() for k in list(range(1, 140)) + [4095, 4096, 4097, 65535, 65536]]
    blob, offs, lens = pack(datas)
    caps = [z.deflate_bound(k) for k in lens]
    ooffs = np.concatenate([[0], np.cumsum(caps)]).astype(np.uint64)
    items = z.make_items(len(datas))
    items["in_off"], items["in_len"], items["out_off"], items["out_cap"] = offs, lens, ooffs[:-1], caps
    d_out = torch.zeros(int(ooffs[-1]), dtype=torch.uint8, device="cuda")
    for mode in (z.MODE_FAST, z.MODE_FAST | z.MODE_LAZY, z.mode_fast(3) | z.MODE_LAZY | z.MODE_PRIMED):
        d_out.zero_()
        res = engine.deflate_batch(torch.from_numpy(blob).cuda(), d_out, items, mode=mode)
        h = d_out.cpu().numpy()
        for d, o, r in zip(datas, ooffs[:-1], res):
            assert int(r["status"]) == 0
            assert zlib.decompress(h[int(o):int(o) + int(r["out_len"])].tobytes(), -15) == d, (mode, len(d))


def _fuzz_inputs(rng, count, big_every=50, small_max=6000):
    datas = []
    for i in range(count):
        kind = i % 8
        n = int(rng.integers(1, small_max)) if i % big_every else int(rng.integers(60000, 65537))
        if kind == 0:
            d = rand_bytes(rng, n, int(rng.choice([1, 2, 3, 4, 8]))).tobytes()
        elif kind == 1:
            d = rand_bytes(rng, n, 256).tobytes()
        elif kind == 2:  # repeated blocks with mutations: many equal-length candidates
            blk = rand_bytes(rng, int(rng.integers(3, 300)), int(rng.choice([2, 16, 256]))).tobytes()
            b = bytearray((blk * (n // len(blk) + 1))[:n])
            for _ in range(int(rng.integers(0, 6))):
                b[int(rng.integers(0, n))] ^= 1
            d = bytes(b)
        elif kind == 3:  # runs of varying length
            d = b"".join(bytes([int(rng.integers(0, 4))]) * int(rng.integers(1, 700)) for _ in range(40))[:n] or b"z"
        elif kind == 4:  # words
            words = [rand_bytes(rng, int(rng.integers(2, 9)), 26).tobytes() for _ in range(int(rng.integers(2, 60)))]
            d = b" ".join(words[int(k)] for k in rng.integers(0, len(words), n // 4 + 1))[:n]
        elif kind == 5:  # skewed literals: long code lengths, limit may bind
            p = 0.5 ** np.arange(1, 30)
            p = np.concatenate([p, np.full(256 - p.size, p[-1] / 400)])
            d = rng.choice(256, size=n, p=p / p.sum()).astype(np.uint8).tobytes()
        elif kind == 6:  # records
            d = b"".join(int(7 * k).to_bytes(4, "little") + bytes([int(rng.integers(0, 16)), 0, 0, 0]) for k in range(n // 8 + 1))[:n]
        else:  # tails: lengths around the 3-byte search cut-off and 258
            d = (b"ab" * 200)[:int(rng.integers(1, 8))] if i % 3 else b"q" * int(rng.integers(255, 265))
        datas.append(d)
    return datas


def _walk_joined_blocks(o, blocks, stored):
    """`o` must be the blocks in order, each followed (except the last) by the byte-aligning empty stored block:
    00 00 FF FF when the block left at least 3 free bits in its last byte, otherwise 00 00 00 FF FF -- always the
    latter behind a block that is stored itself."""
    pos = 0
    for k, want in enumerate(blocks):
        assert o[pos:pos + len(want)] == want, ("block", k)
        pos += len(want)
        if k + 1 < len(blocks):
            if not stored[k] and o[pos:pos + 4] == b"\x00\x00\xff\xff":
                pos += 4
            else:
                assert o[pos:pos + 5] == b"\x00\x00\x00\xff\xff", ("join", k)
                pos += 5
    assert pos == len(o)


def test_differential_fuzz_against_oracle(engine):
    """2000 structured random inputs in one batch: every output must equal the oracle's RawDeflate bytes (this is where
    LZ77 tie-breaks, heap order and run-length-coding corner cases show up), for DYNAMIC and FIXED."""
    datas = _fuzz_inputs(np.random.default_rng(2026), 2000)
    for btype in (oracle.DYNAMIC, oracle.FIXED):
        outs, res = _deflate_items(engine, datas, btype)
        assert int(res["status"].max()) == 0
        bad = [i for i, (d, o) in enumerate(zip(datas, outs)) if o != oracle.raw_deflate(d, btype)]
        assert not bad, (btype, bad[:5], [len(datas[i]) for i in bad[:5]])


def test_primed_mode_blocks_equal_the_oracle_with_history(engine):
    """every chunk also searches the 32 KiB in front of it. Block k must equal the oracle's block
    construction run with that history (zo_raw_deflate_dict), the blocks are joined as in the compat mode, and the
    item is one stream for zlib and for the reference's decoder."""
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import synth
    rng = np.random.default_rng(23)
    datas = [synth.mixed(300000, 21).tobytes(), synth.text(200001, 22).tobytes(), rand_bytes(rng, 70000, 3).tobytes(),
             b"z" * 131072, rand_bytes(rng, 98305, 256).tobytes(), b"", b"q", synth.text(32768, 5).tobytes(),
             synth.text(32769, 6).tobytes()]
    for chunk in (0, 4096, 1000, 20000):
        cb = chunk or 32768
        blob, offs, lens = pack(datas)
        caps = [z.deflate_bound(len(d), chunk, oracle.DYNAMIC, z.MODE_PRIMED) for d in datas]
        ooffs = np.concatenate([[0], np.cumsum(caps)]).astype(np.uint64)
        items = z.make_items(len(datas))
        items["in_off"], items["in_len"], items["out_off"], items["out_cap"] = offs, lens, ooffs[:-1], caps
        d_out = torch.zeros(int(ooffs[-1]), dtype=torch.uint8, device="cuda")
        res = engine.deflate_batch(torch.from_numpy(blob).cuda(), d_out, items, oracle.DYNAMIC, chunk, 0, z.MODE_PRIMED)
        h = d_out.cpu().numpy()
        for d, o0, r in zip(datas, ooffs[:-1], res):
            o = h[int(o0):int(o0) + int(r["out_len"])].tobytes()
            assert int(r["status"]) == 0
            assert zlib.decompress(o, -15) == d
            ref, ip = oracle.raw_inflate(o + b"\0\0\0\0", 0, out_cap=len(d) + 8)
            assert ref == d and ip == len(o)
            if not d:
                continue
            blocks = oracle.primed_blocks(d, cb)
            assert int(r["blocks"]) == len(blocks)
            pos = 0
            for k, want in enumerate(blocks):
                assert o[pos:pos + len(want)] == want, (chunk, len(d), k)
                pos += len(want)
                if k + 1 < len(blocks):  # join marker: [00] 00 00 FF FF
                    if o[pos:pos + 4] == b"\x00\x00\xff\xff":
                        pos += 4
                    else:
                        assert o[pos:pos + 5] == b"\x00\x00\x00\xff\xff", (chunk, k)
                        pos += 5
            assert pos == len(o)


def test_primed_mode_ratio_fast_variant_and_split_inflate(engine):
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import synth
    d = synth.text(4 << 20, 31)
    d_in = torch.from_numpy(d).cuda()
    sizes = {}
    for name, mode in (("compat", z.MODE_COMPAT), ("primed", z.MODE_PRIMED), ("fast", z.mode_fast()),
                       ("fast-primed", z.mode_fast() | z.MODE_PRIMED)):
        cap = z.deflate_bound(d.size, 0, oracle.DYNAMIC, mode)
        items = z.make_items(1)
        items["in_len"], items["out_cap"] = d.size, cap
        d_z = torch.zeros(cap, dtype=torch.uint8, device="cuda")
        r = engine.deflate_batch(d_in, d_z, items, mode=mode)
        assert int(r["status"][0]) == 0
        n = int(r["out_len"][0])
        sizes[name] = n
        comp = d_z[:n].cpu().numpy().tobytes()
        assert zlib.decompress(comp, -15) == d.tobytes()
        # our own inflate, with the chunk-parallel route requested: primed streams must take the serial route
        it2 = z.make_items(1)
        it2["in_len"], it2["out_cap"] = n, d.size
        d_o = torch.zeros(d.size, dtype=torch.uint8, device="cuda")
        r2 = engine.inflate_batch(d_z, d_o, it2, z.INFLATE_SPLIT | z.INFLATE_WANT_CRC32)
        assert int(r2["status"][0]) == 0 and int(r2["out_len"][0]) == d.size and int(r2["in_used"][0]) == n
        assert torch.equal(d_o, d_in) and int(r2["crc32"][0]) == zlib.crc32(d.tobytes())
    assert sizes["primed"] < 0.97 * sizes["compat"]          # text: history across chunk boundaries pays
    assert sizes["fast-primed"] < sizes["fast"]


def test_smallest_mode_picks_the_shortest_block_per_chunk(engine):
    """per chunk the shortest of dynamic / fixed / stored, each equal to the oracle's construction."""
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import synth
    rng = np.random.default_rng(24)
    # chunks of every kind: incompressible (stored wins), tiny (fixed wins), ordinary (dynamic wins)
    big = rand_bytes(rng, 65536, 256).tobytes() + synth.text(65536, 40).tobytes() + rand_bytes(rng, 65536, 256).tobytes() \
        + b"ab" * 20 + rand_bytes(rng, 30000, 256).tobytes()
    datas = [big, b"hello hello", rand_bytes(rng, 65535, 256).tobytes(), rand_bytes(rng, 65536, 256).tobytes(),
             synth.mixed(200000, 41).tobytes(), b"", b"x", rand_bytes(rng, 100, 256).tobytes()]
    for mode, cb in ((z.MODE_SMALLEST, 65536), (z.MODE_SMALLEST | z.MODE_PRIMED, 32768)):
        blob, offs, lens = pack(datas)
        caps = [z.deflate_bound(len(d), 0, oracle.DYNAMIC, mode) for d in datas]
        ooffs = np.concatenate([[0], np.cumsum(caps)]).astype(np.uint64)
        items = z.make_items(len(datas))
        items["in_off"], items["in_len"], items["out_off"], items["out_cap"] = offs, lens, ooffs[:-1], caps
        d_out = torch.zeros(int(ooffs[-1]), dtype=torch.uint8, device="cuda")
        res = engine.deflate_batch(torch.from_numpy(blob).cuda(), d_out, items, oracle.DYNAMIC, 0, 0, mode)
        h = d_out.cpu().numpy()
        kinds = set()
        for d, o0, r in zip(datas, ooffs[:-1], res):
            o = h[int(o0):int(o0) + int(r["out_len"])].tobytes()
            assert int(r["status"]) == 0 and zlib.decompress(o, -15) == d
            ref, ip = oracle.raw_inflate(o + b"\0\0\0\0", 0, out_cap=len(d) + 8)
            assert ref == d and ip == len(o)
            if not d:   # fixed wins; upstream drops the end-of-block symbol of an empty fixed block, the engine keeps it
                assert o == b"\x03\x00"
                continue
            n_chunks = max(1, -(-len(d) // cb))
            pos = 0
            for k in range(n_chunks):
                lo, hi = k * cb, min(len(d), (k + 1) * cb)
                dl = min(lo, 32768) if mode & z.MODE_PRIMED else 0
                want, kind = oracle.smallest_block(d[lo - dl:hi], dl, k + 1 == n_chunks)
                kinds.add(kind)
                assert o[pos:pos + len(want)] == want, (mode, len(d), k, kind)
                pos += len(want)
                if k + 1 < n_chunks:
                    if kind == "stored":
                        assert o[pos:pos + 5] == b"\x00\x00\x00\xff\xff"
                        pos += 5
                    elif o[pos:pos + 4] == b"\x00\x00\xff\xff":
                        pos += 4
                    else:
                        assert o[pos:pos + 5] == b"\x00\x00\x00\xff\xff", (mode, k)
                        pos += 5
            assert pos == len(o)
            assert len(o) <= len(d) + 5 * (len(d) // 65535 + 1) + 5 * n_chunks + 16   # never much larger than the input
        assert kinds == {"dynamic", "fixed", "stored"}
    # our own inflate reads it, chunk-parallel route requested (stored bytes may mimic a marker: must still be right)
    d = np.frombuffer(big, dtype=np.uint8)
    cap = z.deflate_bound(d.size)
    items = z.make_items(1)
    items["in_len"], items["out_cap"] = d.size, cap
    d_in = torch.from_numpy(d.copy()).cuda()
    d_z = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    r = engine.deflate_batch(d_in, d_z, items, mode=z.MODE_SMALLEST)
    it2 = z.make_items(1)
    it2["in_len"], it2["out_cap"] = int(r["out_len"][0]), d.size
    d_o = torch.zeros(d.size, dtype=torch.uint8, device="cuda")
    r2 = engine.inflate_batch(d_z, d_o, it2, z.INFLATE_SPLIT)
    assert int(r2["status"][0]) == 0 and torch.equal(d_o, d_in)


def test_host_path_wave_plans_equal_the_device_path(engine):
    """The host entry point cuts large jobs into waves (equal waves for medium jobs; ramp-up, cruise, small last wave
    for large ones) whose copies overlap the kernels. Whatever the plan, the bytes are those of the one-wave device
    path, also for items that straddle wave boundaries."""
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import synth
    for sizes in ([48 << 20], [(50 << 20) + 12345, (61 << 20) + 1, 50 << 20, 777]):
        total = sum(sizes)
        data = synth.mixed(total, 91)
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
        caps = [z.deflate_bound(n) for n in sizes]
        ooffs = np.concatenate([[0], np.cumsum(caps)]).astype(np.uint64)
        items = z.make_items(len(sizes))
        items["in_off"], items["in_len"], items["out_off"], items["out_cap"] = offs, sizes, ooffs[:-1], caps
        flags = z.DEFLATE_WANT_CRC32
        d_out = torch.zeros(int(ooffs[-1]), dtype=torch.uint8, device="cuda")
        rd = engine.deflate_batch(torch.from_numpy(data).cuda(), d_out, items, flags=flags)
        h_out = np.zeros(int(ooffs[-1]), dtype=np.uint8)
        rh = engine.deflate_batch_host(data, h_out, items, flags=flags)
        assert int(rd["status"].max()) == 0 and int(rh["status"].max()) == 0
        assert np.array_equal(rd["out_len"], rh["out_len"]) and np.array_equal(rd["crc32"], rh["crc32"])
        dev = d_out.cpu().numpy()
        for o, n in zip(ooffs[:-1], rd["out_len"]):
            assert np.array_equal(dev[int(o):int(o) + int(n)], h_out[int(o):int(o) + int(n)])
        # spot check: the first item decodes
        n0 = int(rh["out_len"][0])
        assert zlib.decompress(h_out[:n0].tobytes(), -15) == data[:sizes[0]].tobytes()


def test_host_path_items_out_of_order_with_pinned_input(engine):
    """Items that are NOT laid out in ascending order switch the host path from wave-by-wave input copies to one
    whole-input copy; the second compute stream must wait for that copy too (it used to wait only for the tables:
    with a pinned input its first wave could start before the bytes had arrived). Several waves, two streams."""
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import synth
    sizes = [(24 << 20) + 333, 31 << 20, (17 << 20) + 5]   # > 4 * sm_count chunks in total: several waves
    total = sum(sizes)
    data_t = torch.from_numpy(synth.mixed(total, 92)).pin_memory()
    data = data_t.numpy()
    # reversed placement: item 0 reads the last bytes of the input, item 2 the first
    ends = np.cumsum(sizes[::-1])[::-1]
    offs = np.array([total - int(e) for e in ends], dtype=np.uint64)
    offs = np.array([int(sum(sizes[i + 1:])) for i in range(len(sizes))], dtype=np.uint64)
    caps = [z.deflate_bound(n) for n in sizes]
    ooffs = np.concatenate([[0], np.cumsum(caps)]).astype(np.uint64)
    items = z.make_items(len(sizes))
    items["in_off"], items["in_len"], items["out_off"], items["out_cap"] = offs, sizes, ooffs[:-1], caps
    out_t = torch.zeros(int(ooffs[-1]), dtype=torch.uint8).pin_memory()
    h_out = out_t.numpy()
    for rep in range(3):
        h_out[:] = 0
        rh = engine.deflate_batch_host(data, h_out, items, flags=z.DEFLATE_WANT_CRC32)
        assert int(rh["status"].max()) == 0
        for i, n in enumerate(sizes):
            o, m = int(ooffs[i]), int(rh["out_len"][i])
            want = data[int(offs[i]):int(offs[i]) + n].tobytes()
            assert int(rh["crc32"][i]) == zlib.crc32(want)
            assert zlib.decompress(h_out[o:o + m].tobytes(), -15) == want, (rep, i)


def test_differential_fuzz_primed_and_smallest(engine):
    """Structured random inputs cut into small chunks, so that most blocks have history in front of them: every block
    of the primed mode must equal the oracle's construction with that history, and with SMALLEST the shortest of the
    oracle's three constructions."""
    import torch
    import zlibts_b200 as z
    datas = _fuzz_inputs(np.random.default_rng(2027), 480, big_every=60, small_max=12000)
    for mode, chunk in ((z.MODE_PRIMED, 1500), (z.MODE_PRIMED, 4096), (z.MODE_PRIMED | z.MODE_SMALLEST, 2048),
                        (z.MODE_SMALLEST, 3000)):
        blob, offs, lens = pack(datas)
        caps = [z.deflate_bound(len(d), chunk, oracle.DYNAMIC, mode) for d in datas]
        ooffs = np.concatenate([[0], np.cumsum(caps)]).astype(np.uint64)
        items = z.make_items(len(datas))
        items["in_off"], items["in_len"], items["out_off"], items["out_cap"] = offs, lens, ooffs[:-1], caps
        d_out = torch.zeros(int(ooffs[-1]), dtype=torch.uint8, device="cuda")
        res = engine.deflate_batch(torch.from_numpy(blob).cuda(), d_out, items, oracle.DYNAMIC, chunk, 0, mode)
        h = d_out.cpu().numpy()
        assert int(res["status"].max()) == 0
        for i, (d, o0, r) in enumerate(zip(datas, ooffs[:-1], res)):
            o = h[int(o0):int(o0) + int(r["out_len"])].tobytes()
            n_chunks = max(1, -(-len(d) // chunk))
            blocks, stored = [], []
            for k in range(n_chunks):
                lo, hi = k * chunk, min(len(d), (k + 1) * chunk)
                dl = min(lo, 32768) if mode & z.MODE_PRIMED else 0
                if mode & z.MODE_SMALLEST:
                    blk, kind = oracle.smallest_block(d[lo - dl:hi], dl, k + 1 == n_chunks)
                else:
                    blk, kind = oracle.raw_deflate_dict(d[lo - dl:hi], dl, k + 1 == n_chunks), "dynamic"
                blocks.append(blk)
                stored.append(kind == "stored")
            try:
                _walk_joined_blocks(o, blocks, stored)
            except AssertionError as e:
                raise AssertionError((mode, chunk, i, len(d), e.args))
            if i % 16 == 0:
                assert zlib.decompress(o, -15) == d


def test_repeated_runs_write_identical_bytes(engine):
    """The LZ77 kernel's threads read each other's visited bits while they are being written and splice afterwards; a
    race would show up as a run that differs. 64 MiB of the C2 data, 12 runs per mode, byte-identical every time, and
    the first compat run decodes to the input (tools/stress_determinism.py is the long version)."""
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import synth
    n = 64 << 20
    data = synth.mixed(n, 2)
    d_in = torch.from_numpy(data).cuda()
    for name, mode in (("compat", z.MODE_COMPAT), ("primed", z.MODE_PRIMED), ("fast", z.MODE_FAST), ("lazy", z.MODE_FAST | z.MODE_LAZY)):
        cap = z.deflate_bound(n, 0, z.DYNAMIC, mode)
        d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        items = z.make_items(1)
        items["in_len"], items["out_cap"] = n, cap
        first = None
        for k in range(12):
            d_out.zero_()
            r = engine.deflate_batch(d_in, d_out, items, mode=mode)
            m = int(r["out_len"][0])
            assert int(r["status"][0]) == 0
            if first is None:
                first = d_out[:m].clone()
            assert m == first.numel() and torch.equal(d_out[:m], first), (name, k)
        if name == "compat":
            assert zlib.decompress(first.cpu().numpy().tobytes(), -15) == data.tobytes()
