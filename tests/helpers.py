"""Shared test helpers: seeded inputs, batching through the C-ABI, CPython-zlib cross-oracle."""
import zlib

import numpy as np


def rand_bytes(rng, n, alphabet=256):
    return rng.integers(0, alphabet, size=n, dtype=np.uint8)


def zlib_raw(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return co.compress(bytes(data)) + co.flush()


def pack(buffers, align=1):
    """Concatenate buffers; returns (uint8 array, offsets, lengths)."""
    offs, lens, pos = [], [], 0
    for b in buffers:
        pos = (pos + align - 1) // align * align
        offs.append(pos)
        lens.append(len(b))
        pos += len(b)
    out = np.zeros(max(pos, 1), dtype=np.uint8)
    for o, b in zip(offs, buffers):
        out[o:o + len(b)] = np.frombuffer(bytes(b), dtype=np.uint8)
    return out, offs, lens


def gpu_inflate_many(engine, streams, out_sizes, flags=0, trailer=b"", slack=0):
    """Inflate a list of raw deflate streams on the GPU; returns (list of outputs, results)."""
    import torch
    import zlibts_b200 as z
    blob, offs, lens = pack([bytes(s) + trailer for s in streams])
    caps = [n + slack for n in out_sizes]
    ooffs = np.concatenate([[0], np.cumsum(caps)]).astype(np.uint64)
    items = z.make_items(len(streams))
    items["in_off"] = offs
    items["in_len"] = lens
    items["out_off"] = ooffs[:-1]
    items["out_cap"] = caps
    d_in = torch.from_numpy(blob).cuda()
    d_out = torch.zeros(max(int(ooffs[-1]), 1), dtype=torch.uint8, device="cuda")
    res = engine.inflate_batch(d_in, d_out, items, flags)
    h = d_out.cpu().numpy()
    outs = [h[int(o):int(o) + int(r["out_len"])].tobytes() for o, r in zip(ooffs[:-1], res)]
    return outs, res


class _BitWriter:
    """LSB-first bit writer of a raw DEFLATE stream (test streams built by hand)."""

    def __init__(self):
        self.buf = bytearray()
        self.acc = 0
        self.n = 0

    def bits(self, value, count):  # data element, LSB first
        self.acc |= (value & ((1 << count) - 1)) << self.n
        self.n += count
        while self.n >= 8:
            self.buf.append(self.acc & 0xFF)
            self.acc >>= 8
            self.n -= 8

    def code(self, code, length):  # Huffman code, MSB first
        for i in range(length - 1, -1, -1):
            self.bits((code >> i) & 1, 1)

    def align(self):
        if self.n:
            self.bits(0, 8 - self.n)

    def bytes(self):
        self.align()
        return bytes(self.buf)


def _canonical_codes(lengths):
    """RFC 1951 3.2.2: codes of a list of code lengths."""
    max_len = max(lengths)
    bl_count = [0] * (max_len + 1)
    for l in lengths:
        if l:
            bl_count[l] += 1
    code, next_code = 0, [0] * (max_len + 2)
    for b in range(1, max_len + 1):
        code = (code + bl_count[b - 1]) << 1
        next_code[b] = code
    out = []
    for l in lengths:
        if l:
            out.append(next_code[l])
            next_code[l] += 1
        else:
            out.append(0)
    return out


def long_code_match_stream(history, n_matches, seed=5):
    """A raw DEFLATE stream whose second block spends the most bits a symbol can take: stored blocks holding `history`
    (>= 32768 bytes), then one dynamic block of `n_matches` matches in which the length symbol 284 (5 extra bits) and
    the distance symbols 28 / 29 (13 extra bits) own 15-bit codes -- 48 bits per match. The short codes belong to
    symbols the block never uses; both codes are complete. Returns (stream, expected output)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    assert len(history) >= 32768
    w = _BitWriter()
    out = bytearray(history)
    for at in range(0, len(history), 65535):
        piece = history[at:at + 65535]
        w.bits(0, 1)
        w.bits(0, 2)
        w.align()
        w.bits(len(piece), 16)
        w.bits(len(piece) ^ 0xFFFF, 16)
        w.buf.extend(piece)
    ll = [0] * 286
    for i in range(14):
        ll[i] = i + 1          # literals 0..13: lengths 1..14 (never sent)
    ll[256] = 15               # end of block
    ll[284] = 15               # length 227..257, 5 extra bits
    dl = [0] * 30
    for i in range(14):
        dl[i] = i + 1
    dl[28] = 15                # distance 16385..24576, 13 extra bits
    dl[29] = 15                # distance 24577..32768
    cl = [4] * 13 + [5] * 6    # code-length alphabet 0..18: complete (13 / 16 + 6 / 32)
    order = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]
    llc, dlc, clc = _canonical_codes(ll), _canonical_codes(dl), _canonical_codes(cl)
    w.bits(1, 1)
    w.bits(2, 2)
    w.bits(286 - 257, 5)
    w.bits(30 - 1, 5)
    w.bits(19 - 4, 4)
    for sym in order:
        w.bits(cl[sym], 3)
    for l in ll + dl:
        w.code(clc[l], cl[l])
    for _ in range(n_matches):
        length = int(rng.integers(227, 258))
        ds = int(rng.integers(28, 30))
        base = 16385 if ds == 28 else 24577
        dist = min(base + int(rng.integers(0, 8192)), len(out))
        if dist < base:
            ds, base = 28, 16385
            dist = max(dist, base)
        w.code(llc[284], 15)
        w.bits(length - 227, 5)
        w.code(dlc[ds], 15)
        w.bits(dist - base, 13)
        for _k in range(length):
            out.append(out[-dist])
    w.code(llc[256], 15)
    return w.bytes(), bytes(out)
