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
