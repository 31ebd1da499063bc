#!/usr/bin/env python
"""The other BASELINE.json configs at full size on one GPU, with parity checks (not the contract bench: bench.py is).

    python tools/bench_configs.py [c1] [c3] [c4]  -> one JSON object per config on stdout

C1  Zlib.Deflate / Inflate round trip of 1 MiB synthetic text through the host API.
C3  batched inflate of 65 536 independent 64 KiB zlib streams (stream i = reference-compatible encoder output of
    text(65536, 1000+i) for even i, mixed(65536, 1000+i) for odd i; produced by the GPU compat encoder, whose bytes
    are checked against the oracle on a sample), device resident, 4 GiB of output.
C4  Zlib.Zip of 10 000 synthetic files (256 B .. 256 KiB, log-uniform) + Unzip round trip with per-entry CRC-32.
"""
import json
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import zlibts_b200 as z
from zlibts_b200 import synth


def lcg_sizes(n, seed=4):
    x, out = seed, []
    for _ in range(n):
        x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
        k = (x >> 16) % 81
        out.append(int(256 * 2 ** (k / 8)))
    return out


def c1(eng):
    d = synth.text(1 << 20, 1).tobytes()
    z.api.set_engine(eng)
    t0 = time.perf_counter()
    c = z.Deflate(d).compress()
    t1 = time.perf_counter()
    out = z.Inflate(c, {"verify": True}).decompress()
    t2 = time.perf_counter()
    assert out.tobytes() == d and zlib.decompress(c.tobytes()) == d
    return {"config": "C1", "bytes": len(d), "compressed": int(c.size), "ratio": c.size / len(d),
            "deflate_ms_host_api": (t1 - t0) * 1e3, "inflate_ms_host_api": (t2 - t1) * 1e3, "roundtrip_ok": True}


def c3(eng, n_streams=65536):
    import oracle
    stream = torch.cuda.Stream()
    eng2 = z.Engine(0, stream.cuda_stream)
    CH = 65536
    slot = z.deflate_bound(CH) + 8
    group = 4096  # generate + compress in groups, keep only the packed zlib streams
    packed, lens = [], []
    samples = {}
    for g0 in range(0, n_streams, group):
        g = min(group, n_streams - g0)
        h = np.empty(g * CH, dtype=np.uint8)
        for k in range(g):
            i = g0 + k
            if i % 2 == 0:
                synth.text(CH, 1000 + i, out=h[k * CH:(k + 1) * CH])
            else:
                synth.mixed(CH, 1000 + i, 4096, out=h[k * CH:(k + 1) * CH])
        items = z.make_items(g)
        items["in_off"] = np.arange(g, dtype=np.uint64) * CH
        items["in_len"] = CH
        items["out_off"] = np.arange(g, dtype=np.uint64) * slot + 2   # room for the 2-byte zlib header
        items["out_cap"] = slot - 8
        with torch.cuda.stream(stream):
            d_in = torch.from_numpy(h).cuda()
            d_out = torch.zeros(g * slot, dtype=torch.uint8, device="cuda")
            r = eng2.deflate_batch(d_in, d_out, items, flags=z.DEFLATE_WANT_ADLER32)
            ho = d_out.cpu().numpy()
        assert int(r["status"].max()) == 0
        for k in range(g):
            n = int(r["out_len"][k])
            s = ho[k * slot:k * slot + 2 + n + 4]
            s[0], s[1] = 0x78, 0x9C
            s[2 + n:2 + n + 4] = np.frombuffer(int(r["adler32"][k]).to_bytes(4, "big"), dtype=np.uint8)
            packed.append(s.copy())
            lens.append(2 + n + 4)
        for k in (0, 1):   # parity sample: the GPU encoder's bytes are the reference's
            i = g0 + k
            samples[i] = (bytes(packed[g0 + k][2:-4]) == oracle.raw_deflate(h[k * CH:(k + 1) * CH]))
        if g0 == 0:
            first_plain = h[:4 * CH].copy()
    lens = np.array(lens, dtype=np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    blob = np.concatenate(packed)
    del packed
    items = z.make_items(n_streams)
    items["in_off"] = offs + 2                      # the zlib container's `index` (src/Inflate.ts:61-66)
    items["in_len"] = lens - 2
    items["out_off"] = np.arange(n_streams, dtype=np.uint64) * CH
    items["out_cap"] = CH
    with torch.cuda.stream(stream):
        d_c = torch.from_numpy(blob).cuda()
        d_o = torch.empty(n_streams * CH, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            res = eng2.inflate_batch(d_c, d_o, items, z.INFLATE_WANT_ADLER32)
        ms = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(stream)
            res = eng2.inflate_batch(d_c, d_o, items)
            e1.record(stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        res = eng2.inflate_batch(d_c, d_o, items, z.INFLATE_WANT_ADLER32)
        head = d_o[:4 * CH].cpu().numpy()
    assert int(res["status"].max()) == 0 and int(res["out_len"].min()) == CH
    # Adler-32 of every output equals the trailer of its stream; in_used lands on the trailer
    trailers = np.array([int.from_bytes(blob[int(o + l) - 4:int(o + l)].tobytes(), "big") for o, l in zip(offs, lens)], dtype=np.uint64)
    assert np.array_equal(res["adler32"].astype(np.uint64), trailers)
    assert np.array_equal(res["in_used"].astype(np.uint64), lens - 6)
    assert np.array_equal(head, first_plain)
    best = min(ms)
    total_out = n_streams * CH
    return {"config": "C3", "streams": n_streams, "out_bytes": total_out, "in_bytes": int(lens.sum()),
            "inflate_ms": best, "inflate_output_GBps": total_out / best / 1e6,
            "roofline_frac_hbm": (total_out + int(lens.sum())) / best / 1e6 / 6538.6,
            "encoder_bytes_equal_oracle_on_samples": all(samples.values()), "samples": len(samples),
            "adler32_of_all_outputs_match_trailers": True}


def c4(eng, n_files=10000):
    import datetime
    z.api.set_engine(eng)
    sizes = lcg_sizes(n_files)
    files = {}
    for i, n in enumerate(sizes):
        files["f%05d.bin" % i] = (synth.text(n, 4000 + i) if i % 2 == 0 else synth.mixed(n, 4000 + i)).tobytes()
    total = sum(sizes)
    date = datetime.datetime(2026, 10, 18, 0, 0, 0)
    warm = z.Zip()  # untimed first call: the engine's arenas grow to this job's size once
    for name, data in files.items():
        warm.addFile(data, name, {"date": date})
    z.Unzip(warm.compress(), {"verify": True}).decompressAll()
    del warm
    zp = z.Zip()
    t0 = time.perf_counter()
    for name, data in files.items():
        zp.addFile(data, name, {"date": date})
    arc = zp.compress()
    t1 = time.perf_counter()
    uz = z.Unzip(arc, {"verify": True})
    out = uz.decompressAll()
    t2 = time.perf_counter()
    assert list(out) == list(files)
    for k in files:
        assert out[k].tobytes() == files[k]
    import io
    import zipfile
    with zipfile.ZipFile(io.BytesIO(arc.tobytes())) as zf:   # CPython reads the archive, CRCs included
        assert zf.testzip() is None and len(zf.namelist()) == n_files
    # the same archive through the C ABI alone (zlb_archive_host on pinned buffers): what the N-API addon would call
    import struct
    names = [k.encode() for k in files]
    mt = z.api._dos_time(date)
    heads = [b"PK\x03\x04" + struct.pack("<HHH", 20, 0, 8) + mt + struct.pack("<IIIHH", 0, 0, n, len(nm), 0) + nm
             for nm, n in zip(names, sizes)]
    cdirs = [b"PK\x01\x02" + bytes([20, 0]) + h[4:30] + struct.pack("<HHHII", 0, 0, 0, 0, 0) + nm
             for h, nm in zip(heads, names)]
    eocd = b"PK\x05\x06" + struct.pack("<HHHHIIH", 0, 0, n_files, n_files, 0, 0, 0)
    parts = heads + cdirs + [eocd]
    plens = np.array([len(p) for p in parts], dtype=np.uint64)
    poffs = np.concatenate([[0], np.cumsum(plens)[:-1]]).astype(np.uint64)
    ent = z.make_entries(n_files)
    lens = np.array(sizes, dtype=np.uint64)
    ent["in_off"] = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    ent["in_len"], ent["method"] = lens, 8
    ent["head_off"], ent["head_len"] = poffs[:n_files], plens[:n_files]
    ent["cdir_off"], ent["cdir_len"] = poffs[n_files:2 * n_files], plens[n_files:2 * n_files]
    h_in = torch.from_numpy(np.frombuffer(b"".join(files.values()), dtype=np.uint8).copy()).pin_memory()
    h_meta = torch.from_numpy(np.frombuffer(b"".join(parts), dtype=np.uint8).copy()).pin_memory()
    h_out = torch.empty(z.archive_bound(z.FRAME_ZIP, ent, 22), dtype=torch.uint8).pin_memory()
    best = 1e9
    for _ in range(4):
        t3 = time.perf_counter()
        a2, _r = eng.archive_host(z.FRAME_ZIP, h_in, h_meta, ent, (int(poffs[-1]), 22), h_out=h_out.numpy())
        best = min(best, time.perf_counter() - t3)
    assert a2.tobytes() == arc.tobytes()
    return {"config": "C4", "files": n_files, "bytes": total, "archive_bytes": int(arc.size),
            "zip_seconds_host_api": t1 - t0, "unzip_verify_seconds_host_api": t2 - t1,
            "zip_GBps_host_api": total / (t1 - t0) / 1e9, "unzip_GBps_host_api": total / (t2 - t1) / 1e9,
            "zip_seconds_c_abi": best, "zip_GBps_c_abi": total / best / 1e9,
            "c_abi_note": "zlb_archive_host on pinned buffers: H2D, CRC-32 + deflate of all entries, headers / central "
                          "directory / end record written and packed on the device, one D2H; same bytes as the host API",
            "roundtrip_ok": True, "cpython_zipfile_testzip_ok": True}


def c5(eng, gib=1):
    """C5 on one GPU: `gib` GiB as 1 MiB shards mixed(1 MiB, 5000+s); gzip member = header + raw deflate (64 KiB
    chunks, CRC-32 fused into the call) + trailer; gunzip = marker-split inflate + CRC-32 check. Device resident."""
    import struct
    stream = torch.cuda.Stream()
    e = z.Engine(0, stream.cuda_stream)
    n = gib << 30
    h = np.empty(n, dtype=np.uint8)
    for s_ in range(n >> 20):
        synth.mixed(1 << 20, 5000 + s_, 4096, out=h[s_ << 20:(s_ + 1) << 20])
    cap = z.deflate_bound(n)
    items = z.make_items(1)
    items["in_len"], items["out_cap"] = n, cap
    with torch.cuda.stream(stream):
        d_in = torch.from_numpy(h).cuda()
        d_z = torch.empty(cap, dtype=torch.uint8, device="cuda")
        d_o = torch.empty(n, dtype=torch.uint8, device="cuda")
        best_c, best_d = 1e9, 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(stream)
            r = e.deflate_batch(d_in, d_z, items, flags=z.DEFLATE_WANT_CRC32)
            e1.record(stream)
            e1.synchronize()
            best_c = min(best_c, e0.elapsed_time(e1))
        clen = int(r["out_len"][0])
        it2 = z.make_items(1)
        it2["in_len"], it2["out_cap"] = clen, n
        for _ in range(3):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(stream)
            r2 = e.inflate_batch(d_z, d_o, it2, z.INFLATE_SPLIT | z.INFLATE_WANT_CRC32)
            e1.record(stream)
            e1.synchronize()
            best_d = min(best_d, e0.elapsed_time(e1))
        same = bool(torch.equal(d_o, d_in))
        head = bytes(d_z[:1 << 16].cpu().numpy())
    crc = zlib.crc32(h)
    assert int(r["status"][0]) == 0 and int(r2["status"][0]) == 0 and same
    assert int(r["crc32"][0]) == crc == int(r2["crc32"][0]) and int(r2["out_len"][0]) == n and int(r2["in_used"][0]) == clen
    # the member a GZip writer would emit around it is valid for CPython's gzip on a prefix-sized sample
    small = h[:3 << 20]
    its = z.make_items(1)
    its["in_len"], its["out_cap"] = small.size, z.deflate_bound(small.size)
    with torch.cuda.stream(stream):
        d_s = torch.empty(int(its["out_cap"][0]), dtype=torch.uint8, device="cuda")
        rs = e.deflate_batch(d_in[:small.size], d_s, its, flags=z.DEFLATE_WANT_CRC32)
        body = bytes(d_s[:int(rs["out_len"][0])].cpu().numpy())
    import gzip
    member = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" + body + struct.pack("<II", int(rs["crc32"][0]), small.size)
    assert gzip.decompress(member) == small.tobytes()
    # multi-member gzip, one member per 1 MiB shard, framed and packed on the device (zlb_archive)
    n_sh = n >> 20
    ent = z.make_entries(n_sh)
    ent["in_off"] = np.arange(n_sh, dtype=np.uint64) << 20
    ent["in_len"], ent["head_len"] = 1 << 20, 10
    hdr = np.frombuffer(b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03", dtype=np.uint8).copy()
    with torch.cuda.stream(stream):
        d_meta = torch.from_numpy(hdr).cuda()
        d_arc = torch.empty(z.archive_bound(z.FRAME_GZIP, ent), dtype=torch.uint8, device="cuda")
        best_m = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(stream)
            tot, rm = e.archive(z.FRAME_GZIP, d_in, d_meta, ent, d_arc)
            e1.record(stream)
            e1.synchronize()
            best_m = min(best_m, e0.elapsed_time(e1))
        first = bytes(d_arc[:int(rm["out_len"][:8].sum())].cpu().numpy())
    assert gzip.decompress(first) == h[:8 << 20].tobytes() and int(rm["status"].max()) == 0
    return {"config": "C5 (1 GPU)", "bytes": n, "compressed": clen, "ratio": clen / n,
            "gzip_deflate_plus_crc_ms": best_c, "gzip_GBps": n / best_c / 1e6,
            "gzip_multimember_on_device_ms": best_m, "gzip_multimember_GBps": n / best_m / 1e6,
            "gzip_multimember_bytes": tot, "cpython_gzip_reads_first_8_members": True,
            "gunzip_split_inflate_plus_crc_ms": best_d, "gunzip_GBps": n / best_d / 1e6, "roundtrip_ok": True,
            "crc32_matches_cpython": True, "cpython_gzip_reads_member_sample": True}


def main():
    which = [a.lower() for a in sys.argv[1:]] or ["c1", "c3", "c4", "c5"]
    eng = z.Engine(0)
    for name in which:
        arg = None
        if ":" in name:
            name, arg = name.split(":")
        fn = {"c1": c1, "c3": c3, "c4": c4, "c5": c5}[name]
        r = fn(eng) if arg is None else fn(eng, int(arg))
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
