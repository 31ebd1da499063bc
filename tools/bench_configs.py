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
    # the first call of each kind pays for scratch and page-locked staging allocations: one untimed call, then the
    # median of three
    c = z.Deflate(d).compress()
    out = z.Inflate(c, {"verify": True}).decompress()
    assert out.tobytes() == d and zlib.decompress(c.tobytes()) == d
    td, ti = [], []
    for _ in range(3):
        t0 = time.perf_counter()
        c = z.Deflate(d).compress()
        t1 = time.perf_counter()
        out = z.Inflate(c, {"verify": True}).decompress()
        t2 = time.perf_counter()
        td.append((t1 - t0) * 1e3)
        ti.append((t2 - t1) * 1e3)
    assert out.tobytes() == d
    return {"config": "C1", "bytes": len(d), "compressed": int(c.size), "ratio": c.size / len(d),
            "deflate_ms_host_api": sorted(td)[1], "inflate_ms_host_api": sorted(ti)[1],
            "timing": "median of 3 calls after one untimed call", "roundtrip_ok": True}


def c3(eng, n_streams=65536, host_leg=True):
    """65 536 independent 64 KiB zlib streams PRODUCED BY THE REFERENCE's algorithm (oracle/, all host threads):
    stream i = 78 9C | RawDeflate(text(65536, 1000 + i) for even i, mixed(65536, 1000 + i) for odd i) | Adler-32.
    Batched inflate, device resident (4 GiB of output), then end to end through zlb_inflate_batch_host."""
    import oracle
    oracle.build()
    stream = torch.cuda.Stream()
    eng2 = z.Engine(0, stream.cuda_stream)
    CH = 65536
    threads = os.cpu_count() or 1
    group = 4096  # generate + compress in groups, keep only the packed zlib streams
    packed, lens = [], []
    gpu_equal = []
    t_ref = 0.0
    for g0 in range(0, n_streams, group):
        g = min(group, n_streams - g0)
        h = np.empty(g * CH, dtype=np.uint8)
        for k in range(g):
            i = g0 + k
            if i % 2 == 0:
                synth.text(CH, 1000 + i, out=h[k * CH:(k + 1) * CH])
            else:
                synth.mixed(CH, 1000 + i, 4096, out=h[k * CH:(k + 1) * CH])
        t0 = time.perf_counter()
        slots, slot, ln = oracle.zlib_chunks_keep_mt(h, CH, threads)
        t_ref += time.perf_counter() - t0
        for k in range(g):
            packed.append(slots[k * slot:k * slot + int(ln[k])].copy())
        lens.extend(int(x) for x in ln)
        if g0 % (8 * group) == 0:   # parity sample: the GPU encoder writes the same bytes for 64 of these chunks
            m = 64
            items = z.make_items(m)
            items["in_off"] = np.arange(m, dtype=np.uint64) * CH
            items["in_len"] = CH
            items["out_off"] = np.arange(m, dtype=np.uint64) * slot
            items["out_cap"] = slot
            with torch.cuda.stream(stream):
                d_out = torch.zeros(m * slot, dtype=torch.uint8, device="cuda")
                r = eng2.deflate_batch(torch.from_numpy(h[:m * CH]).cuda(), d_out, items)
                ho = d_out.cpu().numpy()
            for k in range(m):
                gpu_equal.append(bytes(ho[k * slot:k * slot + int(r["out_len"][k])]) == bytes(packed[g0 + k][2:-4]))
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
    out = {"config": "C3", "streams": n_streams, "out_bytes": total_out, "in_bytes": int(lens.sum()),
           "streams_made_by": "oracle RawDeflate (reference algorithm) on %d host threads, %.1f s" % (threads, t_ref),
           "inflate_ms": best, "inflate_output_GBps": total_out / best / 1e6,
           "roofline_frac_hbm": (total_out + int(lens.sum())) / best / 1e6 / 6538.6,
           "gpu_encoder_bytes_equal_reference_streams": all(gpu_equal), "gpu_encoder_samples": len(gpu_equal),
           "adler32_of_all_outputs_match_trailers": True}
    if host_leg:
        try:
            del d_o
            torch.cuda.empty_cache()
            h_c = torch.from_numpy(blob).pin_memory()
            h_o = torch.empty(total_out, dtype=torch.uint8).pin_memory()
            with torch.cuda.stream(stream):
                eng2.inflate_batch_host(h_c, h_o, items)
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record(stream)
                rh = eng2.inflate_batch_host(h_c, h_o, items)
                e1.record(stream)
                e1.synchronize()
            assert int(rh["status"].max()) == 0 and np.array_equal(h_o[:4 * CH].numpy(), first_plain)
            msh = e0.elapsed_time(e1)
            out["e2e"] = {"ms": msh, "inflate_output_GBps": total_out / msh / 1e6, "h2d_bytes": int(blob.size),
                          "d2h_bytes": total_out, "api": "zlb_inflate_batch_host (page-locked buffers)"}
        except Exception as e:  # host memory for 6 GiB of page-locked buffers may not be there
            out["e2e"] = {"error": repr(e)[:200]}
    return out


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


def c5_sharded(eng, stream, rank, world, dist, gib_total=8, gib_cap_per_gpu=2):
    """C5: `gib_total` GiB of 1 MiB shards (mixed(1 MiB, 5000 + s)) gzip-compressed and decompressed, the shards dealt
    out over the ranks in contiguous runs (at most gib_cap_per_gpu GiB per GPU: a single GPU takes that much of the
    8 GiB and says so). Every shard becomes one gzip member (header, raw deflate with the CRC-32 fused in, trailer,
    framed and packed on the device: zlb_archive); the file is the members of all ranks back to back, so the only
    exchange is the exclusive scan of the ranks' byte counts. Decompression: every member marker-split-inflated and
    CRC-checked. Device resident; times are CUDA events, max over the ranks."""
    import gzip
    shards_total = gib_total << 10
    per = shards_total // world
    per = min(per, gib_cap_per_gpu << 10)
    s0 = rank * per
    n = per << 20
    h = np.empty(n, dtype=np.uint8)
    for k in range(per):
        synth.mixed(1 << 20, 5000 + s0 + k, 4096, out=h[k << 20:(k + 1) << 20])
    ent = z.make_entries(per)
    ent["in_off"] = np.arange(per, dtype=np.uint64) << 20
    ent["in_len"], ent["head_len"] = 1 << 20, 10
    hdr = np.frombuffer(b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03", dtype=np.uint8).copy()

    def timed(fn, reps=2):
        best, r = 1e30, None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(stream)
            r = fn()
            e1.record(stream)
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best, r

    tot, ok, ms_c, ms_d, err = 0, False, 0.0, 0.0, None
    try:  # a rank that fails still takes part in the one collective below
        with torch.cuda.stream(stream):
            d_in = torch.from_numpy(h).cuda()
            d_meta = torch.from_numpy(hdr).cuda()
            d_arc = torch.empty(z.archive_bound(z.FRAME_GZIP, ent), dtype=torch.uint8, device="cuda")
            d_o = torch.empty(n, dtype=torch.uint8, device="cuda")
            eng.archive(z.FRAME_GZIP, d_in, d_meta, ent, d_arc)
            if world > 1:
                dist.barrier()
            ms_c, (tot, rm) = timed(lambda: eng.archive(z.FRAME_GZIP, d_in, d_meta, ent, d_arc))
            assert int(rm["status"].max()) == 0
            # members of this rank: [in_used, in_used + out_len) inside d_arc; deflate data sits behind the 10-byte header
            it = z.make_items(per)
            it["in_off"] = rm["in_used"].astype(np.uint64) + 10
            it["in_len"] = rm["out_len"].astype(np.uint64) - 18
            it["out_off"] = np.arange(per, dtype=np.uint64) << 20
            it["out_cap"] = 1 << 20
            eng.inflate_batch(d_arc, d_o, it, z.INFLATE_WANT_CRC32 | z.INFLATE_SPLIT)
            ms_d, r2 = timed(lambda: eng.inflate_batch(d_arc, d_o, it, z.INFLATE_WANT_CRC32 | z.INFLATE_SPLIT))
            same = bool(torch.equal(d_o, d_in))
            first = bytes(d_arc[:int(rm["out_len"][:2].sum())].cpu().numpy())
        ok = (same and int(r2["status"].max()) == 0 and np.array_equal(r2["crc32"], rm["crc32"]) and
              gzip.decompress(first) == h[:2 << 20].tobytes())
    except Exception as e:
        err = repr(e)[:200]
    mine = torch.tensor([int(tot), n, int(ok), int(ms_c * 1e3), int(ms_d * 1e3)], dtype=torch.int64, device="cuda")
    gathered = [mine]
    if world > 1:
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
    g = [[int(v) for v in t.tolist()] for t in gathered]
    tot_c, tot_n = sum(x[0] for x in g), sum(x[1] for x in g)
    ok_all = all(x[2] for x in g)
    ms_c, ms_d = max(x[3] for x in g) / 1e3, max(x[4] for x in g) / 1e3    # max over the ranks
    out = {"config": "C5", "n_gpus": world, "bytes": tot_n, "bytes_per_gpu": n, "members": per * world,
           "note": ("8 GiB over the ranks" if per * world == shards_total else
                    "capped at %d GiB per GPU: %d GiB of the 8 GiB" % (gib_cap_per_gpu, tot_n >> 30)),
           "compressed": tot_c, "ratio": tot_c / max(tot_n, 1), "gzip_ms": ms_c, "gzip_GBps": tot_n / max(ms_c, 1e-9) / 1e6,
           "gunzip_ms": ms_d, "gunzip_GBps": tot_n / max(ms_d, 1e-9) / 1e6,
           "file_offset_of_rank": [sum(x[0] for x in g[:r]) for r in range(world)],
           "roundtrip_and_crc32_ok_on_every_rank": bool(ok_all), "cpython_gzip_reads_first_members": bool(ok_all),
           "exchange": "one all-gather of 5 integers per rank (member bytes -> file offsets, times); no data-path collective"}
    if err:
        out["error"] = err
    return out


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
