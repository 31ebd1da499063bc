#!/usr/bin/env python
"""BASELINE config C5 at full size: 8 GiB as 8192 shards of 1 MiB, shard s = mixed(1 MiB, 5000 + s), gzip-compressed
and decompressed sharded across the GPUs of one box (not the contract bench: bench.py is).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c5_multi.py [gib]

Every rank takes a contiguous range of shards (zlib.ts_b200/shard.py), writes one gzip member per shard with the
members framed and packed on the device (zlb_archive), and the ranks exchange one integer each (exclusive scan of
their archive sizes): rank r's bytes belong at that offset of the one multi-member gzip file. Decompression: every
rank inflates its own members (independent streams) and checks the CRC-32 of each against the member trailer.
Whole-file CRC-32 of the plain data = the per-rank CRCs combined (zlb_crc32_combine). Device resident, CUDA events,
max over ranks. Rank 0 prints one JSON object."""
import json
import os
import struct
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import zlibts_b200 as z
from zlibts_b200 import shard, synth


def main():
    gib = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    SH = 1 << 20
    n_shards = gib << 10
    b = shard.plan_ranges(np.full(n_shards, SH, dtype=np.uint64), world)
    s0, s1 = b[rank], b[rank + 1]
    mine = s1 - s0
    h = np.empty(mine * SH, dtype=np.uint8)
    for k in range(mine):
        synth.mixed(SH, 5000 + s0 + k, 4096, out=h[k * SH:(k + 1) * SH])
    stream = torch.cuda.Stream()
    eng = z.Engine(local, stream.cuda_stream)
    ent = z.make_entries(mine)
    ent["in_off"] = np.arange(mine, dtype=np.uint64) * SH
    ent["in_len"], ent["head_len"] = SH, 10
    hdr = np.frombuffer(b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03", dtype=np.uint8).copy()

    def timed(fn, reps=3):
        best = 1e30
        for _ in range(reps):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(stream)
            out = fn()
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            if dist is not None:
                t = torch.tensor([ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            best = min(best, ms)
        return best, out

    with torch.cuda.stream(stream):
        d_in = torch.from_numpy(h).cuda()
        d_meta = torch.from_numpy(hdr).cuda()
        d_arc = torch.empty(z.archive_bound(z.FRAME_GZIP, ent), dtype=torch.uint8, device="cuda")
        ms_c, (total, res) = timed(lambda: eng.archive(z.FRAME_GZIP, d_in, d_meta, ent, d_arc))
        assert int(res["status"].max()) == 0
        # the one exchange: where this rank's members go in the file
        off, file_bytes = shard.exclusive_scan(total, rank, world, dist, "cuda")
        # decompress: one item per member (the 10 header bytes skipped), CRC-32 of the output requested
        it = z.make_items(mine)
        it["in_off"] = res["in_used"] + 10
        it["in_len"] = res["out_len"] - 10
        it["out_off"] = np.arange(mine, dtype=np.uint64) * SH
        it["out_cap"] = SH
        d_out = torch.empty(mine * SH, dtype=torch.uint8, device="cuda")
        ms_d, r2 = timed(lambda: eng.inflate_batch(d_arc, d_out, it, z.INFLATE_WANT_CRC32 | z.INFLATE_SPLIT))
        same = bool(torch.equal(d_out, d_in))
        tail = d_arc[:total].cpu().numpy()
    assert int(r2["status"].max()) == 0 and same
    # every member's trailer = CRC-32 + ISIZE of its shard, and the decoder's CRC agrees
    ends = (res["in_used"] + res["out_len"]).astype(np.int64)
    for k in (0, mine // 2, mine - 1):
        crc, isize = struct.unpack("<II", tail[int(ends[k]) - 8:int(ends[k])].tobytes())
        assert crc == zlib.crc32(h[k * SH:(k + 1) * SH]) == int(r2["crc32"][k]) == int(res["crc32"][k]) and isize == SH
    assert np.array_equal(res["crc32"], r2["crc32"])
    # whole-file CRC of the plain data from per-rank partials
    crc_local = 0
    for k in range(mine):
        crc_local = z.crc32_combine(crc_local, int(res["crc32"][k]), SH) if k else int(res["crc32"][0])
    parts = shard.gather_parts((crc_local, 1, mine * SH), rank, world, dist, "cuda")
    crc_all, _, n_all = shard.combine_checksums(parts, z.crc32_combine, z.adler32_combine)
    if rank == 0:
        import gzip
        n0 = int(res["out_len"][:4].sum())
        assert gzip.decompress(tail[:n0].tobytes()) == h[:4 * SH].tobytes()   # CPython reads the first members
        print(json.dumps({
            "config": "C5", "gib": gib, "n_gpus": world, "shards": n_shards, "file_bytes": file_bytes,
            "ratio": file_bytes / (gib << 30), "gzip_ms": ms_c, "gzip_GBps": (gib << 30) / ms_c / 1e6,
            "gunzip_ms": ms_d, "gunzip_GBps": (gib << 30) / ms_d / 1e6, "rank0_offset": off,
            "crc32_of_8GiB_from_partials": crc_all, "plain_bytes": n_all, "roundtrip_ok": True,
            "member_trailers_match_cpython_crc32": True, "exchange": "exclusive scan of per-rank archive sizes"}), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
