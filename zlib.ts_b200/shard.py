"""Multi-GPU sharding of the hot path: units (64 KiB chunks, zip entries, streams, gzip
members) are independent, so every rank owns a contiguous range of units, balanced by input bytes, and runs the
single-GPU engine on it. The only cross-rank step is an exclusive scan of the per-rank output byte counts -- a few
integers, exchanged through whatever torch.distributed backend the job already has (gloo in the CPU tests, NCCL
under torchrun on a GPU box). No data-path collective exists. Whole-buffer CRC-32 / Adler-32 / ISIZE come from
combining the per-rank partials (zlb_crc32_combine / zlb_adler32_combine)."""
import numpy as np


def plan_ranges(sizes, world):
    """Contiguous unit ranges per rank, balanced by bytes: returns `world + 1` boundaries into the unit list."""
    sizes = np.asarray(sizes, dtype=np.uint64)
    n = sizes.size
    cum = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        k = int(np.searchsorted(cum, target, side="left"))
        k = min(max(k, bounds[-1]), n)
        bounds.append(k)
    bounds.append(n)
    return bounds


def chunk_ranges(n_bytes, chunk, world):
    """Byte ranges per rank for one buffer cut on chunk boundaries: [(lo, hi)] * world."""
    n_chunks = (n_bytes + chunk - 1) // chunk
    b = plan_ranges(np.full(n_chunks, chunk, dtype=np.uint64), world) if n_chunks else [0] * (world + 1)
    return [(min(b[r] * chunk, n_bytes), min(b[r + 1] * chunk, n_bytes)) for r in range(world)]


def exclusive_scan(value, rank, world, dist=None, device=None):
    """(offset of this rank, total) of one integer per rank -- the single cross-device exchange of the path."""
    if world == 1 or dist is None:
        return 0, int(value)
    import torch
    t = torch.zeros(world, dtype=torch.int64, device=device or "cpu")
    t[rank] = int(value)
    dist.all_reduce(t)
    return int(t[:rank].sum().item()), int(t.sum().item())


def combine_checksums(parts, crc32_combine, adler32_combine):
    """parts = [(crc32, adler32, length)] in rank order -> (crc32, adler32, total length) of the concatenation."""
    crc, adler, total = 0, 1, 0
    for c, a, n in parts:
        crc = crc32_combine(crc, c, n) if total else c
        adler = adler32_combine(adler, a, n) if total else a
        total += n
    return crc, adler, total


def gather_parts(local, rank, world, dist=None, device=None):
    """All ranks' (crc32, adler32, length) triples, in rank order."""
    if world == 1 or dist is None:
        return [tuple(int(x) for x in local)]
    import torch
    t = torch.zeros(world, 3, dtype=torch.int64, device=device or "cpu")
    t[rank] = torch.tensor([int(x) for x in local], dtype=torch.int64)
    dist.all_reduce(t)
    return [tuple(int(x) for x in row) for row in t.cpu().tolist()]
