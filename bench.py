#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 DEFLATE engine (BASELINE.json metric / configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one chunk-parallel raw deflate (DYNAMIC blocks, 64 KiB chunks, reference-compatible bytes)
of one 256 MiB `mixed(268435456, seed)` buffer per GPU. The same JSON line also
carries the batched-inflate leg (the 4096 independent 64 KiB streams of that buffer, C3-shaped) under
"inflate". Multi-GPU: one process per GPU (torchrun), every rank owns its own buffer (weak scaling),
no data-path collective; the only exchange is the exclusive scan of the per-rank output sizes.

`--impl reference` times the reference's CPU algorithm (oracle/: C restatement of RawDeflate, because no
JavaScript engine exists in this image) on all host threads, on bounded samples of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

CHUNK = 65536
WORKLOAD_BYTES = 256 << 20
METRIC = "deflate_input_GBps"
UNIT = "GB/s"
WORKLOAD = ("C2: 256 MiB mixed(seed=2+rank, seg=4096) per GPU, chunk-parallel raw deflate, 64 KiB chunks, "
            "DYNAMIC blocks, compat mode (bytes == reference RawDeflate per chunk), joined into one stream")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bytes", type=int, default=WORKLOAD_BYTES, help="per-GPU workload bytes (default = C2)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", os.environ.get("BENCH_CLOCK_MS", "20")],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        # nvidia-smi needs a moment before its first line: wait for it so that the timed region is covered
        t0 = time.time()
        while self.p is not None and time.time() - t0 < 3.0 and os.path.getsize(self.f.name) == 0:
            time.sleep(0.02)

    def mark(self):
        """call at the start of the timed region: only samples written after this point are reported"""
        self.f.flush()
        self.mark_offset = os.path.getsize(self.f.name)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()

        def parse(text):
            sm, mx, reasons = [], [], set()
            for line in text.splitlines():
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1]))
                    mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons

        self.f.seek(0)
        text = self.f.read()
        off = getattr(self, "mark_offset", 0)
        sm, mx, reasons = parse(text[off:])
        window = "timed region"
        if len(sm) < 2:  # region shorter than two sampling periods: report the whole run under load, say so
            sm, mx, reasons = parse(text)
            window = "warm-up + timed region"
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       window=window)
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_deflate_rate(data, threads, target_s, max_bytes=None):
    """Times oracle RawDeflate per 64 KiB chunk over a bounded prefix sample of `data`.
    Returns (GB/s, sample bytes, compressed bytes of the sample, seconds)."""
    import oracle
    n = len(data)
    probe = min(n, 32 * CHUNK * max(1, threads // 4))
    t0 = time.perf_counter()
    oracle.deflate_chunks_mt(data[:probe], CHUNK, oracle.DYNAMIC, threads)
    dt = max(time.perf_counter() - t0, 1e-4)
    sample = int(min(n, max(probe, probe * target_s / dt)))
    if max_bytes:
        sample = min(sample, max_bytes)
    sample = max(CHUNK, sample // CHUNK * CHUNK)
    t0 = time.perf_counter()
    cbytes = oracle.deflate_chunks_mt(data[:sample], CHUNK, oracle.DYNAMIC, threads)
    dt = time.perf_counter() - t0
    return sample / dt / 1e9, sample, cbytes, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    import oracle
    oracle.build()
    from zlibts_b200 import synth
    threads = os.cpu_count() or 1
    # a prefix of the C2 buffer is enough for the bounded samples
    data = synth.mixed(min(args.bytes, max(64 << 20, threads * (4 << 20))), 2)
    per_step_s = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    # size the sample once, then time K steps on it
    _, sample, _, _ = cpu_deflate_rate(data, threads, per_step_s)
    for _ in range(args.warmup):
        oracle.deflate_chunks_mt(data[:sample], CHUNK, oracle.DYNAMIC, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cbytes = oracle.deflate_chunks_mt(data[:sample], CHUNK, oracle.DYNAMIC, threads)
    dt = time.perf_counter() - t0
    gbs = sample * args.steps / dt / 1e9
    # single-thread figure (north_star: "both single-threaded and across all stated cores")
    st_gbs, st_sample, _, _ = cpu_deflate_rate(data, 1, 3.0)
    sample_txt = f"first {sample >> 10} KiB of the C2 buffer ({sample // CHUNK} chunks) per step, {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "chunk_bytes": CHUNK, "sampled": sample_txt,
                   "note": "reference = C restatement of zlib.ts RawDeflate (oracle/); Node is absent from this image"},
        "cpu_baseline": {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_txt,
                         "single_thread_value": st_gbs, "ratio": cbytes / sample},
        "e2e": {"value": gbs, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import zlibts_b200 as z
    from zlibts_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    z.load_library()
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.bytes // CHUNK * CHUNK
    n_chunks = n // CHUNK
    data = synth.mixed(n, 2 + rank)
    stream = torch.cuda.Stream()
    eng = z.Engine(local_rank, stream.cuda_stream)
    cap = z.deflate_bound(n)
    cap_primed = z.deflate_bound(n, 0, z.DYNAMIC, z.MODE_PRIMED)   # twice the chunks: larger worst case
    h_in = torch.from_numpy(data).pin_memory()
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    with torch.cuda.stream(stream):
        d_in = h_in.cuda(non_blocking=True)
        d_out = torch.empty(cap_primed, dtype=torch.uint8, device="cuda")
    stream.synchronize()
    items = z.make_items(1)
    items["in_len"], items["out_cap"] = n, cap

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K calls bracketed by barrier + synchronize; CUDA events on the engine's stream; max over ranks (ms)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    res = {}

    # every rank but the last leaves its stream open, so the ranks' outputs concatenate into one stream
    dflags = z.DEFLATE_NOT_FINAL if rank < world - 1 else 0

    def step_device():
        res["r"] = eng.deflate_batch(d_in, d_out, items, flags=dflags)

    def step_host():
        res["h"] = eng.deflate_batch_host(h_in, h_out, items, flags=dflags)

    # ---- device-resident leg (value) ---------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(args.warmup):
        step_device()
    eng.profile_enable(True)
    eng.profile_reset()
    clocks.mark()
    l0 = eng.launch_count
    ms = timed(step_device, args.steps)
    launches = eng.launch_count - l0
    clk = clocks.stop()
    prof = eng.profile_read()
    eng.profile_enable(False)
    r = res["r"]
    assert int(r["status"][0]) == 0, r
    clen = int(r["out_len"][0])
    value = world * n * args.steps / (ms * 1e-3) / 1e9

    # ---- fast mode (bounded candidate depth; valid stream, not byte-identical): speed and ratio beside compat ----
    for _ in range(2):
        rf = eng.deflate_batch(d_in, d_out, items, flags=dflags, mode=z.MODE_FAST)
    ms_f = timed(lambda: res.__setitem__("f", eng.deflate_batch(d_in, d_out, items, flags=dflags, mode=z.MODE_FAST)),
                 max(1, min(args.steps, 5)))
    fast_steps = max(1, min(args.steps, 5))
    fast = {"value": world * n * fast_steps / (ms_f * 1e-3) / 1e9, "unit": UNIT, "chain_depth": 16,
            "ratio": int(res["f"]["out_len"][0]) / n, "ratio_vs_compat": int(res["f"]["out_len"][0]) / clen,
            "note": "ratio tolerance vs reference RawDeflate per chunk: 3 % (north_star); compat ratio == reference"}
    # ---- primed mode (32 KiB of history in front of every 32 KiB chunk): ratio recovered, speed paid
    items_p = items.copy()
    items_p["out_cap"] = cap_primed
    primed = {}
    for name, mode in (("exhaustive", z.MODE_PRIMED), ("fast", z.MODE_FAST | z.MODE_PRIMED)):
        for _ in range(2):
            eng.deflate_batch(d_in, d_out, items_p, flags=dflags, mode=mode)
        ms_p = timed(lambda: res.__setitem__("p", eng.deflate_batch(d_in, d_out, items_p, flags=dflags, mode=mode)),
                     fast_steps)
        assert int(res["p"]["status"][0]) == 0
        primed[name] = {"value": world * n * fast_steps / (ms_p * 1e-3) / 1e9, "unit": UNIT,
                        "ratio": int(res["p"]["out_len"][0]) / n, "ratio_vs_compat": int(res["p"]["out_len"][0]) / clen}
    primed["note"] = ("32 KiB chunks, each searching the 32 KiB before it as well; exhaustive = the reference's matcher "
                      "(blocks equal the oracle's block construction with that history), fast = depth 16")
    step_device()  # leave the compat output in d_out

    # ---- end-to-end leg: host buffers through the C-ABI host entry point (H2D + D2H inside) -----------
    for _ in range(max(1, min(2, args.warmup))):
        step_host()
    e2e_steps = max(1, min(args.steps, 5))
    ms_h = timed(step_host, e2e_steps)
    assert int(res["h"]["status"][0]) == 0 and int(res["h"]["out_len"][0]) == clen
    e2e_value = world * n * e2e_steps / (ms_h * 1e-3) / 1e9

    # ---- inflate leg: the chunks as independent streams (C3-shaped), device resident + host e2e -------
    slot = z.deflate_bound(CHUNK)
    it_c = z.make_items(n_chunks)
    it_c["in_off"] = np.arange(n_chunks, dtype=np.uint64) * CHUNK
    it_c["in_len"] = CHUNK
    it_c["out_off"] = np.arange(n_chunks, dtype=np.uint64) * slot
    it_c["out_cap"] = slot
    with torch.cuda.stream(stream):
        d_slots = torch.empty(n_chunks * slot, dtype=torch.uint8, device="cuda")
    rc = eng.deflate_batch(d_in, d_slots, it_c)
    assert int(rc["status"].max()) == 0
    # pack the streams back to back (what a zip / multi-member container holds)
    lens = rc["out_len"].astype(np.uint64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)
    total_c = int(lens.sum())
    h_slots = d_slots.cpu().numpy()
    h_packed = torch.empty(total_c, dtype=torch.uint8).pin_memory()
    hp = h_packed.numpy()
    for k in range(n_chunks):
        hp[int(offs[k]):int(offs[k] + lens[k])] = h_slots[k * slot:k * slot + int(lens[k])]
    del h_slots, d_slots
    with torch.cuda.stream(stream):
        d_packed = h_packed.cuda(non_blocking=True)
        d_plain = torch.empty(n, dtype=torch.uint8, device="cuda")
    it_i = z.make_items(n_chunks)
    it_i["in_off"], it_i["in_len"] = offs, lens
    it_i["out_off"] = np.arange(n_chunks, dtype=np.uint64) * CHUNK
    it_i["out_cap"] = CHUNK
    h_plain = torch.empty(n, dtype=torch.uint8).pin_memory()

    def inf_device():
        res["i"] = eng.inflate_batch(d_packed, d_plain, it_i)

    def inf_host():
        res["ih"] = eng.inflate_batch_host(h_packed, h_plain, it_i)

    for _ in range(args.warmup):
        inf_device()
    eng.profile_enable(True)
    eng.profile_reset()
    li0 = eng.launch_count
    ms_i = timed(inf_device, args.steps)
    inf_launches = eng.launch_count - li0
    prof_i = eng.profile_read()
    eng.profile_enable(False)
    assert int(res["i"]["status"].max()) == 0
    assert torch.equal(d_plain, d_in), "inflate(deflate(x)) != x"
    inf_value = world * n * args.steps / (ms_i * 1e-3) / 1e9
    inf_host()
    ms_ih = timed(inf_host, e2e_steps)
    assert torch.equal(h_plain, h_in)
    inf_e2e = world * n * e2e_steps / (ms_ih * 1e-3) / 1e9

    # ---- the one cross-rank exchange: exclusive scan of the per-rank output sizes ----------------------
    rank_off = 0
    if dist is not None:
        sizes = torch.zeros(world, dtype=torch.int64, device="cuda")
        sizes[rank] = clen
        dist.all_reduce(sizes)
        rank_off = int(sizes[:rank].sum().item())
        total_clen = int(sizes.sum().item())
    else:
        total_clen = clen

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    peak, peak_src = peaks()
    top = max(prof.items(), key=lambda kv: kv[1]["ms"])
    top_name, top_ms, top_launches = top[0], top[1]["ms"], top[1]["launches"]
    step_kernel_ms = sum(v["ms"] for v in prof.values())
    alg_bytes_per_step = n + clen                      # N + C per chunk, summed over the step's chunks
    alg_bytes_per_launch = alg_bytes_per_step * args.steps / max(1, top_launches)
    avg_launch_ms = top_ms / max(1, top_launches)
    achieved = alg_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_per_launch, "avg_launch_ms": avg_launch_ms,
                "kernel_share_of_step": top_ms / max(step_kernel_ms, 1e-9),
                "kernels_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items() if v["launches"]}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tj = json.load(open(traffic_file))
            # the capture may be of a smaller launch than this run's: traffic scales with the chunks per launch
            per_unit = tj.get(top_name, 0) / max(1, tj.get("_units_per_launch", 1))
            roofline["traffic"] = per_unit * (n_chunks * args.steps / max(1, top_launches)) or None
            roofline["traffic_source"] = tj.get("_note")
        except Exception:
            pass
    inf_top_ms = prof_i["inflate_warp_kernel"]["ms"] / max(1, prof_i["inflate_warp_kernel"]["launches"])
    inf_roofline = {"bound": "hbm", "kernel": "inflate_warp_kernel",
                    "achieved": (n + total_c) / (inf_top_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s"}
    inf_roofline["frac"] = inf_roofline["achieved"] / peak

    # ---- CPU baseline beside it (rank 0, N = 1 only): oracle port on a bounded sample ------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        oracle.build()
        threads = os.cpu_count() or 1
        gbs, sample, cbytes, secs = cpu_deflate_rate(data, threads, args.cpu_seconds, max_bytes=n)
        st_gbs, _, _, _ = cpu_deflate_rate(data, 1, 3.0)
        # the sample doubles as a parity check: compat bytes => identical sizes for those chunks
        gpu_sample = int(rc["out_len"][: sample // CHUNK].sum())
        cpu = {"value": gbs, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {sample >> 20} MiB of the C2 buffer ({sample // CHUNK} chunks), {secs:.1f} s, "
                         f"{threads} threads; C restatement of reference RawDeflate (Node absent)",
               "single_thread_value": st_gbs, "sample_ratio": cbytes / sample,
               "gpu_bytes_equal_on_sample": bool(gpu_sample == cbytes)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "bytes_per_gpu": n, "chunk_bytes": CHUNK, "chunks_per_gpu": n_chunks,
                       "mode": "compat", "block_type": "DYNAMIC", "parallelism": f"shard{world}",
                       "l2": "inputs (256 MiB) larger than L2 (126 MB), no flush"},
            "ratio": clen / n,
            "fast_mode": fast,
            "primed_mode": primed,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": clen,
                    "ms_per_step": ms_h / e2e_steps, "api": "zlb_deflate_batch_host (pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": {"sm_mhz": clk["sm_mhz"], "sm_max_mhz": clk["sm_max_mhz"], "reasons": clk["reasons"],
                       "samples": clk["samples"], "window": clk.get("window")},
            "roofline": roofline,
            "cpu_baseline": cpu,
            "inflate": {"metric": "inflate_output_GBps", "value": inf_value, "unit": UNIT, "streams_per_gpu": n_chunks,
                        "ms_per_step": ms_i / args.steps, "gpu_launches": int(inf_launches),
                        "e2e": {"value": inf_e2e, "unit": UNIT, "h2d_bytes_per_step": total_c,
                                "d2h_bytes_per_step": n, "ms_per_step": ms_ih / e2e_steps},
                        "roofline": inf_roofline},
            "multi_gpu": {"output_bytes_total": total_clen, "rank0_offset": rank_off,
                          "exchange": "exclusive scan of per-rank output sizes (8-byte all-reduce), no data-path collective"},
        }
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else the libraries print (NCCL banner, ...) was
    redirected to stderr at start-up so that stdout carries exactly one line."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import __graft_entry__ as g
    if rank == 0 or not os.path.exists(os.path.join(ROOT, "zlib.ts_b200", "libzlibts_b200.so")):
        g.build()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
