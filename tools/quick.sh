#!/bin/bash
# quick GPU check: deflate parity tests + per-kernel timings (mixed 256 MiB, text 64 MiB)
mkdir -p gpurun_out
python -m pytest tests/test_deflate_gpu.py -x -q -m gpu 2>&1 | tail -5
python tools/probe.py 256 mixed > gpurun_out/p1.log 2>&1; grep -A5 "^deflate" gpurun_out/p1.log | tail -6; grep "^inflate" gpurun_out/p1.log | tail -1
python tools/probe.py 64 text > gpurun_out/p2.log 2>&1; grep -A5 "^deflate" gpurun_out/p2.log | tail -6; grep "^inflate" gpurun_out/p2.log | tail -1
