#!/bin/bash
# quick GPU check: deflate parity tests + LZ77 time per kind of data
mkdir -p gpurun_out
python -m pytest tests/test_deflate_gpu.py -x -q -m gpu 2>&1 | tail -8
python tools/probe_kinds.py 64 > gpurun_out/kinds_quick.log 2>&1; cat gpurun_out/kinds_quick.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 1500 gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
