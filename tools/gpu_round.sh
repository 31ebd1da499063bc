#!/bin/bash
# One gpurun call: GPU tests, bench (both arms), ncu launch list, ncu full capture of the hot kernels, per-kind times.
# usage: gpurun --timeout 1800 -- 'bash tools/gpu_round.sh [tag]'
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi -L
nproc
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>&1; tail -c 300 gpurun_out/bench_ref_$tag.json
SMALL="timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --configs none"
$SMALL > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $SMALL > gpurun_out/ncu_launch_$tag.log 2>&1
echo "ncu launches rc=$?"
$SMALL > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"lz77_chunk|huffman_build|bitpack" -s 4 -c 3 -f -o gpurun_out/prof_deflate_$tag $SMALL > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu deflate rc=$?"
$SMALL > gpurun_out/plain3_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"inflate_warp" -s 1 -c 1 -f -o gpurun_out/prof_inflate_$tag $SMALL > gpurun_out/ncu_full_inf_$tag.log 2>&1
echo "ncu inflate rc=$?"
timeout 200 python tools/probe_kinds.py 64 > gpurun_out/kinds_$tag.log 2>&1; cat gpurun_out/kinds_$tag.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
ls -la gpurun_out | tail -12
