#!/bin/bash
# One gpurun call: GPU tests, bench, ncu launch list, ncu full capture of the hot kernels.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi -L
nproc
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>&1; cat gpurun_out/bench_ref_$tag.json
SMALL="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$SMALL > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $SMALL > gpurun_out/ncu_launch_$tag.log 2>&1
echo "ncu launches rc=$?"
$SMALL > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"lz77_chunk|huffman_build|bitpack" -s 4 -c 3 -f -o gpurun_out/prof_deflate_$tag $SMALL > gpurun_out/ncu_full_$tag.log 2>&1
echo "ncu deflate rc=$?"
$SMALL > gpurun_out/plain4_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"lz77_fast" -s 1 -c 1 -f -o gpurun_out/prof_fast_$tag $SMALL > gpurun_out/ncu_full_fast_$tag.log 2>&1
echo "ncu fast rc=$?"
$SMALL > gpurun_out/plain3_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"inflate_warp" -s 1 -c 1 -f -o gpurun_out/prof_inflate_$tag $SMALL > gpurun_out/ncu_full_inf_$tag.log 2>&1
echo "ncu inflate rc=$?"
python tools/bench_configs.py c4:2000 > gpurun_out/plain5_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"frame_" -c 2 -f -o gpurun_out/prof_frame_$tag python tools/bench_configs.py c4:2000 > gpurun_out/ncu_full_frame_$tag.log 2>&1
echo "ncu frame rc=$?"
python tools/probe_kinds.py 64 > gpurun_out/kinds_$tag.log 2>&1; cat gpurun_out/kinds_$tag.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python tools/bench_configs.py c1 c3 c4 c5 > gpurun_out/configs_$tag.jsonl 2> gpurun_out/configs_$tag.err; echo "configs rc=$?"; cat gpurun_out/configs_$tag.jsonl
ls -la gpurun_out
