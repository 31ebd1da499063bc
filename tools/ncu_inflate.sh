#!/bin/bash
# Development aid (on the GPU box): one ncu --set full capture of inflate_warp_kernel on N streams of 64 KiB
# usage: bash tools/ncu_inflate.sh [streams] -> gpurun_out/prof_inf_warp.ncu-rep
N=${1:-32768}
mkdir -p gpurun_out
python tools/probe_inflate.py $N > gpurun_out/inf_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"inflate_warp" -s 1 -c 1 -f -o gpurun_out/prof_inf_warp python tools/probe_inflate.py $N > gpurun_out/ncu_inf_warp.log 2>&1
echo rc=$?; tail -1 gpurun_out/inf_plain.log
