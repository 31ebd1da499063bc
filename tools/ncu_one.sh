#!/bin/bash
# ncu --set full capture of one kernel (regex) of a 64 MiB bench run. usage: bash tools/ncu_one.sh <regex> <tag> [skip] [count]
rx=$1; tag=$2; skip=${3:-1}; cnt=${4:-1}
mkdir -p gpurun_out
SMALL="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --bytes 67108864"
$SMALL > gpurun_out/plain_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o gpurun_out/prof_$tag $SMALL > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/plain_$tag.log | cut -c1-600
