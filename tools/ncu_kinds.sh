#!/bin/bash
# ncu --set full capture of lz77_chunk_kernel on one kind of data each. usage: bash tools/ncu_kinds.sh kind...
mkdir -p gpurun_out
for k in "$@"; do
  python tools/kind_case.py $k > gpurun_out/kind_$k.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"lz77_chunk" -s 1 -c 1 -f -o gpurun_out/prof_kind_$k python tools/kind_case.py $k > gpurun_out/ncu_kind_$k.log 2>&1
  echo "$k rc=$?"
done
