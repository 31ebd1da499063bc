#!/bin/bash
# Development aid: builds kernel variants of libzlibts_b200.so into variants/ (name=flags pairs) for A/B runs on
# the GPU: ZLB_LIB_OVERRIDE=variants/<name>.so python tools/probe_kinds.py 64
# usage: bash tools/build_variants.sh name1="-DLZ_PRIV_CAP=128" name2="-DLZ_KMAX=16 -DLZ_WAIT_MUL=1" ...
cd "$(dirname "$0")/../zlib.ts_b200/csrc" || exit 1
mkdir -p ../../variants
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 $flags -Xcompiler -fPIC -shared \
    -o ../../variants/$name.so zts_ctx.cu zts_hoststage.cu zts_checksum.cu zts_inflate.cu zts_lz77.cu zts_huffman.cu zts_deflate.cu zts_container.cu &
done
wait
ls -la ../../variants
