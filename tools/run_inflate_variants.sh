#!/bin/bash
# Development aid (on the GPU box): batched inflate time for the built library and every variants/*.so
for rep in 1 2; do
echo "== base"; for n in 4096 32768; do python tools/probe_inflate.py $n 2>&1 | tail -1; done
for f in variants/*.so; do v=$(basename $f .so); echo "== $v"; for n in 4096 32768; do ZLB_LIB_OVERRIDE=$PWD/$f python tools/probe_inflate.py $n 2>&1 | tail -1; done; done
done
