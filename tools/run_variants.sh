#!/bin/bash
# Development aid (on the GPU box): LZ77 time per kind of data for the built library and every variants/*.so
echo "== base"; python tools/probe_kinds.py 64 2>&1 | grep compat
for f in variants/*.so; do v=$(basename $f .so); echo "== $v"; ZLB_LIB_OVERRIDE=$PWD/$f python tools/probe_kinds.py 64 2>&1 | grep compat; done
