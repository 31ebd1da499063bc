python -m pytest tests/test_inflate_gpu.py -x -q -m gpu 2>&1 | tail -3
for g in 32 16 8; do ZLB_INFLATE_GROUP=$g python tools/probe_inflate.py 16384 2>&1 | tail -1; done
for g in 32 16 8; do ZLB_INFLATE_GROUP=$g python -m pytest tests/test_inflate_gpu.py -x -q -m gpu 2>&1 | tail -1; done
