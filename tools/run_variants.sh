python -m pytest tests/test_deflate_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== base"; python tools/probe_kinds.py 64 2>&1 | grep compat
for v in r4 cap32 cap128; do echo "== $v"; ZLB_LIB_OVERRIDE=$PWD/variants/$v.so python tools/probe_kinds.py 64 2>&1 | grep compat; done
