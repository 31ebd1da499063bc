mkdir -p gpurun_out
python tools/probe_inflate.py 4096 > gpurun_out/inf_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"inflate_lane" -s 1 -c 1 -f -o gpurun_out/prof_inf_lane python tools/probe_inflate.py 4096 > gpurun_out/ncu_inf_lane.log 2>&1
echo rc=$?
