#!/bin/bash
# Round-2 profiling pass (run under gpurun): plain run first, then the ncu launch list and the --set full captures.
set -x
B="python bench.py --steps 2 --warmup 3 --frames 2368 --no-e2e --no-cpu"
$B > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_front -s 2 -c 2 -f -o gpurun_out/r2_fused_final $B > gpurun_out/r2_ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"link_kernel|link_prep|label_kernel|geometry_kernel" -s 8 -c 6 -f -o gpurun_out/r2_back_final $B > gpurun_out/r2_ncu_back.log 2>&1
ls -la gpurun_out/r2_*final* gpurun_out/r2_launches.csv
