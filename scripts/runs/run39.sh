set -x
LINKPROF_FRAMES=4096 LINKPROF_PLAIN=1 ncu --set full --clock-control none --import-source on -k regex:link_kernel -s 1 -c 1 -o gpurun_out/prof_link_v5 python scripts/linkprof.py > gpurun_out/ncu_link_v5.log 2>&1
ls -la gpurun_out/prof_link_v5.ncu-rep
