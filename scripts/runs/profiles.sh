set -x
# 4736 frames: the first chunks of the call are full 1184-frame batches (track_chunks tapers the chunk size towards the end)
CMD="python bench.py --frames 4736 --steps 1 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/r1_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"blur_prepass|plane_margins|gauss_decide|pack_masks|label_kernel|geometry|link_kernel|link_reset" -c 400 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/r1_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blur_prepass|plane_margins|gauss_decide|pack_masks|label_kernel|geometry_kernel|link_kernel" -c 9 -o gpurun_out/r1_full -f $CMD > gpurun_out/r1_ncu_full.log 2>&1
ls -la gpurun_out/r1_full.ncu-rep
