set -x
CMD="python bench.py --frames 1184 --steps 2 --warmup 3 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"blur_prepass|plane_margins|gauss_decide|pack_masks|label_kernel|geometry|link_kernel|link_reset|scalar_decide|frame_moments|moving_threshold" -c 400 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/r1_ncu_launches.log 2>&1
