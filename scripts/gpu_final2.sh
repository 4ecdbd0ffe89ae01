#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_cfg2.json').read().strip().splitlines()[-1])
print('cfg2', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['kernel_ms_per_step'], 'e2e', d['e2e']['value'], d['e2e_grey']['value'], d['e2e_ysmr'].get('value'), d['parity_check']['ok'])
PY
B="python bench.py --steps 2 --warmup 3 --frames 2368 --no-e2e --no-cpu"
YSMR_LINK=nogate timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_front|label_kernel|geometry|link_|rows_|blur_prepass|gauss_decide|pack_masks|plane_margins" -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
ls -la gpurun_out/r2_launches.csv
