#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun): tests, smoke, the bench lines of every configuration, the reference
# arm; with "ncu" as first argument also the ncu passes (launch list + --set full captures) of the same bench command
# (YSMR_LINK=nogate there: per-launch times without the flag wait).
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for cfg in cfg3 cfg4; do
  timeout 900 python bench.py --config $cfg > gpurun_out/bench_${cfg}.json 2> gpurun_out/bench_${cfg}.err
done
python - <<'PY'
import json
for c in ('cfg2', 'cfg3', 'cfg4', 'reference'):
    try:
        d = json.loads(open('gpurun_out/bench_%s.json' % c).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        print(c, round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'frac', r.get('frac'), 'whole', r.get('whole_path_frac'), r.get('kernel_ms_per_step'),
              'link us', r.get('linker_us_per_frame'), 'e2e', (d.get('e2e') or {}).get('value'), 'grey', (d.get('e2e_grey') or {}).get('value'),
              'ysmr', (d.get('e2e_ysmr') or {}).get('value'), 'parity', (d.get('parity_check') or {}).get('ok'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as ex:
        print(c, 'FAILED', ex)
PY
if [ "$1" = ncu ]; then
  B="python bench.py --steps 2 --warmup 3 --frames 2368 --no-e2e --no-cpu"
  export YSMR_LINK=nogate
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fused_front|label_kernel|geometry|link_|rows_|blur_prepass|gauss_decide|pack_masks|plane_margins" -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_front -s 3 -c 1 -f -o gpurun_out/r2_fused_final $B > gpurun_out/r2_ncu_fused.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"link_kernel|label_kernel|geometry_kernel" -s 9 -c 3 -f -o gpurun_out/r2_back_final $B > gpurun_out/r2_ncu_back.log 2>&1
  ls -la gpurun_out/r2_*final* gpurun_out/r2_launches.csv
fi
