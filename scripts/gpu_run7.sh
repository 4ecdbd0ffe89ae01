#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for mode in gate nogate; do
  if [ $mode = nogate ]; then export YSMR_LINK=nogate; else unset YSMR_LINK; fi
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${mode}.json 2> gpurun_out/bench_${mode}.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_${mode}.json').read().strip().splitlines()[-1]); print('${mode}', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
unset YSMR_LINK
for cfg in cfg3 cfg4; do
  timeout 600 python bench.py --config $cfg --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_${cfg}_q.json 2> gpurun_out/bench_${cfg}_q.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_${cfg}_q.json').read().strip().splitlines()[-1]); print('${cfg}', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
