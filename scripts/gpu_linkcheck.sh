#!/bin/bash
# quick check of a linker change: parity tests, kernel time alone / beside a copy loop, the cfg2 bench line
timeout 900 python -m pytest tests/test_gpu_link.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -2
python scripts/link_alone.py cfg2 4736 2>&1 | tail -1; python scripts/link_alone.py cfg2 4736 hog 2>&1 | tail -1
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1]); print('cfg2', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
