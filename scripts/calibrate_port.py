"""How fast is the oracle's restated tracker (oracle/tracker_port.py: LinkerPort) compared with the reference's own
CentroidTracker + GaussianSumFIR (ysmr/tracker.py, ysmr/gsff.py)?  bench.py's CPU legs time the port (`cpu_baseline.kind =
"port"`); this script measures, in the build container (needs /root/reference), how the tracker half of that figure
relates to the real classes and writes oracle/port_calibration.json, which bench.py attaches to its cpu_baseline block.

  python scripts/calibrate_port.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import import_reference, random_detection_sequence  # noqa: E402
from oracle.tracker_port import LinkerPort  # noqa: E402


def rects_of(rec):
    return [((float(r[0]), float(r[1])), (float(r[2]), float(r[3]), float(r[4]))) for r in rec]


def main():
    _, _, tracker = import_reference()
    cases = {
        'cfg2-like (50 tracks)': dict(n_frames=300, n_cells=50, width=1228., height=922.),
        'cfg4-like (200 tracks)': dict(n_frames=120, n_cells=200, width=2048., height=2048.),
        'cfg3-like (2,000 tracks)': dict(n_frames=40, n_cells=2000, width=1228., height=922., p_miss=0.02),
    }
    out = {'host': {'cpus': os.cpu_count()}, 'cases': {}}
    for name, kw in cases.items():
        seq = [rects_of(r) for r in random_detection_sequence(np.random.default_rng(7), **kw)]
        ref = tracker.CentroidTracker(max_disappeared=30.0, fps=30.0)
        t0 = time.perf_counter()
        ids_ref = []
        for rects in seq:
            objects, _ = ref.update(rects)
            ids_ref.append(list(objects.keys()))
        t_ref = time.perf_counter() - t0
        port = LinkerPort(max_disappeared=30.0, fps=30.0)
        t0 = time.perf_counter()
        ids_port = []
        for rects in seq:
            ids_port.append([i for (i, _, _) in port.update(rects)])
        t_port = time.perf_counter() - t0
        assert ids_ref == ids_port, name
        out['cases'][name] = {'frames': len(seq), 'reference_ms_per_frame': 1e3 * t_ref / len(seq),
                              'port_ms_per_frame': 1e3 * t_port / len(seq), 'reference_over_port': t_ref / t_port}
        print(name, out['cases'][name])
    with open(os.path.join(ROOT, 'oracle', 'port_calibration.json'), 'w') as fh:
        json.dump(out, fh, indent=1)


if __name__ == '__main__':
    main()
