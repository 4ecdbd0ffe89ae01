#!/usr/bin/env python
"""Turns ncu output brought back from the GPU box (gpurun_out/) into the tracked summaries under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches_summary.md "command line"
  python scripts/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r1_ncu_full_summary.md "command line" [frames_per_launch]

`full` also updates profiles/r2_traffic.json (DRAM bytes per frame of the front-end kernels, per configuration) which
bench.py reports as roofline.traffic.
"""
import collections
import csv
import json
import os
import subprocess
import sys

OURS = 'ysmr::'


def launches(src, dst, cmd):
    rows = list(csv.reader(open(src)))
    hdr, agg, geo = None, collections.defaultdict(list), {}
    for r in rows:
        if len(r) > 5 and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get('Metric Name') == 'gpu__time_duration.sum':
                v = float(d['Metric Value'].replace(',', ''))
                v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(d['Metric Unit'], 1e-6)
                name = d['Kernel Name']
                agg[name].append(v)
                geo[name] = (d.get('Grid Size', ''), d.get('Block Size', ''))
    names = ('fused_front', 'blur_prepass', 'plane_margins', 'gauss_decide', 'pack_masks', 'label_kernel', 'geometry', 'link_kernel',
             'link_prep', 'link_general', 'link_reset', 'scalar_decide', 'frame_moments', 'moving_threshold', 'frontend_tile',
             'unpack_bits', 'rows_', 'select_')
    ours = {k: v for k, v in agg.items() if OURS in k or k.startswith('void ysmr') or any(n in k for n in names)}
    total = sum(sum(v) for v in ours.values())
    with open(dst, 'w') as f:
        f.write(f'# ncu launch list\n\nCommand: `{cmd}`\n\n')
        f.write('`--metrics gpu__time_duration.sum --clock-control none`: per-launch times are cold-cache and serialised, so compare '
                'SHARES, not absolutes (torch kernels of the synthetic renderer omitted).\n\n')
        f.write('| kernel | launches | total ms | ms / launch | share of our kernels | grid | block |\n|---|---|---|---|---|---|---|\n')
        for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
            short = k.split('(')[0].replace('void ', '')
            f.write(f'| `{short}` | {len(v)} | {sum(v):.3f} | {sum(v) / len(v):.4f} | {100 * sum(v) / total:.1f} % | {geo[k][0]} | {geo[k][1]} |\n')
        det = {k: v for k, v in ours.items() if 'link_' not in k}
        dtot = sum(sum(v) for v in det.values())
        f.write('\nIn the un-profiled run `link_kernel` sits on its own stream and SM and overlaps detection of the next chunk, so '
                'the step time is the detection stream. Shares of the detection stream alone: ')
        f.write(', '.join(f"`{k.split('(')[0].replace('void ', '')}` {100 * sum(v) / dtot:.1f} %"
                          for k, v in sorted(det.items(), key=lambda kv: -sum(kv[1]))) + '.\n')
        f.write(f'\nRaw list: `{os.path.basename(src)}` (same directory).\n')
    print(open(dst).read())


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__icc_request_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']


def full(src, dst, cmd, frames_per_launch, config='cfg2'):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = collections.defaultdict(list)
    with open(dst, 'w') as f:
        f.write(f'# ncu --set full capture\n\nCommand: `{cmd}`\n\n')
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            name = d['Kernel Name']
            f.write(f'## {name[:110]}\n\n')
            for k in WANT:
                if k in d:
                    f.write(f'- {k}: {d[k]} {units[hdr.index(k)]}\n')
            stalls = []
            for k in hdr:
                if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
                    try:
                        v = float(d[k])
                    except ValueError:
                        continue
                    if v > 0.05:
                        stalls.append((k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
            f.write('- stall reasons (warps per issue): ' + ', '.join(f'{a} {b:.2f}' for a, b in stalls) + '\n')
            try:
                def to_bytes(key):
                    v = float(d[key].replace(',', ''))
                    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(units[hdr.index(key)], 1)
                b = to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')
                f.write(f'- dram traffic (read+write): {b / 1e6:.1f} MB per launch')
                if frames_per_launch:
                    f.write(f' = {b / frames_per_launch / 1e6:.3f} MB per frame ({frames_per_launch} frames per launch)')
                    traffic[name.split('(')[0].replace('void ', '')].append(b / frames_per_launch)
                f.write('\n')
            except Exception:
                pass
            f.write('\n')
    if frames_per_launch and traffic:
        per = {k: sum(v) / len(v) for k, v in traffic.items()}
        fe = sum(v for k, v in per.items() if any(t in k for t in ('fused_front', 'blur_prepass', 'plane_margins', 'gauss_decide', 'pack_masks')))
        # bench.py reads profiles/r2_traffic.json: front-end DRAM bytes per frame, per configuration
        path = os.path.join(os.path.dirname(dst), 'r2_traffic.json')
        out = json.load(open(path)) if os.path.isfile(path) else {'frontend_dram_bytes_per_frame': {}, 'detail': {}}
        out['frontend_dram_bytes_per_frame'][config] = fe
        out['detail'][config] = {'frames_per_launch': frames_per_launch, 'dram_bytes_per_frame': per, 'source': os.path.basename(dst)}
        json.dump(out, open(path, 'w'), indent=1)
    print(open(dst).read()[:6000])


if __name__ == '__main__':
    mode, src, dst, cmd = sys.argv[1:5]
    if mode == 'launches':
        launches(src, dst, cmd)
    else:
        full(src, dst, cmd, int(sys.argv[5]) if len(sys.argv) > 5 else 0, sys.argv[6] if len(sys.argv) > 6 else 'cfg2')
