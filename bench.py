#!/usr/bin/env python
"""bench.py -- frames/s of the detect + link (+GSFF) hot path at 1228x922 on N B200s, with the HBM roofline of the
dominant kernel and the reference's CPU path timed beside it.

Contract (one JSON line on stdout from rank 0):
  python bench.py --gpus N --steps K --warmup W            # ours
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU implementation of the same path

Workload at N=1: BASELINE.json configs[1] -- the cfg1 scene (1228x922, ~50 rod-shaped bacteria, default tracking.ini)
continued to 9,000 frames, BGR frames as cap.read() delivers them, resident in HBM.  A "step" is one pass of
detect+link+gsff over the whole 9,000-frame video (30.6 GB of input, far larger than L2).  At N>1 every rank holds its
own 9,000-frame range of an N*9,000-frame video (weak scaling); detections are gathered over NCCL and linked by the one
sequential linker on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='cfg2', choices=['cfg2', 'cfg3', 'cfg4', 'cfg5'],
                    help='BASELINE.json configuration (default cfg2 = configs[1], the one the metric is quoted on)')
    ap.add_argument('--frames', type=int, default=0, help='frames per GPU (0: the default of the configuration)')
    ap.add_argument('--channels', type=int, default=3, choices=[1, 3])
    ap.add_argument('--cells', type=int, default=0, help='override the number of cells of the scene')
    ap.add_argument('--batch', type=int, default=0,
                    help='frames per detect launch (0: 1184 on one GPU, 592 = one chunk of the streamed hand-over on several)')
    ap.add_argument('--cpu-frames', type=int, default=0, help='frames of the bounded CPU sample (0: about 15 s for the configuration)')
    ap.add_argument('--e2e-frames', type=int, default=1000, help='frames in the pinned host buffer of the e2e leg')
    ap.add_argument('--multi', default='stream', choices=['stream', 'gather'],
                    help='N > 1: stream detection records chunk by chunk to the linker (default) or gather whole ranges')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--ysmr-frames', type=int, default=600, help='frames of the FFV1 video of the drop-in leg (0: skip)')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    setup_config(args)
    return args


# ---- the BASELINE.json configurations ------------------------------------------------------------------------------------
# frames: per GPU (cfg5: of the whole video, divided over the GPUs = strong scaling).  cfg3 / cfg4 run a prefix of the 9,000 /
# 3,000 frames BASELINE names so that rendering the synthetic video and the default run stay within minutes; the prefix
# length is part of config.workload.
CONFIG_TABLE = {
    'cfg2': dict(scene='cfg2', frames=9000, cpu_frames=900, max_blobs=128, max_tracks=1024, batch1=1184, white=True,
                 what='cfg2: 1228x922, ~50 rods, default tracking.ini'),
    'cfg3': dict(scene='cfg3', frames=592, cpu_frames=40, max_blobs=4096, max_tracks=8192, batch1=296, white=True,
                 what='cfg3 (dense field): 1228x922, 2,000 rods/frame, adaptive double threshold 2.0'),
    'cfg4': dict(scene='cfg4', frames=1480, cpu_frames=200, max_blobs=1024, max_tracks=4096, batch1=296, white=False,
                 what='cfg4: 2048x2048, ~200 coccoid cells, dark on light'),
    'cfg5': dict(scene='cfg2', frames=54000, cpu_frames=900, max_blobs=128, max_tracks=1024, batch1=1184, white=True,
                 what='cfg5: one 54,000-frame 1228x922 video (cfg2 scene) sharded by frame range'),
}


def setup_config(args):
    t = CONFIG_TABLE[args.config]
    world = int(os.environ.get('WORLD_SIZE', '1'))
    args.table = t
    args.strong = args.config == 'cfg5'
    if not args.frames:
        args.frames = t['frames'] // world if args.strong else t['frames']
    if not args.cpu_frames:
        args.cpu_frames = t['cpu_frames']
    args.white = t['white']


def scene_config(args, n_total):
    import dataclasses
    from ysmr_b200.synth import CONFIGS
    cfg = dataclasses.replace(CONFIGS[args.table['scene']], n_frames=n_total)
    if args.cells:
        cfg = dataclasses.replace(cfg, n_cells=args.cells)
    return cfg


def detect_settings(args):
    from oracle import ref_stages
    return ref_stages.DetectSettings(white_on_dark=args.white)


# ---- clocks ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 10 ms from a thread (NVML); `mark()` brackets the timed region so
    that only samples taken DURING it are reported.  Falls back to one nvidia-smi query when NVML is unavailable."""

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.samples = []          # (t, sm_mhz, reasons_bitmask)
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None
        self.t0 = self.t1 = None

    def _run(self):
        import pynvml as nv
        h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.perf_counter(), float(mhz), int(rs)))
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML indexes physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            if vis:
                try:
                    self.gpu = int(vis.split(',')[self.gpu])
                except Exception:
                    pass
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def mark(self, which):
        if which == 0:
            self.t0 = time.perf_counter()
        else:
            self.t1 = time.perf_counter()

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            inside = [x for x in self.samples if self.t0 is not None and self.t0 <= x[0] <= self.t1]
            if not inside and self.samples and self.t0 is not None:      # very short region: nearest samples
                mid = 0.5 * (self.t0 + self.t1)
                inside = sorted(self.samples, key=lambda x: abs(x[0] - mid))[:3]
            if inside:
                names = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown', 0x4: 'sw_power_cap'}
                bits = 0
                for x in inside:
                    bits |= x[2]
                out = {'sm_mhz': float(np.median([x[1] for x in inside])), 'sm_max_mhz': self.max_mhz,
                       'reasons': sorted(v for k, v in names.items() if bits & k), 'samples': len(inside)}
            return out
        try:
            q = subprocess.run(['nvidia-smi', f'--id={self.gpu}', '--query-gpu=clocks.sm,clocks.max.sm', '--format=csv,noheader,nounits'],
                               capture_output=True, text=True, timeout=10).stdout.strip().split(',')
            out = {'sm_mhz': float(q[0]), 'sm_max_mhz': float(q[1]), 'reasons': [], 'samples': 1, 'note': 'single nvidia-smi query after the run'}
        except Exception:
            pass
        return out


# ---- CPU baseline: the reference's loop body (oracle = cv2/scipy call sites + restated linker) ---------------------------
def cpu_reference_fps(frames_bgr_or_grey, st, fps=30.0, keep_rows=False):
    """Replays track_eval.py:180-316 on in-RAM frames (oracle/; the reference itself cannot travel to the GPU box).
    Returns (frames/s, seconds, rows or row count, cv2 threads); rows = [(frame, id, x, y, w, h, deg)]."""
    import cv2
    from oracle import ref_stages
    from oracle.tracker_port import LinkerPort
    lp = LinkerPort(max_disappeared=fps, fps=fps)
    t0 = time.perf_counter()
    rows = []
    n_rows = 0
    for t, f in enumerate(frames_bgr_or_grey):
        r = ref_stages.detect_frame(f, st)
        out = lp.update(r['rects'])
        n_rows += len(out)
        if keep_rows:
            rows += [(t, i, xy[0], xy[1], info[0], info[1], info[2]) for (i, xy, info) in out]
    dt = time.perf_counter() - t0
    return len(frames_bgr_or_grey) / dt, dt, (np.array(rows, np.float64).reshape(-1, 7) if keep_rows else n_rows), cv2.getNumThreads()


def scene_for(args, n_total):
    from ysmr_b200.synth import make_scene
    return make_scene(scene_config(args, n_total))


def parity_check(gpu_rows, ref_rows):
    """GPU rows of the first frames of the timed video against the oracle's rows of the same bytes (north_star bars): frame
    and track id bit-exact; w/h/deg within 1e-3 (exact-area ties of cv2.minAreaRect are counted); x/y (GSFF)
    within 1e-5 relative on EVERY row -- no exemption for coasting tracks: the device restates NumPy's roundings
    (csrc/link.cuh), the fraction of rows whose float64 bits equal the oracle's is reported as well."""
    out = {'frames': int(ref_rows[:, 0].max()) + 1 if len(ref_rows) else 0, 'rows': int(len(ref_rows)), 'rows_gpu': int(len(gpu_rows))}
    if len(gpu_rows) != len(ref_rows):
        out['ok'] = False
        return out
    ids_ok = bool((gpu_rows['frame'] == ref_rows[:, 0]).all() and (gpu_rows['track_id'] == ref_rows[:, 1]).all())
    geo = np.stack([gpu_rows['w'], gpu_rows['h'], gpu_rows['deg']], 1).astype(np.float64)
    gerr = np.abs(geo - ref_rows[:, 4:7].astype(np.float32)).max(1)
    flips = int((gerr > 1e-3).sum())
    err = np.maximum(np.abs(gpu_rows['x'] - ref_rows[:, 2]) / np.maximum(1, np.abs(ref_rows[:, 2])),
                     np.abs(gpu_rows['y'] - ref_rows[:, 3]) / np.maximum(1, np.abs(ref_rows[:, 3])))
    exact = (gpu_rows['x'] == ref_rows[:, 2]) & (gpu_rows['y'] == ref_rows[:, 3])
    out.update({'ids_bit_exact': ids_ok, 'rect_max_err_excl_ties': float(gerr[gerr <= 1e-3].max()) if (gerr <= 1e-3).any() else 0.0,
                'rect_area_tie_flips': flips, 'xy_rel_err_max_all_rows': float(err.max()) if len(err) else 0.0,
                'xy_bit_exact_fraction': float(exact.mean()) if len(err) else 1.0})
    # a tie flip moves one measurement by a fraction of a pixel and stays in that track's filter history: allow for it
    bar = 1e-5 if flips == 0 else 1e-2
    out['ok'] = bool(ids_ok and flips <= max(2, len(ref_rows) // 5000) and out['xy_rel_err_max_all_rows'] < bar)
    return out


def workload_text(args, channels, frames_per_gpu, world):
    t = args.table
    h, w = scene_config(args, 1).height, scene_config(args, 1).width
    return (f'{t["what"]}; {w}x{h}x{channels} ({"BGR as cap.read() delivers" if channels == 3 else "grey plane"}), '
            f'{frames_per_gpu} frames per GPU x {world} GPU(s)' + (' of one video (strong scaling)' if args.strong else ''))


def port_calibration():
    """oracle/port_calibration.json (scripts/calibrate_port.py, build container): ms per frame of the real
    ysmr.tracker.CentroidTracker + GaussianSumFIR over the restated tracker the CPU legs time."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'oracle', 'port_calibration.json')) as fh:
            cal = json.load(fh)
        return {k: round(v['reference_over_port'], 3) for k, v in cal['cases'].items()}
    except Exception:
        return None


def ysmr_leg(args, frames, Cn, H, W):
    """Wall time of the drop-in on a video FILE: what a YSMR user sees (reference: 19.6 frames/s for the same 300-frame
    scene in the survey container, BASELINE.md section 2, of which FFV1 decode was 7.9 ms per frame)."""
    import shutil
    import tempfile

    import cv2
    from ysmr_b200.select import select_tracks
    from ysmr_b200.track_eval import track_bacteria
    n = min(args.ysmr_frames, frames.shape[0])
    grey = (frames[:n, :, :, 0] if Cn == 3 else frames[:n]).cpu().numpy()
    tmp = tempfile.mkdtemp(prefix='ysmr_b200_bench_')
    try:
        video = os.path.join(tmp, 'scene.avi')
        vw = cv2.VideoWriter(video, cv2.VideoWriter_fourcc(*'FFV1'), 30.0, (W, H), True)
        for f in grey:
            vw.write(np.repeat(f[..., None], 3, axis=-1))
        vw.release()
        st = detect_settings(args)
        settings = {'white bacteria on dark background': bool(st.white_on_dark), 'threshold offset for detection': int(st.offset),
                    'adaptive double threshold': float(st.adt), 'minimal frame count': 10, 'display video analysis': False,
                    'minimal length in seconds': 2.0, 'limit track length to x seconds': 2.0, 'store processed .csv file': True,
                    'frame height': H, 'frame width': W}
        out = {}
        for sink in ('append', 'once'):
            t0 = time.perf_counter()
            res = track_bacteria(video, dict(settings), tmp, row_sink=sink, max_blobs=args.table['max_blobs'], max_tracks=args.table['max_tracks'])
            t1 = time.perf_counter()
            if res is None:
                return {'error': 'track_bacteria returned None'}
            df, fps, fh, fw, csv_path = res
            sel = select_tracks(path_to_file=csv_path, df=df, results_directory=tmp, fps=fps, frame_height=fh, frame_width=fw,
                                settings=dict(settings))
            t2 = time.perf_counter()
            out[sink] = {'track_s': t1 - t0, 'select_s': t2 - t1, 'rows': int(len(df)), 'selected_rows': 0 if sel is None else int(len(sel))}
        best = min(out.values(), key=lambda v: v['track_s'] + v['select_s'])
        return {'value': n / (best['track_s'] + best['select_s']), 'unit': 'frames/s', 'frames': n, 'container': 'FFV1 AVI (lossless, intra-only)',
                'what': 'track_bacteria (cv2 decode on reader threads, GPU detect + link, sorted CSV) + select_tracks (GPU), wall clock',
                'row_sink': out, 'host_cpus': os.cpu_count()}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_reference(args):
    """--impl reference: rank 0 only; each step is a bounded sample (args.cpu_frames frames) of the same workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from ysmr_b200.synth import render_frames, to_bgr
    n = args.cpu_frames
    scene = scene_for(args, n)
    grey = render_frames(scene, 0, n)
    frames = to_bgr(grey) if args.channels == 3 else grey
    st = detect_settings(args)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_fps(frames[:min(30, n)], st)
    vals, secs = [], []
    for _ in range(args.steps):
        fps, dt, _, threads = cpu_reference_fps(frames, st)
        vals.append(fps); secs.append(dt)
    v = float(np.mean(vals))
    cfg = scene_config(args, 1)
    line = {
        'impl': 'reference', 'metric': f'frames/sec detect+link at {cfg.width}x{cfg.height}', 'value': v, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(np.mean(secs) * 1000),
        'higher_is_better': True, 'scaling': 'strong' if args.strong else 'weak', 'vs_baseline': None, 'dtype': 'u8/f32/f64',
        'data': 'synthetic',
        'config': {'workload': workload_text(args, args.channels, args.frames, int(os.environ.get('WORLD_SIZE', '1'))) +
                               f'; CPU sample = first {n} frames per step (numpy renderer, same scene), frames in RAM'},
        'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{n} frames/step x {args.steps} steps; oracle = the reference loop body '
                                   f'(track_eval.py:180-316) on cv2/scipy + restated CentroidTracker/GSFF '
                                   f'(speed of the restated tracker against the real class: port_vs_reference_tracker); '
                                   f'host has {os.cpu_count()} cpus',
                         'port_vs_reference_tracker': port_calibration()},
        'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    import hashlib

    import torch
    import torch.distributed as dist
    from ysmr_b200.api import ROW_DTYPE, Context
    from ysmr_b200.synth import render_frames_torch

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not args.batch:
        args.batch = args.table['batch1'] if world == 1 else min(592, args.table['batch1'])
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    scfg = scene_config(args, 1)
    H, W, Cn, F = scfg.height, scfg.width, args.channels, args.frames
    if F * H * W * Cn > 170e9:
        raise SystemExit(f'{args.config} with {F} frames of {H}x{W}x{Cn} per GPU does not fit one B200 (use --channels 1 or more GPUs)')
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s (B200_PROFILING.md)'

    # ---- synthetic video of world*F frames.  N == 1 or --multi gather: this rank keeps the contiguous range
    # [rank*F, (rank+1)*F).  --multi stream (default for N > 1): the video is cut into chunks of `batch` frames and chunk c
    # belongs to rank c % N (frame-range sharding at chunk granularity), so that the one sequential linker on rank 0 can
    # consume chunks in frame order while all ranks are still detecting.  Every rank holds ~F frames either way.
    scene = scene_for(args, world * F)
    stream_mode = world > 1 and args.multi == 'stream'
    B = args.batch
    if stream_mode:
        n_chunks = (world * F + B - 1) // B
        chunk_range = lambda c: (c * B, min(world * F, (c + 1) * B))
        my_chunks = [c for c in range(n_chunks) if c % world == rank]
        spans = [chunk_range(c) for c in my_chunks]
    else:
        spans = [(rank * F, (rank + 1) * F)]
    n_local = sum(b - a for a, b in spans)
    shape = (n_local, H, W) if Cn == 1 else (n_local, H, W, 3)
    frames = torch.empty(shape, dtype=torch.uint8, device=dev)
    G = max(8, min(200, int(4e8 // (scfg.n_cells * 17 * 17 * 12 + H * W * 8))))   # frames per rendering call (memory bound)
    pos, span_pos = 0, []
    for a, b in spans:
        span_pos.append(pos)
        for x in range(a, b, G):
            y = min(b, x + G)
            render_frames_torch(scene, x, y, dev, channels=Cn, out=frames[pos:pos + (y - x)])
            pos += y - x
    torch.cuda.synchronize()

    MB, MT = args.table['max_blobs'], args.table['max_tracks']
    ctx_kw = dict(max_blobs=MB, max_tracks=MT, white_on_dark=args.white)
    if args.config == 'cfg3':
        ctx_kw['max_runs'] = 65536
    ctx = Context(H, W, Cn, local, max_batch=args.batch, **ctx_kw)
    rows_per_frame = max(160, int(1.25 * scfg.n_cells))
    rows_cap = world * F * rows_per_frame
    rows_buf = torch.empty(rows_cap * ROW_DTYPE.itemsize, dtype=torch.uint8, device=dev) if rank == 0 else None

    def step_single():
        ctx.reset()
        return ctx.track_device(frames, 0, rows_capacity=rows_cap, rows_buf=rows_buf, return_device=True)

    # multi-GPU: every rank detects its own frame range; rank 0 runs the whole pipeline on its range (linker
    # overlapped with detection), the other ranks' detection records (count + 5 floats per blob) are gathered over NCCL, and
    # rank 0's one sequential linker continues through them in frame order.  Linking cannot be sharded exactly.
    import ctypes as C

    def detect_all():
        counts = torch.empty(F, dtype=torch.int32, device=dev)
        blobs = torch.empty((F, MB, 5), dtype=torch.float32, device=dev)
        for i in range(0, F, args.batch):
            j = min(F, i + args.batch)
            c, bl = ctx.detect(frames[i:j], rank * F + i)
            counts[i:j] = c; blobs[i:j] = bl
        return counts, blobs

    def step_multi():
        if rank == 0:
            ctx.reset()
            _, n0 = ctx.track_device(frames, 0, rows_capacity=rows_cap, rows_buf=rows_buf, return_device=True)
            counts = torch.zeros(F, dtype=torch.int32, device=dev)
            blobs = torch.zeros((F, MB, 5), dtype=torch.float32, device=dev)
            cs = [torch.empty_like(counts) for _ in range(world)]
            bs = [torch.empty_like(blobs) for _ in range(world)]
        else:
            counts, blobs = detect_all()
            cs = bs = None
        dist.gather(counts, cs, dst=0)
        dist.gather(blobs, bs, dst=0)
        if rank != 0:
            return None, None
        total = n0.clone()
        nr = torch.zeros(1, dtype=torch.int64, device=dev)
        for r in range(1, world):
            off = int(total.item()) * ROW_DTYPE.itemsize
            ctx._check(ctx.lib.ysmr_link(ctx._h, C.c_void_p(cs[r].data_ptr()), C.c_void_p(bs[r].data_ptr()), r * F, F,
                                         C.c_void_p(rows_buf.data_ptr() + off), rows_cap - int(total.item()),
                                         C.c_void_p(nr.data_ptr()), ctx._stream_ptr()))
            total += nr
        return rows_buf, total

    # streamed hand-over: one byte buffer per chunk = int32 counts[nf] followed by float32 blobs[nf][MB][5]
    if stream_mode:
        rec_bytes = 4 + MB * 5 * 4

        def chunk_views(buf, nf):
            counts = buf[:nf * 4].view(torch.int32)
            blobs = buf[nf * 4:nf * rec_bytes].view(torch.float32).view(nf, MB, 5)
            return counts, blobs

        chunk_buf = {}
        if rank == 0:
            for c in range(n_chunks):
                a, b = chunk_range(c)
                chunk_buf[c] = torch.empty(((b - a) * rec_bytes + 15) // 16 * 16, dtype=torch.uint8, device=dev)
        else:
            for c in my_chunks:
                a, b = chunk_range(c)
                chunk_buf[c] = torch.empty(((b - a) * rec_bytes + 15) // 16 * 16, dtype=torch.uint8, device=dev)
        # the serial linker goes first whenever it is ready: highest stream priority, like the library's own pipeline
        det_stream, link_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)
        n_rows_run = torch.zeros(1, dtype=torch.int64, device=dev)

        def detect_chunk(i, c, stream):
            a, b = chunk_range(c)
            counts, blobs = chunk_views(chunk_buf[c], b - a)
            fr = frames[span_pos[i]:span_pos[i] + (b - a)]
            ctx._check(ctx.lib.ysmr_detect(ctx._h, C.c_void_p(fr.data_ptr()), b - a, int(np.prod(shape[1:])), a,
                                           C.c_void_p(counts.data_ptr()), C.c_void_p(blobs.data_ptr()), None,
                                           C.c_void_p(stream.cuda_stream)))

        def step_stream():
            cur = torch.cuda.current_stream(dev)
            if rank != 0:
                reqs = []
                for i, c in enumerate(my_chunks):
                    detect_chunk(i, c, cur)
                    reqs.append(dist.isend(chunk_buf[c], dst=0))
                for r in reqs:
                    r.wait()
                return None, None
            ctx.reset()
            n_rows_run.zero_()
            det_stream.wait_stream(cur); link_stream.wait_stream(cur)
            reqs = {c: dist.irecv(chunk_buf[c], src=c % world) for c in range(n_chunks) if c % world != 0}
            done = {}
            with torch.cuda.stream(det_stream):
                for i, c in enumerate(my_chunks):
                    detect_chunk(i, c, det_stream)
                    done[c] = det_stream.record_event()
            with torch.cuda.stream(link_stream):
                for c in range(n_chunks):
                    a, b = chunk_range(c)
                    if c in done:
                        link_stream.wait_event(done[c])
                    else:
                        reqs[c].wait()                     # NCCL: makes the current (link) stream wait, not the host
                    counts, blobs = chunk_views(chunk_buf[c], b - a)
                    ctx._check(ctx.lib.ysmr_link_append(ctx._h, C.c_void_p(counts.data_ptr()), C.c_void_p(blobs.data_ptr()), a, b - a,
                                                        C.c_void_p(rows_buf.data_ptr()), rows_cap, C.c_void_p(n_rows_run.data_ptr()),
                                                        C.c_void_p(link_stream.cuda_stream)))
            cur.wait_stream(link_stream); cur.wait_stream(det_stream)
            return rows_buf, n_rows_run

    step = step_single if world == 1 else (step_stream if stream_mode else step_multi)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # ---- timed region: no per-launch profiling (the per-kernel split comes from a separate pass below)
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark(0)
    ev0.record()
    for _ in range(args.steps):
        rows_dev, n_rows_dev = step()
    ev1.record()
    barrier()
    sampler.mark(1)
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    n_rows = int(n_rows_dev.item()) if rank == 0 else 0
    rows = rows_dev[:n_rows * ROW_DTYPE.itemsize].cpu().numpy().view(ROW_DTYPE).copy() if rank == 0 else None

    # ---- second, untimed pass with CUDA events around every launch: per-kernel times for the roofline block
    ctx.set_profiling(True)
    barrier()
    step()
    barrier()
    prof = ctx.get_profile()
    ctx.set_profiling(False)
    prof_gen3 = None
    if world == 1:
        # the three-kernel front-end of round 1 on the same frames, for the A/B figure
        ctx.set_option(ctx.OPT_FRONTEND_GEN, 3)
        step(); torch.cuda.synchronize()
        ctx.set_profiling(True)
        step(); torch.cuda.synchronize()
        prof_gen3 = ctx.get_profile()
        ctx.set_profiling(False)
        ctx.set_option(ctx.OPT_FRONTEND_GEN, 4)

    # ---- multi-GPU identity: the streamed / gathered rows must be byte-identical to ONE sequential ysmr_link over the same
    # detection records (different chunking of the linker, no streams, no NCCL in between)
    multi_check = None
    if stream_mode and rank == 0:
        counts_all = torch.cat([chunk_views(chunk_buf[c], chunk_range(c)[1] - chunk_range(c)[0])[0] for c in range(n_chunks)])
        blobs_all = torch.cat([chunk_views(chunk_buf[c], chunk_range(c)[1] - chunk_range(c)[0])[1] for c in range(n_chunks)])
        ctx.reset()
        rows_seq = ctx.link(counts_all.contiguous(), blobs_all.contiguous(), 0, rows_capacity=rows_cap)
        multi_check = {'rows_sha256': hashlib.sha256(rows.tobytes()).hexdigest(),
                       'sequential_link_sha256': hashlib.sha256(rows_seq.tobytes()).hexdigest()}
        multi_check['identical'] = multi_check['rows_sha256'] == multi_check['sequential_link_sha256']
        del counts_all, blobs_all
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    total_frames = world * F * args.steps
    value = total_frames / (ms / 1000.0)
    n_tracks = int(rows['track_id'].max()) + 1 if n_rows else 0

    # ---- roofline: algorithmic bytes = H*W*C per frame (input read once) against the front-end launch that
    # consumes them -- fused_front_kernel (fused.cu), the dominant kernel -- timed with CUDA events on the detection stream.
    # Its own DRAM traffic comes from the committed ncu capture (profiles/r2_traffic.json) when present.
    fe_ms, fe_n = prof['frontend']
    bytes_per_frame = H * W * Cn
    frames_per_launch = n_local / max(fe_n, 1)
    per_launch_ms = fe_ms / max(fe_n, 1)
    achieved = (bytes_per_frame * frames_per_launch) / (per_launch_ms / 1000.0) / 1e9 if fe_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')))
        traffic = float(tj['frontend_dram_bytes_per_frame'][args.config]) * frames_per_launch
    except Exception:
        pass
    det_ms = prof['frontend'][0] + prof['label'][0] + prof['geometry'][0]
    roofline = {'bound': 'hbm', 'kernel': 'fused_front_kernel (BGR -> grey -> blur -> adaptive double threshold -> bit masks, one launch per batch)',
                'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak, 'traffic': traffic,
                'peak_source': peak_src, 'algorithmic_bytes_per_frame': bytes_per_frame, 'frames_per_launch': frames_per_launch,
                'avg_launch_ms': per_launch_ms,
                'kernel_ms_per_step': {k: v[0] for k, v in prof.items() if v[1]},
                'share_of_detection': prof['frontend'][0] / det_ms if det_ms > 0 else None,
                'linker_us_per_frame': 1000.0 * prof['link'][0] / (world * F) if prof['link'][1] else None,
                'whole_path_frac': (value * bytes_per_frame / 1e9) / (hbm_peak * world),
                'note': 'per-kernel times from a separate profiled pass (CUDA events around every launch); value from the unprofiled timed region'}
    if prof_gen3 is not None:
        g3 = prof_gen3['frontend'][0]
        roofline['three_kernel_frontend_of_round_1'] = {'ms_per_step': g3, 'frac': (bytes_per_frame * F / (g3 / 1000.0) / 1e9) / hbm_peak if g3 > 0 else None}

    # ---- e2e: the same metric through ysmr_track_host with pinned HOST frames (H2D + kernels + D2H rows inside) ---------
    e2e = e2e_grey = None
    if not args.no_e2e and world == 1:
        def e2e_leg(channels):
            E = min(args.e2e_frames, F)
            shp = (E, H, W) if channels == 1 else (E, H, W, 3)
            host = torch.empty(shp, dtype=torch.uint8).pin_memory()
            host.copy_(frames[:E] if channels == Cn else (frames[:E, :, :, 0] if Cn == 3 else frames[:E][..., None].expand(-1, -1, -1, 3)))
            host_np = host.numpy()
            rows_out = np.empty(E * rows_per_frame, ROW_DTYPE)
            calls = (F + E - 1) // E
            # host streaming uses the chunk size of the drop-in (ysmr_b200/track_eval.py): small chunks keep the H2D copy of
            # chunk i+1 under the kernels of chunk i and the exposed tail short
            ctx_e = Context(H, W, channels, local, max_batch=min(256, args.batch), **ctx_kw)

            def e2e_step():
                ctx_e.reset()
                got = 0
                for c in range(calls):
                    got += len(ctx_e.track_host(host_np, c * E, rows_capacity=len(rows_out), rows_out=rows_out))
                return got
            e2e_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            got = 0
            for _ in range(args.steps):
                got = e2e_step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            ctx_e.close()
            return {'value': calls * E * args.steps / dt, 'unit': 'frames/s',
                    'h2d_bytes_per_step': calls * E * H * W * channels, 'd2h_bytes_per_step': got * ROW_DTYPE.itemsize,
                    'note': f'ysmr_track_host on a pinned {E}-frame buffer of {channels}-channel frames x {calls} calls per step; '
                            f'wall clock incl. H2D, kernels, D2H of rows'}
        e2e = e2e_leg(Cn)
        if Cn == 3:
            # grey containers: the drop-in reads the decoder's single plane (cv2.CAP_PROP_CONVERT_RGB off) and opens the
            # context with channels=1, so a third of the bytes cross PCIe (cv2.cvtColor is the identity when B == G == R)
            e2e_grey = e2e_leg(1)

    # ---- e2e through the drop-in: ysmr_b200.track_eval.track_bacteria + select_tracks on a lossless FFV1 AVI of the scene
    # (decode by cv2.VideoCapture reader threads -> pinned buffers -> ysmr_track_host -> sorted CSV -> GPU selection)
    e2e_ysmr = None
    if not args.no_e2e and world == 1 and args.ysmr_frames > 0:
        try:
            e2e_ysmr = ysmr_leg(args, frames, Cn, H, W)
        except Exception as ex:          # the leg is a report beside the metric, never a reason to lose the line
            e2e_ysmr = {'error': repr(ex)}

    # ---- CPU baseline on a bounded sample of the same bytes, and the parity check of the timed bytes ----------------------
    cpu = parity = None
    if not args.no_cpu and world == 1:
        S = min(args.cpu_frames, F)
        sample = frames[:S].cpu().numpy()
        fps_cpu, dt_cpu, ref_rows, threads = cpu_reference_fps(sample, detect_settings(args), keep_rows=True)
        cpu = {'value': fps_cpu, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
               'sample': f'first {S} frames of the same video, {dt_cpu:.1f} s; reference loop body (track_eval.py:180-316) '
                         f'replayed on cv2/scipy + restated CentroidTracker/GSFF; host has {os.cpu_count()} cpus',
               'port_vs_reference_tracker': port_calibration()}
        parity = parity_check(rows[rows['frame'] < S], ref_rows)

    line = {
        'metric': f'frames/sec detect+link at {W}x{H}', 'value': value, 'unit': 'frames/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'strong' if args.strong else 'weak', 'vs_baseline': None, 'dtype': 'u8/f32/f64', 'data': 'synthetic',
        'config': {'workload': workload_text(args, Cn, F, world) +
                               f'; tracking.ini defaults (offset 5, adaptive double threshold 2.0, gsff 10/20/30), '
                               f'{"white on dark" if args.white else "dark on light"}',
                   'name': args.config, 'frames_per_gpu': F, 'batch': args.batch,
                   'l2': f'inputs ({F * bytes_per_frame / 1e9:.1f} GB per GPU) far exceed the 126 MB L2',
                   'parallelism': (f'frame-range x{world}, one sequential linker' if not stream_mode else
                                   f'chunk-interleaved frame ranges x{world} ({B}-frame chunks, chunk c on rank c % {world}), '
                                   f'records streamed over NCCL to the one sequential linker on rank 0'),
                   'rows': n_rows, 'tracks': n_tracks, 'max_blobs': MB, 'max_tracks': MT},
        'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'e2e_ysmr': e2e_ysmr, 'gpu_launches': int(launches), 'clocks': clocks,
        'parity_check': parity,
    }
    if e2e_grey is not None:
        line['e2e_grey'] = e2e_grey
    if multi_check is not None:
        line['multi_gpu_check'] = multi_check
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
