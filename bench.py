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
    ap.add_argument('--frames', type=int, default=9000, help='frames per GPU (configs[1]: 9000)')
    ap.add_argument('--channels', type=int, default=3, choices=[1, 3])
    ap.add_argument('--cells', type=int, default=50)
    ap.add_argument('--batch', type=int, default=0,
                    help='frames per detect launch (0: 1184 on one GPU, 592 = one chunk of the streamed hand-over on several)')
    ap.add_argument('--cpu-frames', type=int, default=900, help='frames of the bounded CPU sample (about 15 s)')
    ap.add_argument('--e2e-frames', type=int, default=1000, help='frames in the pinned host buffer of the e2e leg')
    ap.add_argument('--multi', default='stream', choices=['stream', 'gather'],
                    help='N > 1: stream detection records chunk by chunk to the linker (default) or gather whole ranges')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    return ap.parse_args()


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
def cpu_reference_fps(frames_bgr_or_grey, fps=30.0):
    """Replays track_eval.py:180-316 on in-RAM frames (oracle/; the reference itself cannot travel to the GPU box)."""
    import cv2
    from oracle import ref_stages
    from oracle.tracker_port import LinkerPort
    st = ref_stages.DetectSettings()
    lp = LinkerPort(max_disappeared=fps, fps=fps)
    t0 = time.perf_counter()
    n_rows = 0
    for f in frames_bgr_or_grey:
        r = ref_stages.detect_frame(f, st)
        n_rows += len(lp.update(r['rects']))
    dt = time.perf_counter() - t0
    return len(frames_bgr_or_grey) / dt, dt, n_rows, cv2.getNumThreads()


def scene_for(args, n_total):
    from ysmr_b200.synth import SceneConfig, make_scene
    return make_scene(SceneConfig(n_frames=n_total, n_cells=args.cells, seed=0))


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
    for _ in range(min(args.warmup, 1)):
        cpu_reference_fps(frames[:30])
    vals, secs = [], []
    for _ in range(args.steps):
        fps, dt, _, threads = cpu_reference_fps(frames)
        vals.append(fps); secs.append(dt)
    v = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': 'frames/sec detect+link at 1228x922', 'value': v, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': float(np.mean(secs) * 1000),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/f32/f64', 'data': 'synthetic',
        'config': {'workload': f'cfg2 scene: 1228x922x{args.channels}, {args.cells} rods, default tracking.ini; '
                               f'CPU sample = first {n} frames per step, frames in RAM'},
        'cpu_baseline': {'value': v, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{n} frames/step x {args.steps} steps; oracle = the reference loop body '
                                   f'(track_eval.py:180-316) on cv2/scipy + restated CentroidTracker/GSFF; '
                                   f'host has {os.cpu_count()} cpus'},
        'e2e': {'value': v, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from ysmr_b200.api import ROW_DTYPE, Context
    from ysmr_b200.synth import render_frames_torch

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not args.batch:
        args.batch = 1184 if world == 1 else 592
    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    H, W, Cn, F = 922, 1228, args.channels, args.frames
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
    G = 200
    pos, span_pos = 0, []
    for a, b in spans:
        span_pos.append(pos)
        for x in range(a, b, G):
            y = min(b, x + G)
            render_frames_torch(scene, x, y, dev, channels=Cn, out=frames[pos:pos + (y - x)])
            pos += y - x
    torch.cuda.synchronize()

    MB, MT = 128, 1024              # capacities: detections per frame (the scene has ~50), live tracks
    ctx = Context(H, W, Cn, local, max_batch=args.batch, max_blobs=MB, max_tracks=MT)
    rows_cap = world * F * 160
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
    ctx.set_profiling(True)
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
    prof = ctx.get_profile()
    ctx.set_profiling(False)
    launches = ctx.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    total_frames = world * F * args.steps
    value = total_frames / (ms / 1000.0)
    n_rows = int(n_rows_dev.item())
    rows = rows_dev[:n_rows * ROW_DTYPE.itemsize].cpu().numpy().view(ROW_DTYPE)
    n_tracks = int(rows['track_id'].max()) + 1 if n_rows else 0

    # ---- roofline: algorithmic bytes = H*W*C per frame (input read once), against the front-end launches that
    # consume them (K1a blur pre-pass, K1b Gaussian + decisions -- the dominant kernel --, K1c mask packing), timed with
    # CUDA events on the detection stream inside the timed region.  Per-kernel own DRAM traffic comes from the committed
    # ncu capture (profiles/r1_traffic.json) when present.
    fe_ms, fe_n = prof['frontend']
    bytes_per_frame = H * W * Cn
    frames_per_launch = (n_local * args.steps) / max(fe_n, 1)
    per_launch_ms = fe_ms / max(fe_n, 1)
    achieved = (bytes_per_frame * frames_per_launch) / (per_launch_ms / 1000.0) / 1e9 if fe_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')))
        traffic = float(tj['frontend_dram_bytes_per_frame']) * frames_per_launch
    except Exception:
        pass
    kb = prof['k1b']
    k1b_ms = kb[0] / max(kb[1], 1)
    roofline = {'bound': 'hbm', 'kernel': 'front-end launches of one batch: K1a blur_prepass + margins, K1b gauss_decide (dominant), K1c pack_masks',
                'achieved': achieved, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak, 'traffic': traffic,
                'peak_source': peak_src, 'algorithmic_bytes_per_frame': bytes_per_frame, 'frames_per_launch': frames_per_launch,
                'avg_launch_ms': per_launch_ms,
                'dominant_kernel': {'name': 'gauss_decide_kernel (K1b)', 'avg_launch_ms': k1b_ms,
                                    'share_of_frontend': kb[0] / fe_ms if fe_ms > 0 else None,
                                    'note': 'issue-bound (FP32 + integer pipes), reads 1 B/px and writes 0.25 B/px; see DESIGN.md section 4'},
                'kernel_ms_per_step': {k: v[0] / args.steps for k, v in prof.items()},
                'whole_path_frac': (value * bytes_per_frame / 1e9) / (hbm_peak * world)}

    # ---- e2e: the same metric through ysmr_track_host with pinned HOST frames (H2D + kernels + D2H rows inside) ---------
    e2e = None
    if not args.no_e2e and world == 1:
        E = min(args.e2e_frames, F)
        host = torch.empty((E,) + shape[1:], dtype=torch.uint8).pin_memory()
        host.copy_(frames[:E])
        host_np = host.numpy()
        rows_out = np.empty(E * 160, ROW_DTYPE)
        calls = (F + E - 1) // E
        # host streaming uses the chunk size of the drop-in (ysmr_b200/track_eval.py): small chunks keep the H2D copy of chunk
        # i+1 under the kernels of chunk i and the exposed tail short
        ctx_e = Context(H, W, Cn, local, max_batch=256, max_blobs=MB, max_tracks=MT)
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
        e2e = {'value': calls * E * args.steps / dt, 'unit': 'frames/s',
               'h2d_bytes_per_step': calls * E * bytes_per_frame, 'd2h_bytes_per_step': got * ROW_DTYPE.itemsize,
               'note': f'ysmr_track_host on a pinned {E}-frame buffer x {calls} calls per step; wall clock incl. H2D, '
                       f'kernels, D2H of rows'}

    # ---- CPU baseline on a bounded sample of the same bytes ---------------------------------------------------------------
    cpu = None
    if not args.no_cpu and world == 1:
        S = min(args.cpu_frames, F)
        sample = frames[:S].cpu().numpy()
        fps_cpu, dt_cpu, _, threads = cpu_reference_fps(sample)
        cpu = {'value': fps_cpu, 'unit': 'frames/s', 'cores': threads, 'kind': 'port',
               'sample': f'first {S} frames of the same video, {dt_cpu:.1f} s; reference loop body (track_eval.py:180-316) '
                         f'replayed on cv2/scipy + restated CentroidTracker/GSFF; host has {os.cpu_count()} cpus'}

    line = {
        'metric': 'frames/sec detect+link at 1228x922', 'value': value, 'unit': 'frames/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/f32/f64', 'data': 'synthetic',
        'config': {'workload': f'cfg2: 1228x922x{Cn} (BGR as cap.read() delivers) x {F} frames per GPU, {args.cells} rods, '
                               f'default tracking.ini (white-on-dark, offset 5, adaptive double threshold 2.0, gsff 10/20/30)',
                   'frames_per_gpu': F, 'batch': args.batch, 'l2': 'inputs (>= 10 GB) far exceed the 126 MB L2',
                   'parallelism': (f'frame-range x{world}, one sequential linker' if not stream_mode else
                                   f'chunk-interleaved frame ranges x{world} ({B}-frame chunks, chunk c on rank c % {world}), '
                                   f'records streamed over NCCL to the one sequential linker on rank 0'),
                   'rows': n_rows, 'tracks': n_tracks,
                   'max_blobs': MB, 'max_tracks': MT},
        'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
