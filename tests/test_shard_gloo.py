"""Multi-process host logic of the sharded path (ysmr_b200/shard.py) on CPU: world_size 2 and 3, gloo backend, with the
oracle standing in for the two GPU steps.  The sharded result must equal the single-process result row for row."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ysmr_b200.shard import frame_ranges, read_range


def test_frame_ranges_cover_exactly():
    for n in (0, 1, 7, 300, 54000):
        for w in (1, 2, 3, 4, 8):
            r = frame_ranges(n, w)
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    assert read_range(100, 200, 0) == (100, 0) and read_range(100, 200, 150) == (0, 100) and read_range(400, 500, 150) == (250, 150)


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _frames():
    from ysmr_b200.synth import SceneConfig, make_scene, render_frames
    return render_frames(make_scene(SceneConfig(width=200, height=160, n_frames=45, n_cells=8, seed=6, margin=30.0)))


def _settings(mode):
    from oracle import ref_stages
    return ref_stages.DetectSettings(True, 25, -1.0, fps=2.0) if mode == 'meanstd' else ref_stages.DetectSettings()


def _oracle_detect(grey, mode, read_start, stop, max_blobs=64):
    """oracle detection of frames [read_start, stop) as the C-ABI would return it (counts, dense blobs)"""
    from oracle import ref_stages
    st = _settings(mode)
    counts, blobs = [], []
    for t in range(read_start, stop):
        a = ref_stages.rects_to_array(ref_stages.detect_frame(grey[t], st)['rects'])
        counts.append(len(a))
        dense = np.zeros((max_blobs, 5), np.float32); dense[:len(a)] = a
        blobs.append(dense)
    return torch.tensor(counts, dtype=torch.int32), torch.from_numpy(np.stack(blobs) if blobs else np.zeros((0, max_blobs, 5), np.float32))


class _OracleLinker:
    def __init__(self):
        from oracle.tracker_port import LinkerPort
        self.lp = LinkerPort(max_disappeared=30.0, fps=30.0)

    def __call__(self, counts, blobs, first):
        rows = []
        for i in range(counts.numel()):
            rec = blobs[i, :int(counts[i])].numpy()
            rects = [((float(r[0]), float(r[1])), (float(r[2]), float(r[3]), float(r[4]))) for r in rec]
            rows += [(first + i, k, xy[0], xy[1], info[0], info[1], info[2]) for (k, xy, info) in self.lp.update(rects)]
        return np.array(rows, np.float64).reshape(-1, 7)


def _worker(rank, world, port, mode, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ysmr_b200.shard import track_sharded
    grey = _frames()
    lead = 10 if mode == 'meanstd' else 0          # window = floor(5 * 2.0) + 1 = 11 frames -> 10 earlier ones
    linker = _OracleLinker()
    rows = track_sharded(len(grey), world, rank, lambda a, b: _oracle_detect(grey, mode, a, b), linker, lead_in=lead,
                         dist=dist, device=torch.device('cpu'))
    if rank == 0:
        np.save(out_path, np.concatenate(rows))
    dist.barrier()
    dist.destroy_process_group()


def _worker_streamed(rank, world, port, mode, out_path):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ysmr_b200.shard import track_streamed
    grey = _frames()
    lead = 10 if mode == 'meanstd' else 0
    linker = _OracleLinker()
    rows = track_streamed(len(grey), 7, world, rank, lambda a, b: _oracle_detect(grey, mode, a, b), linker, lead_in=lead,
                          dist=dist, device=torch.device('cpu'), max_blobs=64)
    if rank == 0:
        np.save(out_path, np.concatenate(rows))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,mode', [(2, 'adaptive'), (3, 'adaptive'), (2, 'meanstd')])
def test_streamed_chunks_equal_single_process(tmp_path, world, mode):
    """chunk-interleaved sharding (chunk c on rank c % world, 7-frame chunks, last one ragged) with the streamed hand-over"""
    out = str(tmp_path / 'rows.npy')
    mp.spawn(_worker_streamed, args=(world, _free_port(), mode, out), nprocs=world, join=True)
    got = np.load(out)
    grey = _frames()
    c, b = _oracle_detect(grey, mode, 0, len(grey))
    ref = _OracleLinker()(c, b, 0)
    assert got.shape == ref.shape and (got == ref).all()


def test_chunk_spans_and_owner():
    from ysmr_b200.shard import chunk_owner, chunk_spans
    sp = chunk_spans(45, 7)
    assert sp[0] == (0, 7) and sp[-1] == (42, 45) and all(a[1] == b[0] for a, b in zip(sp, sp[1:]))
    assert [chunk_owner(c, 3) for c in range(7)] == [0, 1, 2, 0, 1, 2, 0]
    assert chunk_spans(0, 7) == []


@pytest.mark.parametrize('world,mode', [(2, 'adaptive'), (3, 'adaptive'), (2, 'meanstd')])
def test_sharded_equals_single_process(tmp_path, world, mode):
    out = str(tmp_path / 'rows.npy')
    mp.spawn(_worker, args=(world, _free_port(), mode, out), nprocs=world, join=True)
    got = np.load(out)
    grey = _frames()
    c, b = _oracle_detect(grey, mode, 0, len(grey))
    ref = _OracleLinker()(c, b, 0)
    assert got.shape == ref.shape and (got == ref).all()
