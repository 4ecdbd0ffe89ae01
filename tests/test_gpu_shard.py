"""-m gpu: the sharded multi-GPU path against the single-GPU path (needs >= 2 GPUs; skipped otherwise).  One process per
GPU, NCCL; rows must be bit-identical to the single-GPU rows."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close()
    return p


def _frames():
    from ysmr_b200.synth import SceneConfig, make_scene, render_frames
    return render_frames(make_scene(SceneConfig(width=320, height=240, n_frames=150, n_cells=14, seed=3, margin=30.0)))


def _run_single(grey):
    from ysmr_b200.api import Context
    ctx = Context(240, 320, 1, 0, max_batch=32, max_blobs=256, max_tracks=512)
    rows = ctx.track_device(torch.from_numpy(grey).cuda(), 0)
    ctx.close()
    return rows


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    from ysmr_b200.api import Context
    from ysmr_b200.shard import track_sharded
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    grey = _frames()
    ctx = Context(240, 320, 1, rank, max_batch=32, max_blobs=256, max_tracks=512)

    def detect_range(a, b):
        fr = torch.from_numpy(grey[a:b]).to(dev)
        cs, bs = [], []
        for i in range(0, b - a, 32):
            c, bl = ctx.detect(fr[i:i + 32], a + i)
            cs.append(c); bs.append(bl)
        return torch.cat(cs), torch.cat(bs)

    link_ctx = {}

    def link_range(c, b, first):
        if 'l' not in link_ctx:
            link_ctx['l'] = Context(240, 320, 1, rank, max_batch=4, max_blobs=int(b.shape[1]), max_tracks=512)
        return link_ctx['l'].link(c, b, first, rows_capacity=int(c.numel()) * 512)

    rows = track_sharded(len(grey), world, rank, detect_range, link_range, 0, dist, dev)
    if rank == 0:
        np.save(out_path, np.concatenate(rows))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_rows_bit_identical_to_one_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    out = str(tmp_path / 'rows.npy')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    ref = _run_single(_frames())
    assert got.tobytes() == ref.tobytes()


def _worker_streamed(rank, world, port, out_path):
    import torch.distributed as dist
    from ysmr_b200.api import Context
    from ysmr_b200.shard import track_streamed
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    grey = _frames()
    ctx = Context(240, 320, 1, rank, max_batch=32, max_blobs=256, max_tracks=512)

    def detect_range(a, b):
        return ctx.detect(torch.from_numpy(grey[a:b]).to(dev), a)

    def link_range(c, b, first):
        return ctx.link(c, b, first, rows_capacity=int(c.numel()) * 512)

    rows = track_streamed(len(grey), 32, world, rank, detect_range, link_range, 0, dist, dev, max_blobs=256)
    if rank == 0:
        np.save(out_path, np.concatenate(rows))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_streamed_chunks_bit_identical_to_one_gpu(tmp_path):
    """chunk-interleaved sharding with the point-to-point hand-over (shard.track_streamed) over NCCL"""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    out = str(tmp_path / 'rows.npy')
    mp.spawn(_worker_streamed, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    ref = _run_single(_frames())
    assert got.tobytes() == ref.tobytes()


def test_streamed_single_gpu_equals_pipeline():
    """world == 1: the chunked detect/link calls of track_streamed equal the one-call pipeline"""
    from ysmr_b200.api import Context
    from ysmr_b200.shard import track_streamed
    grey = _frames()
    ctx = Context(240, 320, 1, 0, max_batch=32, max_blobs=256, max_tracks=512)
    dev = torch.device('cuda', 0)
    rows = track_streamed(len(grey), 32, 1, 0, lambda a, b: ctx.detect(torch.from_numpy(grey[a:b]).to(dev), a),
                          lambda c, b, first: ctx.link(c, b, first, rows_capacity=int(c.numel()) * 512))
    ctx.close()
    assert np.concatenate(rows).tobytes() == _run_single(grey).tobytes()


def test_two_devices_in_one_process():
    """Kernel attributes (shared-memory opt-in of the linker and the fused front-end) are per device: a context on device 1
    after one on device 0, in the same process, must work and give the same rows."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from ysmr_b200.api import Context
    from ysmr_b200.synth import SceneConfig, make_scene, render_frames
    cfg = SceneConfig(width=320, height=240, n_frames=40, n_cells=12, seed=3, margin=30.0)
    grey = render_frames(make_scene(cfg))
    rows = []
    for dev in (0, 1):
        ctx = Context(cfg.height, cfg.width, 1, dev, max_batch=16, max_blobs=512, max_tracks=512)
        rows.append(ctx.track_host(grey, 0))
        ctx.close()
    assert rows[0].tobytes() == rows[1].tobytes() and len(rows[0]) > 0
