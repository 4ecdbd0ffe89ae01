"""Multi-GPU: frame-range sharding of detection, one sequential linker.

Detection of a frame depends on no other frame (adaptive modes) or only on per-frame scalars of the preceding
``window - 1`` frames (mean/std mode), so a long video splits into contiguous frame ranges, one per GPU.  Linking does
not split: track ids, disappearance counters and the GSFF recursion make it strictly sequential, and an
"overlap and stitch" scheme cannot reproduce the reference's ids bit for bit (SURVEY finding 4).  So the ranges'
detection records (a few KB per frame) are gathered over NCCL to rank 0, which runs the one linker over them in frame
order.  The only collectives are that gather and one max-reduction for the record width.

The two compute steps are passed in as callables so that the host logic (range arithmetic, lead-in frames, gather
order, hand-over to the linker) can be tested on CPU with the gloo backend and the oracle as the compute.
"""
from __future__ import annotations

import numpy as np

__all__ = ['frame_ranges', 'read_range', 'gather_detections', 'track_sharded', 'chunk_owner', 'chunk_spans', 'track_streamed']


def frame_ranges(n_frames: int, world: int):
    """Contiguous ranges [start, stop) per rank; the first ``n_frames % world`` ranks get one frame more."""
    base, extra = divmod(n_frames, world)
    out, start = [], 0
    for r in range(world):
        stop = start + base + (1 if r < extra else 0)
        out.append((start, stop))
        start = stop
    return out


def read_range(start: int, stop: int, lead_in: int):
    """Frames a rank has to run through detection: its own range plus ``lead_in`` preceding frames whose results are
    dropped (mean/std mode: the moving-average threshold of frame t needs the scalars of up to 5*fps earlier frames;
    adaptive modes use lead_in = 0).  Returns (read_start, n_dropped)."""
    read_start = max(0, start - lead_in)
    return read_start, start - read_start


def gather_detections(counts, blobs, world, rank, dist=None, device=None):
    """counts: int32 [n_local]; blobs: float32 [n_local, max_blobs, 5] (torch tensors on the rank's device).
    Returns on rank 0 a list of (counts_r, blobs_r) in rank order with blobs trimmed to the widest frame of the whole
    video; elsewhere None.  Ranges may differ by one frame, so the payload is padded to the longest range."""
    import torch
    if world == 1:
        k = int(counts.max().item()) if counts.numel() else 0
        return [(counts, blobs[:, :max(k, 1)].contiguous())]
    meta = torch.tensor([int(counts.max().item()) if counts.numel() else 0, counts.numel()], dtype=torch.int64, device=device)
    dist.all_reduce(meta, op=dist.ReduceOp.MAX)
    k, n_max = max(int(meta[0].item()), 1), int(meta[1].item())
    n_local = counts.numel()
    c_pad = torch.zeros(n_max + 1, dtype=torch.int32, device=device)
    c_pad[:n_local] = counts
    c_pad[n_max] = n_local
    b_pad = torch.zeros((n_max, k, 5), dtype=torch.float32, device=device)
    b_pad[:n_local] = blobs[:, :k]
    if rank == 0:
        cs = [torch.empty_like(c_pad) for _ in range(world)]
        bs = [torch.empty_like(b_pad) for _ in range(world)]
    else:
        cs = bs = None
    dist.gather(c_pad, cs, dst=0)
    dist.gather(b_pad, bs, dst=0)
    if rank != 0:
        return None
    out = []
    for c, b in zip(cs, bs):
        n = int(c[n_max].item())
        out.append((c[:n].contiguous(), b[:n].contiguous()))
    return out


def track_sharded(n_frames, world, rank, detect_range, link_range, lead_in=0, dist=None, device=None):
    """detect_range(read_start, stop) -> (counts, blobs) for frames [read_start, stop) of the video (this rank only);
    link_range(counts, blobs, first_frame) -> rows for those frames, called on rank 0 once per range in frame order
    (the callee keeps the linker state between calls).  Returns the list of row arrays on rank 0, None elsewhere."""
    start, stop = frame_ranges(n_frames, world)[rank]
    read_start, dropped = read_range(start, stop, lead_in)
    counts, blobs = detect_range(read_start, stop)
    counts, blobs = counts[dropped:], blobs[dropped:]
    parts = gather_detections(counts, blobs, world, rank, dist, device)
    if rank != 0:
        return None
    rows, first = [], 0
    for c, b in parts:
        rows.append(link_range(c, b, first))
        first += int(c.numel())
    assert first == n_frames
    return rows


# ---- chunk-interleaved sharding with a streamed hand-over --------------------------------------------------------------------
# Contiguous ranges make the one sequential linker wait for whole ranges.  Cutting the video into chunks of `chunk` frames and
# giving chunk c to rank c % world lets the linker (rank 0) consume chunks in frame order while every rank is still detecting:
# the job then runs at max(detection / world, linking) instead of detection + (world - 1) * linking.  Still frame-range
# sharding, still one linker, results identical to the single-GPU run.

def chunk_spans(n_frames: int, chunk: int):
    """[(start, stop)] of the chunks of a video, in frame order."""
    return [(a, min(n_frames, a + chunk)) for a in range(0, n_frames, chunk)]


def chunk_owner(c: int, world: int) -> int:
    return c % world


def track_streamed(n_frames, chunk, world, rank, detect_range, link_range, lead_in=0, dist=None, device=None, max_blobs=None):
    """detect_range(read_start, stop) -> (counts int32 [n], blobs float32 [n, max_blobs, 5]) for frames [read_start, stop);
    link_range(counts, blobs, first_frame) -> rows, called on rank 0 once per chunk in frame order.  Chunks detected by other
    ranks travel to rank 0 as one point-to-point message each (counts packed in front of the blobs as float32).  `lead_in`
    frames before a chunk are detected and dropped (mean/std mode).  Returns the list of row arrays on rank 0, else None."""
    import torch
    spans = chunk_spans(n_frames, chunk)
    rows = []
    if world == 1:
        for a, b in spans:
            rs, dropped = read_range(a, b, lead_in)
            c, bl = detect_range(rs, b)
            rows.append(link_range(c[dropped:], bl[dropped:], a))
        return rows

    def pack(c, bl):
        return torch.cat([c.to(torch.float32).reshape(-1, 1), bl.reshape(bl.shape[0], -1)], 1).contiguous()

    def unpack(buf, k):
        return buf[:, 0].to(torch.int32).contiguous(), buf[:, 1:].reshape(buf.shape[0], k, 5).contiguous()

    if rank != 0:
        reqs, keep = [], []
        for ci, (a, b) in enumerate(spans):
            if chunk_owner(ci, world) != rank:
                continue
            rs, dropped = read_range(a, b, lead_in)
            c, bl = detect_range(rs, b)
            buf = pack(c[dropped:], bl[dropped:])
            keep.append(buf)                           # must stay alive until the send has completed
            reqs.append(dist.isend(buf, dst=0))
        for r in reqs:
            r.wait()
        return None
    assert max_blobs is not None, 'rank 0 needs the record width to post its receives'
    # Receives are posted in frame order (so they match the senders' order per source) but only a bounded window ahead of the
    # chunk being linked: rank-0 memory stays at `window` record buffers however long the video is (a full-width buffer is
    # chunk x (1 + 5 max_blobs) floats: 21 MB at the defaults, almost all of it padding).
    window = max(4, 2 * world)
    remote = [ci for ci in range(len(spans)) if chunk_owner(ci, world) != 0]
    pending, nxt = {}, 0

    def post_until(limit):
        nonlocal nxt
        while nxt < len(remote) and remote[nxt] < limit:
            ci = remote[nxt]
            a, b = spans[ci]
            buf = torch.empty((b - a, 1 + max_blobs * 5), dtype=torch.float32, device=device)
            pending[ci] = (buf, dist.irecv(buf, src=chunk_owner(ci, world)))
            nxt += 1

    for ci, (a, b) in enumerate(spans):
        post_until(ci + window + 1)
        if ci in pending:
            buf, req = pending.pop(ci)
            req.wait()
            c, bl = unpack(buf, max_blobs)
        else:
            rs, dropped = read_range(a, b, lead_in)
            c, bl = detect_range(rs, b)
            c, bl = c[dropped:], bl[dropped:]
        rows.append(link_range(c, bl, a))
    return rows
