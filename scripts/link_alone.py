"""The linker's kernel time with nothing else on the GPU (detections of all frames computed first), from the library's own
per-launch CUDA events -- to hold against bench.py's `kernel_ms_per_step.link`, which is measured while the detection of
the next chunk saturates the other 147 SMs.   python scripts/link_alone.py [cfg] [frames]"""
import dataclasses
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ysmr_b200.api import Context  # noqa: E402
from ysmr_b200.synth import CONFIGS, make_scene, render_frames_torch  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
F = int(sys.argv[2]) if len(sys.argv) > 2 else 2368
cfg = dataclasses.replace(CONFIGS[name], n_frames=F)
scene = make_scene(cfg)
mb, mt = {'cfg2': (128, 1024), 'cfg4': (1024, 4096), 'cfg3': (4096, 8192)}[name]
ctx = Context(cfg.height, cfg.width, 1, 0, max_batch=148, max_blobs=mb, max_tracks=mt, white_on_dark=name != 'cfg4')
cs, bs = [], []
for a in range(0, F, 148):
    e = min(F, a + 148)
    fr = torch.empty((e - a, cfg.height, cfg.width), dtype=torch.uint8, device='cuda')
    render_frames_torch(scene, a, e, 'cuda', 1, out=fr)
    c, b = ctx.detect(fr, a)
    cs.append(c.clone()); bs.append(b.clone())
counts = torch.cat(cs); blobs = torch.cat(bs)
torch.cuda.synchronize()
hog = len(sys.argv) > 3 and sys.argv[3] == 'hog'
mm = len(sys.argv) > 3 and sys.argv[3] == 'mm'
ma = torch.randn(8192, 8192, device='cuda', dtype=torch.bfloat16) if mm else None
big = torch.empty(1 << 30, dtype=torch.uint8, device='cuda') if hog else None
big2 = torch.empty_like(big) if hog else None
side = torch.cuda.Stream()
for rep in range(3):
    ctx.reset(); ctx.set_profiling(True); ctx.get_profile()
    if mm:                       # compute-bound GEMMs on the other SMs (little memory traffic)
        with torch.cuda.stream(side):
            for _ in range(40):
                torch.mm(ma, ma)
    if hog:                      # a DRAM-streaming copy on another stream for the whole duration of the link (memory contention)
        with torch.cuda.stream(side):
            for _ in range(40):
                big2.copy_(big)
    try:
        rows = ctx.link(counts, blobs, 0, int(os.environ.get('LINK_ROWS_CAP', F * (3000 if name == 'cfg3' else 400))))
    except Exception as ex:          # (LINK_ROWS_CAP=1: the row overflow is the point -- no row is stored)
        rows = []
    p = ctx.get_profile()['link']
    torch.cuda.synchronize()
    print(f'{name}: link {"beside a 1 GB copy loop" if hog else ("beside bf16 GEMMs" if mm else "alone")} {p[0]:.3f} ms for {F} frames = {1e3 * p[0] / F:.3f} us/frame ({p[1]} profiled launches), rows {len(rows)}')
