"""Small driver for ncu captures: a few batches of the cfg2 (or cfg3 / cfg4) scene through ysmr_detect."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ysmr_b200.api import Context
from ysmr_b200.synth import CONFIGS, make_scene, render_frames_torch
import dataclasses

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 296
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
gen = int(sys.argv[4]) if len(sys.argv) > 4 else 4
cfg = dataclasses.replace(CONFIGS[name], n_frames=n)
scene = make_scene(cfg)
dev = torch.device('cuda', 0)
frames = torch.empty((n, cfg.height, cfg.width, 3), dtype=torch.uint8, device=dev)
render_frames_torch(scene, 0, n, dev, channels=3, out=frames)
wod = name != 'cfg4'
ctx = Context(cfg.height, cfg.width, 3, 0, max_batch=n, max_blobs=4096 if name == 'cfg3' else 512, max_tracks=8192,
              white_on_dark=wod)
ctx.set_option(ctx.OPT_FRONTEND_GEN, gen)
ctx.set_profiling(True)
for _ in range(reps):
    c, b = ctx.detect(frames, 0)
torch.cuda.synchronize()
prof = ctx.get_profile()
print({k: (round(v[0] / reps, 3), v[1]) for k, v in prof.items()}, 'blobs/frame', float(c.float().mean()))
