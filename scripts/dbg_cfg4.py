"""Finds where the GPU rows of a configuration's first frames leave the oracle's (detection counts, then rows)."""
import sys, os, dataclasses
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ysmr_b200.api import Context
from ysmr_b200.synth import CONFIGS, make_scene, render_frames_torch
from oracle import ref_stages
from oracle.tracker_port import LinkerPort

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg4'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
cfg = dataclasses.replace(CONFIGS[name], n_frames=n)
scene = make_scene(cfg)
fr = torch.empty((n, cfg.height, cfg.width, 3), dtype=torch.uint8, device='cuda')
for a in range(0, n, 50):
    render_frames_torch(scene, a, min(n, a + 50), 'cuda', 3, out=fr[a:min(n, a + 50)])
wod = name != 'cfg4'
ctx = Context(cfg.height, cfg.width, 3, 0, max_batch=64, max_blobs=4096 if name == 'cfg3' else 1024, max_tracks=8192, white_on_dark=wod)
cs, bs = [], []
for a in range(0, n, 64):
    c, b = ctx.detect(fr[a:a + 64], a); cs.append(c.cpu().numpy()); bs.append(b.cpu().numpy())
ctx.status()
counts = np.concatenate(cs); blobs = np.concatenate(bs)
st = ref_stages.DetectSettings(white_on_dark=wod)
host = fr.cpu().numpy()
lp = LinkerPort(max_disappeared=30.0, fps=30.0)
ref_rows = []
bad = None
for t in range(n):
    r = ref_stages.detect_frame(host[t], st)
    ref = ref_stages.rects_to_array(r['rects'])
    if len(ref) != counts[t]:
        print('frame', t, 'count', counts[t], 'vs oracle', len(ref)); bad = t; break
    d = np.abs(blobs[t, :len(ref)] - ref).max() if len(ref) else 0
    if d > 1e-3: print('frame', t, 'rect diff', d)
    ref_rows += [(t, i, xy[0], xy[1]) for (i, xy, info) in lp.update(r['rects'])]
print('detection compared', 'stopped at' if bad is not None else 'all ok', bad)
ref_rows = np.array(ref_rows)
c = torch.from_numpy(counts).cuda(); b = torch.from_numpy(blobs).cuda()
rows = ctx.link(c, b, 0, rows_capacity=n * 2048)
np.savez_compressed('gpurun_out/dbg_seq.npz', counts=counts, blobs=blobs[:, :counts.max()], ref=ref_rows, gpu=np.stack([rows['frame'], rows['track_id']], 1))
print('rows gpu', len(rows), 'oracle', len(ref_rows))
k = min(len(rows), len(ref_rows))
neq = np.nonzero((rows['frame'][:k] != ref_rows[:k, 0]) | (rows['track_id'][:k] != ref_rows[:k, 1]))[0]
if len(neq):
    i = neq[0]; print('first id/frame mismatch at row', i, 'gpu', rows['frame'][i], rows['track_id'][i], 'ref', ref_rows[i, :2])
    f = int(ref_rows[i, 0])
    print('frame', f, 'n gpu rows', (rows['frame'] == f).sum(), 'ref', (ref_rows[:, 0] == f).sum(), 'm', counts[f], 'prev frame rows', (ref_rows[:, 0] == f - 1).sum(), 'm prev', counts[f - 1])
else:
    print('ids identical over', k, 'rows; max xy err', np.abs(rows['x'][:k] - ref_rows[:k, 2]).max())
