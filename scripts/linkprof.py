import numpy as np, torch, sys
sys.path.insert(0,'.')
from ysmr_b200.api import Context
from ysmr_b200.synth import CONFIGS, SceneConfig, make_scene, render_frames_torch
import dataclasses
import os
F=int(os.environ.get('LINKPROF_FRAMES','1024'))
PLAIN=bool(os.environ.get('LINKPROF_PLAIN'))
CFG=os.environ.get('LINKPROF_CFG','cfg2')
cfg=dataclasses.replace(CONFIGS[CFG],n_frames=F)
scene=make_scene(cfg)
H,W=cfg.height,cfg.width
fr=torch.empty((F,H,W),dtype=torch.uint8,device='cuda')
for a in range(0,F,64): render_frames_torch(scene,a,min(F,a+64),'cuda',1,out=fr[a:min(F,a+64)])
ctx=Context(H,W,1,0,max_batch=256,max_blobs=4096 if CFG=='cfg3' else 512,max_tracks=8192,white_on_dark=CFG!='cfg4')
cs=[];bs=[]
for a in range(0,F,256):
    c,b=ctx.detect(fr[a:a+256],a); cs.append(c); bs.append(b)
counts=torch.cat(cs); blobs=torch.cat(bs)
HOG=bool(os.environ.get('LINKPROF_HOG'))
big=torch.empty(1<<30,dtype=torch.uint8,device='cuda') if HOG else None
big2=torch.empty_like(big) if HOG else None
side=torch.cuda.Stream()
for rep in range(2):
    ctx.reset(); ctx.set_profiling(True, link_phases=not PLAIN)
    torch.cuda.synchronize()
    if HOG:
        with torch.cuda.stream(side):
            for _ in range(60): big2.copy_(big)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); rows=ctx.link(counts,blobs,0,F*(3000 if CFG=='cfg3' else 400)); e1.record(); torch.cuda.synchronize()
    pc=ctx.link_phase_cycles()
    print('link ms', e0.elapsed_time(e1), 'rows', len(rows), 'frames', pc[12])
    names=['loop top + staging','candidate/scan/claim','barrier 2','conflict+outcome','barrier 4 (vote)','events','gsff','row + loop end']
    for i in range(8): print('  %-22s %8.0f cyc/frame'%(names[i], pc[i]/max(pc[12],1)))
    print('  total cyc/frame', sum(pc[:8])/max(pc[12],1), ' exact scans/frame', pc[8]/max(pc[12],1), ' conflict frames', pc[9], ' event frames', pc[10])
# association regimes of the scene: n tracks (rows per frame) against m detections
fr_ids = np.asarray(rows['frame']); npf = np.bincount(fr_ids, minlength=F)
m = counts.cpu().numpy().astype(int)
n_before = np.concatenate([[0], npf[:-1]])          # tracks alive when the frame's association runs
print('frames n==m', int((n_before == m).sum()), 'n>m', int((n_before > m).sum()), 'n<m', int((n_before < m).sum()))
print('frames where track count changes', int((np.diff(npf) != 0).sum()), 'mean n', npf.mean(), 'mean m', m.mean())
print('hist of n-m', np.unique(n_before - m, return_counts=True))
