import numpy as np, torch, sys
sys.path.insert(0,'.')
from ysmr_b200.api import Context
from ysmr_b200.synth import SceneConfig, make_scene, render_frames_torch
F=512
scene=make_scene(SceneConfig(n_frames=F,n_cells=50,seed=0))
fr=torch.empty((F,922,1228),dtype=torch.uint8,device='cuda')
for a in range(0,F,128): render_frames_torch(scene,a,a+128,'cuda',1,out=fr[a:a+128])
ctx=Context(922,1228,1,0,max_batch=256,max_blobs=512,max_tracks=1024)
cs=[];bs=[]
for a in range(0,F,256):
    c,b=ctx.detect(fr[a:a+256],a); cs.append(c); bs.append(b)
counts=torch.cat(cs); blobs=torch.cat(bs)
# forward-backward-forward... so that positions stay continuous
counts=torch.cat([counts,counts.flip(0)]*4).contiguous(); blobs=torch.cat([blobs,blobs.flip(0)]*4).contiguous()
ctx.reset()
rows=ctx.link(counts,blobs,0,len(counts)*100)
print('rows',len(rows),'frames',len(counts))
