import numpy as np, torch, os, sys
sys.path.insert(0,'.')
from tests.util import GOLDEN, oracle_rows
from oracle import ref_stages
from ysmr_b200.api import Context
from ysmr_b200.synth import SceneConfig, make_scene, render_frames
g=np.load(os.path.join(GOLDEN,"e2e_small_meanstd.npz"))
kw={k[6:]: g[k].item() for k in g.files if k.startswith("scene_")}
cfg=SceneConfig(**kw); grey=render_frames(make_scene(cfg))
st=ref_stages.DetectSettings(True,int(g["offset"]),float(g["adt"]),float(g["fps"]))
rows,per=oracle_rows(grey,st,fps=float(g["fps"]))
ctx=Context(cfg.height,cfg.width,1,0,white_on_dark=True,offset=int(g["offset"]),adt=float(g["adt"]),fps=float(g["fps"]),max_batch=32,max_blobs=1024,max_tracks=1024)
fr=torch.from_numpy(grey).cuda()
allc=[];allb=[]
for a in range(0,len(grey),32):
    c,b=ctx.detect(fr[a:a+32],a); allc.append(c.cpu().numpy()); allb.append(b.cpu().numpy())
C=np.concatenate(allc); B=np.concatenate(allb)
bad=0
for t in range(len(grey)):
    ref=per[t]
    if C[t]!=len(ref): print('count mismatch',t,C[t],len(ref)); bad+=1; continue
    d=np.abs(B[t,:C[t]]-ref).max() if len(ref) else 0
    if d>1e-3: print('det mismatch frame',t,d); print(B[t,:C[t]]); print(ref); bad+=1
print('detect mismatches',bad)
ctx.reset()
got=ctx.link(torch.from_numpy(C).cuda(), torch.from_numpy(B).cuda(), 0, 200*64)
print(len(got),len(rows))
n=min(len(got),len(rows))
e=np.maximum(np.abs(got['x'][:n]-rows[:n,2]),np.abs(got['y'][:n]-rows[:n,3]))
i=np.argmax(e>1e-6)
print('first row mismatch idx',i,'frame',rows[i,0],'id',rows[i,1],e[i])
for k in range(max(0,i-3),i+6): print(rows[k], got[k])
# which frames have M>N?
