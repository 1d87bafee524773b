import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
import oracle, simplestereo_b200 as ss
from simplestereo_b200.synth import synth_pair
for (W,H,win,maxd,gp) in ((700,40,35,127,17.5),(700,40,35,127,40.0),(700,56,51,127,17.5),(700,56,51,127,34.0),(700,44,39,127,40.0)):
    l,r,_=synth_pair(W,H,maxd,1)
    kw=dict(winSize=win,maxDisparity=maxd,minDisparity=0,gammaC=12.0,gammaP=gp,consistent=False)
    rows=(H//2-1,H//2+1)
    ref=oracle.asw(l,r,stages=True,cost=True,rows=rows,**kw)["cost"]
    out={}
    for tc in ("1","0"):
        os.environ["SS_TCDEN"]=tc
        g=ss.passive.StereoASW(**kw).compute_staged(l,r,cost=True)["cost"][rows[0]:rows[1]].astype(np.float64)
        fin=np.isfinite(ref)
        rel=(g[fin]-ref[fin])/ref[fin]
        out[tc]=(rel.min(),rel.max(),np.abs(rel).mean())
    print(f"win {win} gammaP {gp}: TC rel err min {out['1'][0]:+.2e} max {out['1'][1]:+.2e} mean|.| {out['1'][2]:.2e} | CUDA-core min {out['0'][0]:+.2e} max {out['0'][1]:+.2e} mean|.| {out['0'][2]:.2e}")
