import sys; sys.path.insert(0,'/root/repo')
import numpy as np, vecchio_b200 as vb
from oracle import pyoracle as po
scene = vb.Scene("stress_spheres", seed=1, param=1000); cam = scene.next_camera()
ctx = vb.Context(0); ctx.upload(scene)
for W,H in ((192,108),(960,540),(3840,2160)):
    for fl,name in ((0,"fast"),(vb.VK_FLAG_STRICT_MATH,"strict")):
        for var in (1,2):
            r,_,st = ctx.render(cam, vb.render_params(W,H,4,50,seed=1,flags=fl,variant=var))
            print(W,H,name,"variant",var,"rays/path",st.rays/st.paths,"mean",r.mean(),"dropped",st.dropped_samples, flush=True)
o = po.OracleScene(scene)
for W,H in ((192,108),(480,270)):
    r,_,so = o.render(cam, vb.render_params(W,H,4,50,seed=2))
    print("oracle",W,H,"rays_live/path",so.rays_live/so.paths,"rays/path",so.rays/so.paths,"mean",r.mean(), flush=True)
