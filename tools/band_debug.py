import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from multi_frame_super_resolution_b200 import rowband
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst
n,h,w=4,1152,512
fr,_=synth_burst(n,h,w,seed=11); p=default_params(); dev=fr.cuda()
sr=BurstSuperResolution(p,0,w,h,n); sr.set_input(dev); full=sr.next_frame().cpu().numpy(); sr.close()
for world in (2,3,4):
    for halo in (128,256):
        try: bands=rowband.plan_bands(h,world,128,halo)
        except ValueError as e: print(world,halo,e); continue
        out=np.empty_like(full)
        for b in bands:
            srb=BurstSuperResolution(rowband.band_params(p,b,h),0,w,b.bottom-b.top,n); srb.set_input(dev[:,b.top:b.bottom].contiguous())
            out[2*b.row0:2*b.row1]=srb.next_frame().cpu().numpy(); srb.close()
        d=np.abs(out-full); print('world',world,'halo',halo,'equal',np.array_equal(out,full),'max',d.max(),'frac>1e-3',(d>1e-3).mean(),'frac!=',(d!=0).mean())
