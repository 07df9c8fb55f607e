import sys, torch
sys.path.insert(0, '.')
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst
n,h,w=8,3024,4032
dev=torch.device('cuda',0)
sr=BurstSuperResolution(default_params(), device=0, max_width=w, max_height=h, max_frames=n)
ow,oh=sr.output_size(w,h); out=torch.empty((oh,ow,3),dtype=torch.float32,device=dev)
for seed in (1234,1235,1236,1237):
    fr,sh=synth_burst(n,h,w,seed=seed,device=dev)
    for _ in range(2): sr.set_input(fr); sr.next_frame(out=out)
    torch.cuda.synchronize()
    st=sr.stage_ms()
    print(seed, {k:round(v,2) for k,v in st.items()}, 'sum', round(sum(st.values()),2), 'shifts', [[round(float(x),1) for x in r] for r in sh.tolist()][:8])
