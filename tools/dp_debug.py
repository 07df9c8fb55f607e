import os, sys, time, torch
sys.path.insert(0, '.')
import torch.distributed as dist
rank=int(os.environ.get('RANK','0')); lr=int(os.environ.get('LOCAL_RANK','0')); world=int(os.environ.get('WORLD_SIZE','1'))
torch.cuda.set_device(lr)
dev=torch.device('cuda',lr)
if world>1: dist.init_process_group('nccl', device_id=dev)
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst
n,h,w=8,3024,4032
sr=BurstSuperResolution(default_params(), device=lr, max_width=w, max_height=h, max_frames=n)
fr,_=synth_burst(n,h,w,seed=1234+rank,device=dev)
ow,oh=sr.output_size(w,h); out=torch.empty((oh,ow,3),dtype=torch.float32,device=dev)
for _ in range(3): sr.set_input(fr); sr.next_frame(out=out)
torch.cuda.synchronize(dev)
if world>1: dist.barrier()
t0=time.perf_counter()
for _ in range(5): sr.set_input(fr); sr.next_frame(out=out)
torch.cuda.synchronize(dev)
dt=(time.perf_counter()-t0)/5*1e3
print(f'rank {rank} dev {lr} {torch.cuda.get_device_properties(lr).uuid} wall ms/step {dt:.2f} stages {sum(sr.stage_ms().values()):.2f} cpus {os.cpu_count()} affinity {len(os.sched_getaffinity(0))}', flush=True)
