# config 4 in row bands over N GPUs (run under gpurun --gpus N):   bash tools/rowband_r2.sh N
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus $N --mode rowband --steps 5 --warmup 2 > gpurun_out/r2e_rowband_n$N.json 2> gpurun_out/r2e_rb$N.err; tail -2 gpurun_out/r2e_rb$N.err
python - <<PY
import json
f = "gpurun_out/r2e_rowband_n$N.json"
d = json.loads([l for l in open(f) if l.startswith("{")][-1])
print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("halo"))
PY
