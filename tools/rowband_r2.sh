# config 4 in row bands over N GPUs (run under gpurun --gpus N):   bash tools/rowband_r2.sh N
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
# small size, stitched image compared with the single-GPU run bit for bit; then config 4 with both halo transports
$T bench.py --gpus $N --mode rowband --band-height $((N > 2 ? 768 * N : 1536)) --band-width 1024 --band-frames 4 --verify --steps 3 > gpurun_out/r2e_rowband_verify_n$N.json 2> gpurun_out/r2e_rbv$N.err; tail -3 gpurun_out/r2e_rbv$N.err
$T bench.py --gpus $N --mode rowband --steps 5 --warmup 2 > gpurun_out/r2e_rowband_n$N.json 2> gpurun_out/r2e_rb$N.err; tail -2 gpurun_out/r2e_rb$N.err
MFSR_HALO=nccl $T bench.py --gpus $N --mode rowband --steps 5 --warmup 2 > gpurun_out/r2e_rowband_nccl_n$N.json 2> gpurun_out/r2e_rbn$N.err; tail -2 gpurun_out/r2e_rbn$N.err
python - <<PY
import json
for f in ("gpurun_out/r2e_rowband_verify_n$N.json", "gpurun_out/r2e_rowband_n$N.json", "gpurun_out/r2e_rowband_nccl_n$N.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("halo"), d.get("verify"))
    except Exception as e:
        print(f, "unreadable", e)
PY
