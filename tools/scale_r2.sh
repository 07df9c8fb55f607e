# multi-GPU evidence (run under gpurun --gpus N): data-parallel bench and config 4 in row bands.   bash tools/scale_r2.sh N
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_n$N.err; tail -2 gpurun_out/r2g_n$N.err
$T bench.py --gpus $N --mode rowband --steps 5 --warmup 2 > gpurun_out/r2g_rowband_n$N.json 2> gpurun_out/r2g_rb$N.err; tail -2 gpurun_out/r2g_rb$N.err
python - <<PY
import json
for f in ("gpurun_out/r2g_bench_n$N.json", "gpurun_out/r2g_rowband_n$N.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d.get("per_rank_ms"), d.get("same_seed_control"), d["e2e"]["value"], d.get("halo"))
    except Exception as e:
        print(f, "unreadable", e)
PY
