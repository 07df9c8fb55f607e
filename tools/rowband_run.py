#!/usr/bin/env python
"""One very large burst split into row bands over the GPUs of a box (BASELINE config 4: 8064x6048 x 15 frames, 2x).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/rowband_run.py [--height H --width W --frames N --verify]

Every rank owns a band of every frame (here cut out of the same seeded synthetic burst), receives its halo rows from
the neighbouring ranks over NCCL (the path's only exchange step), runs the chain in row-band mode and keeps its output
rows.  Prints one JSON line on rank 0: output MP/s of the whole burst (max over ranks), halo bytes, and with --verify the
comparison of the stitched image with the single-GPU full-frame run (gathered on rank 0; small sizes only)."""
import argparse, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
from multi_frame_super_resolution_b200 import rowband
from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
from multi_frame_super_resolution_b200.synth import synth_burst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=6048); ap.add_argument("--width", type=int, default=8064)
    ap.add_argument("--frames", type=int, default=15); ap.add_argument("--halo", type=int, default=rowband.DEFAULT_HALO)
    ap.add_argument("--steps", type=int, default=3); ap.add_argument("--verify", action="store_true")
    ap.add_argument("--margin", type=int, default=-1, help="-1: derived from the params (rowband.default_margin)")
    a = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    p = default_params()
    bands = rowband.plan_bands(a.height, world, p.tile_size << (p.levels - 1), a.halo)
    b = bands[rank]
    # the rank's own rows of every frame (a real deployment receives only these); generated band-wise to bound memory
    full, _ = synth_burst(a.frames, a.height, a.width, seed=4321, device=dev)
    own = full[:, b.row0:b.row1].contiguous()
    if not a.verify:
        del full
    torch.cuda.synchronize(dev)
    a.margin = rowband.default_margin(p) if a.margin < 0 else a.margin
    bp = rowband.band_params(p, b, a.height, a.margin)
    sr = BurstSuperResolution(bp, device=lr, max_width=a.width, max_height=b.bottom - b.top, max_frames=a.frames)
    ow, oh = sr.output_size(a.width, b.bottom - b.top)
    out = torch.empty((oh, ow, 3), dtype=torch.float32, device=dev)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    times = []
    for it in range(a.steps + 1):
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0.record()
        banded = rowband.exchange_halos(own, bands, rank) if world > 1 else own
        e1.record()
        sr.set_input(banded)
        sr.next_frame(out=out)
        e2.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e2), e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it > 0:
            times.append(t.tolist())
    ms = sum(x[0] for x in times) / len(times); ms_halo = sum(x[1] for x in times) / len(times)
    line = {"workload": f"{a.width}x{a.height} x {a.frames} frames, 2x, row bands", "n_gpus": world, "ms_per_burst": round(ms, 2),
            "ms_halo_exchange": round(ms_halo, 3), "output_megapixels_per_second": round(4 * a.width * a.height / 1e6 / (ms / 1e3), 1),
            "halo_rows": a.halo, "margin_rows": a.margin, "halo_bytes_per_rank": int((b.halo_up + b.halo_down) * a.width * 2 * a.frames),
            "band_rows": [x.rows for x in bands], "processed_rows": [x.bottom - x.top for x in bands]}
    if a.verify:
        stitched = None
        if world > 1:
            parts = [torch.empty((2 * x.rows, 2 * a.width, 3), dtype=torch.float32, device=dev) for x in bands] if rank == 0 else None
            dist.gather(out, parts, dst=0)
            if rank == 0:
                stitched = torch.cat(parts, 0)
        else:
            stitched = out
        if rank == 0:
            srf = BurstSuperResolution(p, device=lr, max_width=a.width, max_height=a.height, max_frames=a.frames)
            srf.set_input(full); ref = srf.next_frame(); srf.synchronize()
            line["bit_identical_to_single_gpu"] = bool(torch.equal(stitched, ref))
            line["max_abs_diff"] = float((stitched - ref).abs().max())
    if rank == 0:
        print(json.dumps(line))
    sr.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
