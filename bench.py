#!/usr/bin/env python
"""bench.py — output MP/s of the burst super-resolution hot path + merge-stage HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode dp|rowband]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

--mode dp (default): one "step" = one synthetic burst (BASELINE.json configs[1]: 12 MP RGGB, 8 frames, 2x, full frame) per rank
through the whole chain (front end -> pyramid tile alignment -> consolidation -> LK flow -> kernel params -> robustness -> fused
merge).  Independent bursts shard data-parallel, no data-path collective: weak scaling.
--mode rowband: one step = ONE 48 MP x 15-frame burst (configs[3]) split into row bands over the ranks, halo rows exchanged over
NCCL (the path's only exchange step): strong scaling.
--impl reference: the reference's CPU path (the oracle port, C + OpenMP on all host threads) on the same 12 MP burst, rank 0 only.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "output_megapixels_per_second"
UNIT = "MP/s"
CFG = dict(frames=8, height=3024, width=4032, scale=2)
CFG_BAND = dict(frames=15, height=6048, width=8064, scale=2)


def merge_bytes_per_px(n, s, gray=False):
    """compulsory traffic of the fused merge per OUTPUT pixel."""
    return ((11 if gray else 14) * n + 16) / (s * s) + (8 if gray else 24)


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(self.samples[0][1]), "reasons": reasons, "samples": len(sm)}


def hbm_peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_cpu_port(steps, warmup):
    """The reference's CPU path = the oracle port (C + OpenMP, all host threads) on the WHOLE 12 MP x 8 burst.  Loads nothing but
    oracle/ (numpy burst generator, oracle-side defaults): the product library stays out of this process's arm."""
    from oracle import pyoracle as O
    from oracle.synth_np import synth_burst_np
    cores = O.set_threads(os.cpu_count() or 1)
    p = O.default_params()
    fr, _ = synth_burst_np(CFG["frames"], CFG["height"], CFG["width"], seed=1234)
    out_mp = CFG["height"] * CFG["width"] * p.scale * p.scale / 1e6
    for _ in range(warmup):
        O.run_pipeline(fr, p)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.run_pipeline(fr, p)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    sample = f"the whole burst: {CFG['frames']} frames of {CFG['width']}x{CFG['height']} RGGB, whole chain, {steps} timed step(s)"
    return out_mp / dt, dt * 1e3, cores, sample


def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs of the GPU's NUMA node BEFORE any pinned buffer is allocated and touched (first-touch placement):
    eight ranks staging through one node's memory was the e2e limiter of round 1.  Returns a description for the JSON line."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        if bus is None:
            q = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True, timeout=5)
            bus = q.stdout.strip()
        bus = str(bus).lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        base = Path("/sys/bus/pci/devices") / bus
        node = int((base / "numa_node").read_text().strip())
        cpulist = (base / "local_cpulist").read_text().strip()
        if node < 0 or not cpulist:
            return {"numa_node": node, "bound": False}
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "bound": True, "cpus": len(cpus)}
        return {"numa_node": node, "bound": False}
    except Exception as e:      # VMs without NUMA information: nothing to bind to
        return {"numa_node": None, "bound": False, "why": type(e).__name__}


def host_stream_gbs():
    """STREAM-style copy bandwidth of this rank's host memory (numpy copy of 256 MB, best of 3): context for the e2e ceiling."""
    import numpy as np
    a = np.ones(32 * 1024 * 1024, np.float64)
    b = np.empty_like(a)
    best = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        np.copyto(b, a)
        best = max(best, 2 * a.nbytes / (time.perf_counter() - t0) / 1e9)
    return round(best, 1)


def reference_arm(args, config):
    steps, warm = 1, 0          # ~20-40 s per 12 MP burst on the box's host threads: one timed step, no warm-up
    val, ms, cores, sample = run_cpu_port(steps, warm)
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "requested_steps": args.steps, "requested_warmup": args.warmup,
            "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference CPU path = oracle port (C/OpenMP) on the same 12 MP x 8 burst; the reference's kernels have no CPU implementation and its "
                    "OpenCV superres host path is not installable (DESIGN.md); rank 0 only, bursts_per_step = 1"}
    print(json.dumps(line))


def ref_gpu_block(dev):
    """The reference's OWN merge kernels (8 x accumulateImagesSuperRes + ApplyWeighting + GammasRGB, compiled unmodified into
    oracle/_ref) against mfsr_stage_merge on the same buffers and the reference's geometry (12.2 MP out), both on this GPU."""
    import torch
    from oracle import pyref
    if not pyref.available():
        return {"unavailable": "oracle/_ref/libmfsr_ref.so not built"}
    from multi_frame_super_resolution_b200 import stages
    from multi_frame_super_resolution_b200._lib import MergeGeom
    from multi_frame_super_resolution_b200.synth import synth_merge_inputs
    n, h, w = CFG["frames"], CFG["height"], CFG["width"]
    white, black = [959.0] * 3, [64.0] * 3
    raw, mask, flow, kern = synth_merge_inputs(n, h, w, seed=1234, device=dev)
    fb = torch.rand((h, w, 3), device=dev)
    ms_ref = min(pyref.merge_chain_ms(raw, mask, flow, kern, fb, white, black, 0.1, [0, 1, 1, 2], gamma=True)[0] for _ in range(3))
    geom = MergeGeom.reference(w, h)
    out = torch.empty((h, w, 3), device=dev)
    best = 1e30
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        stages.merge(raw, mask, flow, kern, fb, geom, white, black, 0.1, flags=1, out=out)
        e1.record()
        torch.cuda.synchronize(dev)
        if i:
            best = min(best, e0.elapsed_time(e1))
    return {"ms_merge_ref": round(ms_ref, 3), "ms_merge_ours": round(best, 3), "ratio": round(ms_ref / best, 2),
            "workload": f"merge stage alone, {n} frames of {w}x{h}, reference geometry (central crop, {w}x{h} out), synthetic flows / masks",
            "kernels": "reference: accumulateImagesSuperRes x 8 + ApplyWeighting + GammasRGB (DeBayerKernels.cu:379, kernel.cu:426,393) vs mfsr_stage_merge"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="dp", choices=["dp", "rowband"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true", help="rowband: compare the stitched image with the single-GPU run (small sizes)")
    ap.add_argument("--band-height", type=int, default=CFG_BAND["height"])
    ap.add_argument("--band-width", type=int, default=CFG_BAND["width"])
    ap.add_argument("--band-frames", type=int, default=CFG_BAND["frames"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"synthetic {CFG['width']}x{CFG['height']} (12 MP) RGGB burst, {CFG['frames']} frames, {CFG['scale']}x, full frame"
    config = {"workload": workload, "frames": CFG["frames"], "raw": [CFG["width"], CFG["height"]], "scale": CFG["scale"],
              "bursts_per_step": max(world, 1), "parallelism": f"dp{max(world, 1)} (independent bursts, no data-path collective)",
              "l2": "inputs larger than L2 (raw 195 MB + flow 780 MB + masks 390 MB per burst vs 126 MB L2)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        config["bursts_per_step"] = 1
        config["parallelism"] = "host threads of rank 0 (OpenMP)"
        reference_arm(args, config)
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa(local_rank)
    if args.mode == "rowband":
        return bench_rowband(args, rank, world, local_rank, numa)
    return bench_dp(args, rank, world, local_rank, numa, config)


def bench_dp(args, rank, world, local_rank, numa, config):
    import numpy as np
    import torch
    import torch.distributed as dist
    from multi_frame_super_resolution_b200 import dp
    from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
    from multi_frame_super_resolution_b200.synth import synth_burst

    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, h, w, s = CFG["frames"], CFG["height"], CFG["width"], CFG["scale"]
    p = default_params()
    sr = BurstSuperResolution(p, device=local_rank, max_width=w, max_height=h, max_frames=n)
    # every step processes `world` bursts; rank r owns burst r (dp.shard_bursts) and burst b uses seed 1234 + b
    my_bursts = dp.shard_bursts(world, rank, world)
    frames, _ = synth_burst(n, h, w, seed=dp.burst_seed(1234, my_bursts[0]), device=dev)
    ow, oh = sr.output_size(w, h)
    out_dev = torch.empty((oh, ow, 3), dtype=torch.float32, device=dev)
    out_mp = ow * oh / 1e6
    stream = torch.cuda.ExternalStream(sr.stream, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def gather(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        if world > 1:
            parts = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            return [round(float(x.item()), 3) for x in parts]
        return [round(float(v), 3)]

    def time_resident(fr, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launches = 0
        for _ in range(steps):
            sr.set_input(fr)                    # D2D staging / in-place use of the frame stack (part of the step)
            sr.next_frame(out=out_dev)
            launches += sr.launch_count()
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), launches

    # ---------------- device-resident timing (value)
    for _ in range(args.warmup):
        sr.set_input(frames)
        sr.next_frame(out=out_dev)
    sampler = ClockSampler(local_rank); sampler.start()
    torch.cuda.cudart().cudaProfilerStart()     # `ncu --profile-from-start off` lists exactly the timed steps (profiles/*_launches.txt); a no-op otherwise
    ms_total, launches = time_resident(frames, args.steps)
    torch.cuda.cudart().cudaProfilerStop()
    stage_last = sr.stage_ms()
    merge_ms = []
    for _ in range(args.steps):                 # merge kernel: average launch duration, each launch preceded by the rest of the chain
        sr.set_input(frames)
        sr.next_frame(out=out_dev)
        merge_ms.append(sr.stage_ms()["merge"])
    barrier()
    clocks = sampler.summary()
    per_rank_ms = gather(ms_total / args.steps)
    _, ms_step, value = dp.aggregate_throughput(out_mp, ms_total / args.steps, device=dev)
    # same-seed control: every rank processes the SAME burst (rank 0's); the drop of the slowest rank that remains is the machine's,
    # the rest of the N > 1 efficiency gap is the burst dependence of the step time
    same_seed = None
    if world > 1:
        fr0 = frames if rank == 0 else synth_burst(n, h, w, seed=dp.burst_seed(1234, 0), device=dev)[0]
        for _ in range(2):
            sr.set_input(fr0); sr.next_frame(out=out_dev)
        ms0, _ = time_resident(fr0, max(3, args.steps // 2))
        per0 = gather(ms0 / max(3, args.steps // 2))
        same_seed = {"per_rank_ms": per0, "ms_per_step": max(per0), "value": round(out_mp * world / (max(per0) / 1e3), 2)}
        del fr0

    # ---------------- end-to-end through the C ABI with HOST buffers (e2e)
    # Every step: H2D of the 8 raw frames from pinned memory (mfsr_set_frames), the whole chain, D2H of the result into pinned memory
    # (mfsr_run_format).  Handles are driven alternately so that one burst's PCIe transfers overlap the other bursts' kernels; a
    # handle is synchronised before it is re-used.  Headline format: 8-bit RGB, what the reference program delivers
    # (multi_frame_sr.cpp:207 writes the result as an 8-bit PNG); the float3 and half3 deliveries are reported beside it.
    host_in = torch.empty((n, h, w), dtype=torch.int16, pin_memory=True)
    host_in.copy_(frames)
    host_np = host_in.numpy().view(np.uint16)
    n_handles = max(2, int(os.environ.get("BENCH_E2E_HANDLES", "3")))
    extra = [BurstSuperResolution(p, device=local_rank, max_width=w, max_height=h, max_frames=n) for _ in range(n_handles - 1)]
    handles = [sr] + extra

    def step_e2e(i, outs, dtype):
        hd = handles[i % n_handles]
        hd.synchronize()                                      # its previous burst (step i - n_handles) has fully landed
        hd.set_input(host_np)                                 # async H2D inside the timed region
        hd.next_frame(out=outs[i % n_handles], host=True, sync=False, dtype=dtype)   # chain + async D2H

    e2e_steps = max(4, min(args.steps, 10))

    def time_e2e(dtype):
        outs = [torch.empty((oh, ow, 3), dtype=dtype, pin_memory=True) for _ in handles]
        for i in range(2 * n_handles):             # untimed warm-up: two rounds over the handles (first touches of the pinned buffers)
            step_e2e(i, outs, dtype)
        for hd in handles:
            hd.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            step_e2e(i, outs, dtype)
        for hd in handles:
            hd.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        per = gather(ms)
        return max(per), per, outs

    e2e_ms, e2e_per, outs8 = time_e2e(torch.uint8)
    del outs8
    e2e16_ms, _, o16 = time_e2e(torch.float16)
    del o16
    e2e32_ms, _, outs32 = time_e2e(torch.float32)
    for hd in extra:
        hd.close()

    # PCIe context for the e2e number (outside every timed region): plain pinned copies of the step's buffers, alone on the link
    def _copy_gbs(dst, src, reps=3):
        best = 0.0
        for _ in range(reps):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize(dev)
            best = max(best, src.numel() * src.element_size() / (time.perf_counter() - t0) / 1e9)
        return round(best, 1)
    barrier()
    pcie = {"d2h_gbs": _copy_gbs(outs32[0], out_dev), "h2d_gbs": _copy_gbs(frames, host_in)}
    pcie_all = {k: gather(v) for k, v in pcie.items()}
    stream_gbs = gather(host_stream_gbs())

    if rank == 0:
        peak, peak_src = hbm_peak()
        bpp = merge_bytes_per_px(n, s)
        mm = float(np.mean(merge_ms))
        achieved = bpp * ow * oh / (mm / 1e3) / 1e9
        traffic = None
        tf = ROOT / "profiles" / "merge_traffic.json"
        if tf.exists():
            try:
                traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "per_rank_ms": per_rank_ms,
                "e2e": {"value": round(out_mp * world / (e2e_ms / 1e3), 2), "unit": UNIT, "h2d_bytes_per_step": int(n * h * w * 2), "d2h_bytes_per_step": int(ow * oh * 3),
                        "ms_per_step": round(e2e_ms, 3), "per_rank_ms": e2e_per, "result_format": "8-bit RGB (mfsr_run_format MFSR_OUT_U8; the reference program writes an 8-bit PNG, multi_frame_sr.cpp:207)",
                        "pcie_alone": pcie, "pcie_alone_per_rank": pcie_all, "host_copy_gbs_per_rank": stream_gbs, "numa": numa,
                        "timed": f"host wall clock over the steps: mfsr_set_frames(pinned host) + mfsr_run_format(pinned host out, async) on {n_handles} alternating handles (transfers of one burst overlap kernels of the others), all synchronised at the end; max over ranks"},
                "e2e_f16_out": {"value": round(out_mp * world / (e2e16_ms / 1e3), 2), "unit": UNIT, "ms_per_step": round(e2e16_ms, 3), "d2h_bytes_per_step": int(ow * oh * 6)},
                "e2e_f32_out": {"value": round(out_mp * world / (e2e32_ms / 1e3), 2), "unit": UNIT, "ms_per_step": round(e2e32_ms, 3), "d2h_bytes_per_step": int(ow * oh * 12)},
                "gpu_launches": int(launches),
                "roofline": {"kernel": "merge_pf_kernel + merge_band_kernel (mfsr_stage_merge)", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                             "bytes_per_output_px": bpp, "ms_per_launch": round(mm, 4)},
                "stage_ms": {k: round(v, 3) for k, v in stage_last.items()},
                "clocks": clocks}
        if same_seed:
            line["same_seed_control"] = same_seed
        if not args.no_cpu_baseline and world == 1:
            del outs32
            line["ref_gpu"] = ref_gpu_block(dev)
            val, ms, cores, sample = run_cpu_port(1, 0)
            line["cpu_baseline"] = {"value": round(val, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms": round(ms, 1)}
        print(json.dumps(line))
    sr.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def bench_rowband(args, rank, world, local_rank, numa):
    """BASELINE configs[3]: ONE 48 MP x 15-frame burst, row bands over the ranks, halo rows over NCCL."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from multi_frame_super_resolution_b200 import dp, rowband
    from multi_frame_super_resolution_b200.pipeline import BurstSuperResolution, default_params
    from multi_frame_super_resolution_b200.synth import synth_burst

    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    H, W, N = args.band_height, args.band_width, args.band_frames
    p = default_params()
    margin = rowband.default_margin(p)
    bands = rowband.plan_bands(H, world, p.tile_size << (p.levels - 1), rowband.DEFAULT_HALO)
    b = bands[rank]
    # the rank's own rows of every frame (a deployment receives only these); the burst is generated frame by frame to bound memory
    full, _ = synth_burst(N, H, W, seed=4321, device=dev)
    own = full[:, b.row0:b.row1].contiguous()
    if not args.verify:
        full = None
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()
    bp = rowband.band_params(p, b, H, margin)
    sr = BurstSuperResolution(bp, device=local_rank, max_width=W, max_height=b.bottom - b.top, max_frames=N)
    ow, oh = sr.output_size(W, b.bottom - b.top)
    out = torch.empty((oh, ow, 3), dtype=torch.float32, device=dev)
    out_mp = 4.0 * W * H / 1e6

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # halo rows straight out of the neighbours' memory over NVLink (torch symmetric memory); grouped ncclSend/ncclRecv where that is not available
    peer, transport = None, "none (one rank)"
    if world > 1:
        transport = "nccl grouped send/recv"
        if os.environ.get("MFSR_HALO", "peer") == "peer":
            try:
                peer = rowband.PeerHaloExchange(N, bands, rank, W * 2, dev)
                transport = "peer memory over NVLink (torch symmetric memory), device-side barriers"
            except Exception as e:              # noqa: BLE001  (no symmetric memory on this box: the NCCL exchange is the same data)
                transport = f"nccl grouped send/recv (symmetric memory unavailable: {type(e).__name__})"
        flag = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)       # all ranks or none
        if int(flag.item()) == 0:
            peer = None
        if peer is not None:                              # the rank's own rows live in the shared band buffer from now on
            peer.own_view().copy_(own)
            own = peer.own_view()

    def step():
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        if world == 1:
            banded = own
        elif peer is not None:
            banded = peer.exchange()
        else:
            banded = rowband.exchange_halos(own, bands, rank)
        e1.record()
        sr.set_input(banded)
        sr.next_frame(out=out)
        e2.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e2), e0.elapsed_time(e1)

    for _ in range(max(1, min(args.warmup, 3))):
        sync_all(); step()
    sampler = ClockSampler(local_rank); sampler.start()
    steps = max(2, min(args.steps, 5))
    tot, halo, merge_ms, launches = [], [], [], 0
    for _ in range(steps):
        sync_all()
        a, c = step()
        t = torch.tensor([a, c], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tot.append(float(t[0])); halo.append(float(t[1])); merge_ms.append(sr.stage_ms()["merge"]); launches += sr.launch_count()
    clocks = sampler.summary()
    ms = float(np.mean(tot))
    per_rank = None
    if world > 1:
        tt = torch.tensor([float(np.mean(merge_ms))], dtype=torch.float64, device=dev)
        parts = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(parts, tt)
        per_rank = [round(float(x.item()), 3) for x in parts]
    # e2e: own rows from pinned host memory, exchange, chain, 8-bit band result to pinned host memory
    host_own = torch.empty((N, b.rows, W), dtype=torch.int16, pin_memory=True); host_own.copy_(own)
    host_out = torch.empty((oh, ow, 3), dtype=torch.uint8, pin_memory=True)
    own2 = own if peer is not None else torch.empty_like(own)

    def step_e2e():
        for f in range(N):                                # one dense copy per frame (own2 may be a strided view of the band buffer)
            own2[f].copy_(host_own[f], non_blocking=True)
        if world == 1:
            banded = own2
        elif peer is not None:
            banded = peer.exchange()
        else:
            banded = rowband.exchange_halos(own2, bands, rank)
        sr.set_input(banded)
        sr.next_frame(out=host_out, host=True, sync=True, dtype=torch.uint8)

    sync_all(); step_e2e()
    e2e = []
    for _ in range(steps):
        sync_all()
        t0 = time.perf_counter()
        step_e2e()
        torch.cuda.synchronize(dev)
        t = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e.append(float(t[0]))
    e2e_ms = float(np.mean(e2e))
    verify = None
    if args.verify:
        if world > 1:
            parts = [torch.empty((2 * x.rows, 2 * W, 3), dtype=torch.float32, device=dev) for x in bands] if rank == 0 else None
            sr.set_input(peer.exchange() if peer is not None else rowband.exchange_halos(own, bands, rank)); sr.next_frame(out=out); sr.synchronize()
            dist.gather(out, parts, dst=0)
            stitched = torch.cat(parts, 0) if rank == 0 else None
        else:
            stitched = out
        if rank == 0:
            srf = BurstSuperResolution(p, device=local_rank, max_width=W, max_height=H, max_frames=N)
            srf.set_input(full); ref = srf.next_frame(); srf.synchronize()
            verify = {"bit_identical_to_single_gpu": bool(torch.equal(stitched, ref)), "max_abs_diff": float((stitched - ref).abs().max())}
            srf.close()
    if rank == 0:
        peak, peak_src = hbm_peak()
        bpp = merge_bytes_per_px(N, 2)
        mm = float(np.mean(merge_ms))
        achieved = bpp * ow * oh / (mm / 1e3) / 1e9
        line = {"metric": METRIC, "value": round(out_mp / (ms / 1e3), 2), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(1, min(args.warmup, 3)),
                "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "mode": "rowband",
                "config": {"workload": f"ONE synthetic {W}x{H} RGGB burst, {N} frames, 2x, full frame, split into {world} row band(s)",
                           "frames": N, "raw": [W, H], "scale": 2, "bursts_per_step": 1,
                           "parallelism": f"row bands x{world}: halo rows of the raw frames exchanged between neighbours (the only exchange step; transport in halo.transport), then the unmodified chain per band",
                           "l2": "inputs larger than L2"},
                "halo": {"ms_exchange": round(float(np.mean(halo)), 3), "transport": transport, "rows": rowband.DEFAULT_HALO, "margin_rows": margin,
                         "bytes_received_per_rank": int((b.halo_up + b.halo_down) * W * 2 * N), "band_rows": [x.rows for x in bands],
                         "processed_rows": [x.bottom - x.top for x in bands]},
                "e2e": {"value": round(out_mp / (e2e_ms / 1e3), 2), "unit": UNIT, "ms_per_step": round(e2e_ms, 3), "h2d_bytes_per_step": int(N * H * W * 2),
                        "d2h_bytes_per_step": int(4 * W * H * 3), "result_format": "8-bit RGB", "numa": numa,
                        "timed": "host wall clock: H2D of the rank's own rows from pinned memory + halo exchange + chain + D2H of the band's 8-bit rows; max over ranks"},
                "gpu_launches": int(launches),
                "roofline": {"kernel": "merge_pf_kernel (+ band kernel) of rank 0's band, frame-chunked", "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                             "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src, "bytes_per_output_px": bpp, "ms_per_launch": round(mm, 4),
                             "merge_ms_per_rank": per_rank},
                "stage_ms": {k: round(v, 3) for k, v in sr.stage_ms().items()},
                "clocks": clocks}
        if verify:
            line["verify"] = verify
        print(json.dumps(line))
    sr.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
