"""GPU micro-baseline named by the north star: NPP's FilterBox (the library path the reference's NPP sample, boxFilterNPP.cpp,
stands for) on the tensor-smoothing sub-step, next to the fused kernel-parameter stage of this library.

The reference chain smooths the three structure-tensor planes with a box filter between ComputeStructureTensor (kernel.cu:691)
and ComputeKernelParam (kernel.cu:718) — done by the absent host with NPP.  Here: nppiFilterBox_32f_C1R, 5x5, on three
4032x3024 float planes (what that sub-step alone costs as a library call), against mfsr_stage_kernel_params, which does the
derivative, the tensor, the same 5x5 box and the eigen-analysis in one launch.  Prints one JSON line."""
import ctypes as C, json, sys
sys.path.insert(0, '.')
import torch
from multi_frame_super_resolution_b200 import stages
from multi_frame_super_resolution_b200.pipeline import default_params


class NppiSize(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int)]


class NppiPoint(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int)]


def main():
    npp = None
    for name in ("libnppif.so.12", "/usr/local/cuda/lib64/libnppif.so.12", "libnppif.so"):
        try:
            npp = C.CDLL(name); break
        except OSError:
            continue
    if npp is None:
        print(json.dumps({"unavailable": "libnppif not found"})); return 0
    fn = npp.nppiFilterBox_32f_C1R
    fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, NppiSize, NppiSize, NppiPoint]
    fn.restype = C.c_int
    h, w, r = 3024, 4032, 2
    dev = torch.device("cuda", 0)
    planes = [torch.rand((h, w), device=dev) for _ in range(3)]
    outs = [torch.empty_like(p) for p in planes]
    roi = NppiSize(w - 2 * r, h - 2 * r)

    def npp_box():
        for p, o in zip(planes, outs):
            off = (r * w + r) * 4
            rc = fn(C.c_void_p(p.data_ptr() + off), w * 4, C.c_void_p(o.data_ptr() + off), w * 4, roi, NppiSize(2 * r + 1, 2 * r + 1), NppiPoint(r, r))
            assert rc == 0, rc

    def timed(f, reps=20):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    ms_npp = timed(npp_box)
    # NPP's result against a torch box filter (interior): the baseline computes the same thing
    ref = torch.nn.functional.avg_pool2d(planes[0][None, None], 2 * r + 1, stride=1)[0, 0]
    err = float((outs[0][r:h - r, r:w - r] - ref).abs().max())
    p = default_params()
    gray = torch.rand((h, w), device=dev)
    ms_ours = timed(lambda: stages.kernel_params(gray, p.tensor_box_radius, p.Dth, p.Dtr, p.kDetail, p.kDenoise, p.kStretch, p.kShrink))
    print(json.dumps({"image": [w, h], "box": "5x5", "npp_filterbox_3_planes_ms": round(ms_npp, 4),
                      "npp_bytes_moved": 3 * 2 * h * w * 4, "npp_gbs": round(3 * 2 * h * w * 4 / ms_npp / 1e6, 1), "npp_max_abs_err_vs_torch": err,
                      "mfsr_stage_kernel_params_ms": round(ms_ours, 4),
                      "note": "NPP: box smoothing of three planes only; ours: derivative + tensor + same box + eigen-analysis fused (4 B in, 16 B out per pixel)"}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
