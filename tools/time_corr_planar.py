"""CUDA-event timing of the planar correlation kernel (corr_planar.cu) next to the pixel-major one (corr_pipe.cu) at the
HD 1/4-res shape for d in {2,4,6} (BASELINE config 5 sweep) and at the Sintel / KITTI 1/4-res widths that TMA can walk."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upflow_pytorch_b200 import ops, _ext
import json
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, iters=22):
    ts = []
    for i in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1))
    ts.sort()
    return 1e3 * sum(ts) / len(ts), 1e3 * ts[0]

for (N, C, h, w) in ((2, 32, 270, 480), (8, 32, 109, 256), (2, 32, 94, 312), (2, 64, 47, 156)):
    for d in (2, 4, 6):
        g = torch.Generator().manual_seed(1)
        f1 = torch.randn(N, C, h, w, generator=g).cuda(); f2 = torch.randn(N, C, h, w, generator=g).cuda()
        D2 = (2 * d + 1) ** 2
        outp = torch.empty(N, D2, h, w, device="cuda")
        a, b = ops.to_pixel_major(f1), ops.to_pixel_major(f2)
        outm = torch.empty(N, h, w, D2, device="cuda")
        nbytes = 4 * N * h * w * (2 * C + D2)
        for name, fn in ((("planar", lambda: ops.k_corr_planar(f1, f2, outp, d, slope=0.1)),) if d <= 4 else ()) + (
                         ("pixel-major", lambda: ops.k_corr(a, b, outm, d, slope=0.1)),):
            us, best = timeit(fn)
            print("corr %-11s N%d C%d %dx%d d=%d: %.1f us (best %.1f)  %.0f GB/s  %.1f%% of measured HBM peak  (%.2f TFLOP/s)  [%s]" % (
                name, N, C, h, w, d, us, best, nbytes / us / 1e3, 100 * nbytes / us / 1e3 / peak, 2 * D2 * C * N * h * w / us / 1e6,
                ops.last_kernel() if hasattr(ops, "last_kernel") else ""), flush=True)
        if d <= 4:
            print("   max |planar - pixel-major| = %.3g" % (outp - outm.permute(0, 3, 1, 2)).abs().max().item())
