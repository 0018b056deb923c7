"""CUDA-event timing of the correlation kernel at the HD 1/4-res shape for d in {2,4,6} (BASELINE config 5 sweep)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upflow_pytorch_b200 import ops, _ext
if len(sys.argv) > 1:
    _ext.load().upf_debug_corr_pipe(int(sys.argv[1]))   # 0 = tiled kernel only, n > 1 = pipelined kernel from n tiles on
peak = 6547.8
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (N, C, h, w) in ((2, 32, 270, 480), (2, 32, 94, 311), (2, 64, 47, 156), (2, 196, 6, 20)):
    for d in (2, 4, 6):
        g = torch.Generator().manual_seed(1)
        f1 = torch.randn(N, h, w, C, generator=g).cuda(); f2 = torch.randn(N, h, w, C, generator=g).cuda()
        D2 = (2 * d + 1) ** 2
        out = torch.empty(N, h, w, D2, device="cuda")
        st = torch.zeros(N, C, 2, dtype=torch.float64, device="cuda"); ops.k_stats(f1, st)
        for norm in (False, True):
            ts = []
            for i in range(12):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.k_corr(f1, f2, out, d, st if norm else None, st if norm else None, slope=0.1)
                e1.record(); torch.cuda.synchronize()
                if i >= 2: ts.append(e0.elapsed_time(e1))
            us = 1e3 * sum(ts) / len(ts)
            nbytes = 4 * N * h * w * (2 * C + D2)
            print("corr N%d C%d %dx%d d=%d norm=%d: %.1f us  %.0f GB/s  %.1f%% of measured HBM peak  (%.2f TFLOP/s)" % (
                N, C, h, w, d, norm, us, nbytes / us / 1e3, 100 * nbytes / us / 1e3 / peak, 2 * D2 * C * N * h * w / us / 1e6))
