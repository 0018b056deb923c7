"""Per-role cycle counters of CTA 0 (upf_debug_probe) for the fine-level convolution kernels at the 1/4-res KITTI shape."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upflow_pytorch_b200 import _ext, ops
from upflow_pytorch_b200.ops import Slice
lib = _ext.load()
g = torch.Generator().manual_seed(0)
probe = torch.zeros(64, dtype=torch.int64, device="cuda")
names = ["prod wait emptyA", "prod wait emptyB", "prodB total", "mma wait fullA", "mma wait fullB", "mma total", "epi ph1", "epi ph2/total"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (N, h, w, cin, cout) in ((2, 94, 311, 128, 128), (2, 94, 311, 576, 128), (2, 94, 311, 384, 96), (2, 94, 311, 544, 32), (2, 94, 311, 480, 64), (2, 94, 311, 128, 32), (2, 94, 311, 64, 32)):
    X = torch.randn(N, h, w, 576, generator=g).cuda()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) * 0.02).cuda()
    b = torch.zeros(cout).cuda()
    _, wtc = ops.pack_conv_weight(wt, tc=True)
    out = torch.empty(N, h, w, cout, device="cuda")
    lib.upf_debug_probe(None)
    for _ in range(3):
        ops.k_conv(Slice(X, 0, cin), wtc, b, out, 3, 1, 1, 0.1, None, _ext.CONV_TF32)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.k_conv(Slice(X, 0, cin), wtc, b, out, 3, 1, 1, 0.1, None, _ext.CONV_TF32); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    kern = lib.upf_last_kernel().decode()
    probe.zero_()
    lib.upf_debug_probe(ctypes.c_void_p(probe.data_ptr()))
    ops.k_conv(Slice(X, 0, cin), wtc, b, out, 3, 1, 1, 0.1, None, _ext.CONV_TF32)
    torch.cuda.synchronize()
    lib.upf_debug_probe(None)
    p = probe.cpu().tolist()[:8]
    taps = ((cin + 31) // 32) * 9
    print("%-10s N%d %dx%d %d->%d: %.1f us | " % (kern, N, h, w, cin, cout, min(ts)) + ", ".join("%s %.2fus" % (n, v / 1965.0) for n, v in zip(names, p)) + " | mma cycles/tap %.0f" % (p[5] / taps), flush=True)
