"""Validate conv_halo.cu against the per-tap tensor-core kernel (conv_tc.cu) for both descriptor base-offset
conventions, and time both.  python tools/bench_halo.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upflow_pytorch_b200 import _ext, ops
from upflow_pytorch_b200.ops import Slice
lib = _ext.load()
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

MODES = (("tap", (0, 128 << 8)), ("halo", (1, 128 << 8)))


def run(N, h, w, cin, cout, dil, ld=576):
    X = torch.randn(N, h, w, ld, generator=g).cuda()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5).cuda()
    b = (torch.randn(cout, generator=g) * 0.1).cuda()
    _, wtc = ops.pack_conv_weight(wt, tc=True)
    res = {}
    for name, (en, bo) in MODES:
        lib.upf_debug_conv_halo(en, bo)
        out = torch.full((N, h, w, cout), float("nan"), device="cuda")
        ops.k_conv(Slice(X, 0, cin), wtc, b, out, 3, 1, dil, 0.1, None, _ext.CONV_TF32)
        torch.cuda.synchronize()
        ts = []
        for i in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.k_conv(Slice(X, 0, cin), wtc, b, out, 3, 1, dil, 0.1, None, _ext.CONV_TF32); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[name] = (out.clone(), 1e3 * min(ts))
    ref = res[MODES[0][0]][0]
    fl = 2.0 * N * h * w * 9 * cin * cout
    msg = "N%d %dx%d %d->%d dil%d:" % (N, h, w, cin, cout, dil)
    for name, _ in MODES:
        o, us = res[name]
        err = (o - ref).abs().max().item()
        msg += "  [%s %.1f us %.0f TF err %.2g nan %d]" % (name, us, fl / us / 1e6, err, int(torch.isnan(o).sum()))
    print(msg, flush=True)

for args in ((2, 94, 311, 64, 32, 1), (2, 94, 311, 128, 128, 1), (2, 94, 311, 576, 128, 1), (2, 94, 311, 544, 32, 1), (2, 94, 311, 576, 2, 1),
             (2, 94, 311, 128, 128, 2), (2, 94, 311, 128, 128, 4), (2, 47, 156, 256, 128, 1), (2, 47, 156, 128, 96, 4),
             (2, 188, 621, 32, 32, 1, 32), (2, 375, 1242, 16, 16, 1, 16), (2, 24, 78, 576, 128, 1), (2, 6, 20, 576, 128, 1),
             (2, 12, 39, 128, 128, 2)):
    run(*args)
lib.upf_debug_conv_halo(1, 128 << 8)
