"""fp16 / bf16 storage correlation vs the fp32 tiled and pipelined kernels at the HD 1/4-res shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upflow_pytorch_b200 import ops, _ext
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, iters=14):
    ts = []
    for i in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts)
N, C, h, w, d = 2, 32, 270, 480, 4
g = torch.Generator().manual_seed(1)
f1 = torch.randn(N, h, w, C, generator=g).cuda(); f2 = torch.randn(N, h, w, C, generator=g).cuda()
out = torch.empty(N, h, w, 81, device="cuda")
print("fp32 pipelined: %.1f us" % timed(lambda: ops.k_corr(f1, f2, out, d, slope=0.1)))
_ext.load().upf_debug_corr_pipe(0)
print("fp32 tiled:     %.1f us" % timed(lambda: ops.k_corr(f1, f2, out, d, slope=0.1)))
_ext.load().upf_debug_corr_pipe(1)
for dt in (torch.float16, torch.bfloat16):
    a, b, o = f1.to(dt), f2.to(dt), torch.empty(N, h, w, 81, device="cuda", dtype=dt)
    print("%s storage (tiled): %.1f us" % (dt, timed(lambda: ops.k_corr_lp(a, b, o, d, slope=0.1))))
