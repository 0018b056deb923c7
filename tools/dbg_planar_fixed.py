import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upflow_pytorch_b200 import ops, _ext
lib = _ext.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run(n, c, h, w, fl, reps=14):
    g1 = torch.randn(n, c, h, w).cuda(); g2 = torch.randn(n, c, h, w).cuda(); o = torch.empty(n, 81, h, w, device="cuda")
    pp = [(_ext.ctypes.c_longlong * 3)(t.stride(2), t.stride(1), t.stride(0)) for t in (g1, g2, o)]
    ts = []
    for i in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.upf_corr_lrelu_fwd_planar(g1.data_ptr(), pp[0], g2.data_ptr(), pp[1], o.data_ptr(), pp[2], n, h, w, c, 4, 0, 0.1, fl, st)
        e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]
# empty event pair and a trivial torch kernel for scale
ts = []
for i in range(14):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print("empty event pair: median %.1f us" % sorted(ts)[7])
x = torch.zeros(1024, device="cuda"); ts = []
for i in range(14):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); x.add_(1.0); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
print("tiny torch kernel: median %.1f us" % sorted(ts)[7])
for shape in ((1, 32, 4, 120), (1, 32, 148, 480)):
    for fl, name in ((0x2b00, "nothing/no-epi 221 KB smem"), (0x6b00, "nothing/no-epi 16 KB smem")):
        med, best = run(*shape, fl)
        print("shape=%s %-28s median %.1f best %.1f us" % (shape, name, med, best), flush=True)
