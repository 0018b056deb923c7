import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upflow_pytorch_b200 import ops, _ext
lib = _ext.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def run(n, c, h, w, fl, do_flush=True, reps=12):
    g1 = torch.randn(n, c, h, w).cuda(); g2 = torch.randn(n, c, h, w).cuda(); o = torch.empty(n, 81, h, w, device="cuda")
    pp = [(_ext.ctypes.c_longlong * 3)(t.stride(2), t.stride(1), t.stride(0)) for t in (g1, g2, o)]
    ts = []
    for i in range(reps):
        if do_flush: flush.zero_()
        else: torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.upf_corr_lrelu_fwd_planar(g1.data_ptr(), pp[0], g2.data_ptr(), pp[1], o.data_ptr(), pp[2], n, h, w, c, 4, 0, 0.1, fl, st)
        e1.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]
HALO = (65 << 16) | (128 << 8)
for pdl in (1, 0):
    lib.upf_debug_conv_halo(1, HALO | (0 if pdl else 8))
    for shape in ((1, 32, 4, 120), (1, 32, 4 * 148 // 4, 480), (2, 32, 270, 480)):
        for fl, name in ((0x2b00, "nothing/no-epi"), (0xb00, "nothing"), (0, "full")):
            for do_flush in (True, False):
                med, best = run(*shape, fl, do_flush)
                print("pdl=%d shape=%s %-15s flush=%d: median %.1f best %.1f us" % (pdl, shape, name, do_flush, med, best), flush=True)
lib.upf_debug_conv_halo(1, HALO)
