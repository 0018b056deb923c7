"""Where a layer of a persistent chain spends its time: clock64 stamps of CTA 0's roles (upf_debug_conv_chain_probe) for the
estimator + context chain of one coarse level.  python tools/probe_chain.py [h w]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
import bench
from upflow_pytorch_b200 import _ext, ops
from upflow_pytorch_b200.engine import DecoderEngine
lib = _ext.load()
H, W, B = bench.WORKLOADS["kitti_375x1242_b1"]
sd = bench.make_weights()
eng = DecoderEngine({k: v.cuda() for k, v in sd.items()}, precision="tf32")
eng.overlap = False
im1, im2 = bench.synth_inputs(B, H, W, 1234)
im1, im2 = im1.cuda(), im2.cuda()
probe = torch.zeros(256, dtype=torch.int64, device="cuda")
calls = []
orig = ops.k_conv_chain
def hooked(layers):
    probe.zero_()
    lib.upf_debug_conv_chain_probe(ctypes.c_void_p(probe.data_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(layers); e1.record()
    torch.cuda.synchronize()
    lib.upf_debug_conv_chain_probe(None)
    calls.append((layers[0]._shape, [(L.Cin, L.Cout, L.dilation) for L in layers], probe.cpu().view(16, 16).clone(), e0.elapsed_time(e1) * 1e3))
with torch.no_grad():
    for _ in range(2):
        eng.forward(im1, im2)
    ops.k_conv_chain = hooked
    eng.forward(im1, im2)
    ops.k_conv_chain = orig
names = ["at barrier", "barrier passed", "A loads out", "operands landed", "last MMA out", "accum complete", "parked", "cluster parked",
         "stored", "items done", "fenced", "arrived"]
GHZ = 1.965
for shape, layers, p, us in calls:
    print("chain %s: %d layers, %.1f us by events" % (shape, len(layers), us))
    t0 = None
    for l, (cin, cout, dil) in enumerate(layers):
        row = p[l, :12].tolist()
        base = row[1] if row[1] else row[3]
        if t0 is None:
            t0 = base
        prev_arr = p[l - 1, 11].item() if l else base
        txt = "  L%-2d %3d->%-3d d%-2d start %7.2f us |" % (l, cin, cout, dil, (base - t0) / GHZ / 1e3)
        for k in range(12):
            txt += " %s %+.2f" % (names[k].split()[0][:6], (row[k] - base) / GHZ / 1e3) if row[k] else " %s   -  " % names[k].split()[0][:6]
        print(txt)
