"""A/B of the small-grid policy of conv_tc (upf_debug_conv_tc): whole KITTI forward as a CUDA-graph replay, L2 flushed per
step, ONE process.  0 = round-1 heuristic (K split <= 8 only), 8 / 16 = N narrowing + K split with that cluster cap."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from upflow_pytorch_b200 import _ext
from upflow_pytorch_b200.engine import DecoderEngine
lib = _ext.load()
H, W, B = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "kitti_375x1242_b1"]
sd = bench.load_weights()[0]
im1, im2 = bench.synth_inputs(B, H, W, 1234)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
for rep in range(2):
    for cap in (0, 8, 16):
        lib.upf_debug_conv_tc(cap)
        eng = DecoderEngine({k: v.cuda() for k, v in sd.items()}, precision="tf32")
        with torch.no_grad():
            g = eng.capture(B, H, W)
        g.im1.copy_(im1.cuda()); g.im2.copy_(im2.cuda())
        for _ in range(5):
            g.replay()
        ts = []
        for _ in range(30):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        out = g.flow_f.clone()
        if ref is None:
            ref = out
        print("cap %2d  median %.3f ms  min %.3f ms  (%.1f pairs/s)  mean|flow diff vs first| %.3g" % (
            cap, ts[len(ts) // 2], ts[0], B * 1e3 / ts[len(ts) // 2], (out - ref).abs().mean().item()), flush=True)
lib.upf_debug_conv_tc(0)
