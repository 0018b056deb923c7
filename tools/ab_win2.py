import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from upflow_pytorch_b200 import _ext
from upflow_pytorch_b200.engine import DecoderEngine
lib = _ext.load()
sd = bench.make_weights()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for wl in ("kitti_375x1242_b1", "sintel_436x1024_b8", "hd_1080x1920_b2"):
    H, W, B = bench.WORKLOADS[wl]
    im1, im2 = bench.synth_inputs(B, H, W, 1234)
    for rep in range(2):
        for name, wforce in (("default", 0), ("win two CTAs", 16)):
            lib.upf_debug_conv_win(1, 0, wforce)
            eng = DecoderEngine({k: v.cuda() for k, v in sd.items()}, precision="tf32")
            eng.chain = False
            with torch.no_grad():
                g = eng.capture(B, H, W)
            g.im1.copy_(im1.cuda()); g.im2.copy_(im2.cuda())
            for _ in range(3):
                g.replay()
            ts = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            print("%-20s %-14s median %.3f ms  min %.3f ms" % (wl, name, ts[len(ts) // 2], ts[0]), flush=True)
            del g, eng
            torch.cuda.empty_cache()
lib.upf_debug_conv_win(1, 0, 0)
