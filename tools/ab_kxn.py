"""A/B timing of whole forwards (CUDA-graph replay, L2 flushed per step) for conv_win variants, ONE box, ONE process:
the horizontal taps along N (default) against one MMA per tap, and eight against four epilogue warps.
   python tools/ab_kxn.py [workload ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from upflow_pytorch_b200 import _ext
from upflow_pytorch_b200 import engine as E
from upflow_pytorch_b200.engine import DecoderEngine
lib = _ext.load()
sd = bench.make_weights()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
# name, conv_win mode, conv_win force bits, largest Cout that runs as 1x1-expand + tap-combine
VARIANTS = (("default", 1, 0, E.EXPAND_MAX_COUT), ("conv_win: ring item per channel block", 17, 0, E.EXPAND_MAX_COUT),
            ("per tap, 4 epilogue warps, expand (before)", 5, 32, 8, 16))
if os.environ.get("AB_FULL"):
    VARIANTS += (("kxn, 4 epilogue warps", 1, 32, E.EXPAND_MAX_COUT), ("per tap, 8 warps", 5, 0, E.EXPAND_MAX_COUT))
for wl in (sys.argv[1:] or ["kitti_375x1242_b1"]):
    H, W, B = bench.WORKLOADS[wl]
    im1, im2 = bench.synth_inputs(B, H, W, 1234)
    ref = None
    for rep in range(2):
        for var in VARIANTS:
            name, wmode, wforce, exp_cout = var[:4]
            lib.upf_debug_conv_win(wmode, var[5] if len(var) > 5 else 1003, wforce)
            lib.upf_debug_conv_halo(1, (65 << 16) | (128 << 8) | (var[4] if len(var) > 4 else 0))
            E.EXPAND_MAX_COUT = exp_cout
            eng = DecoderEngine({k: v.cuda() for k, v in sd.items()}, precision="tf32")
            with torch.no_grad():
                g = eng.capture(B, H, W)
            g.im1.copy_(im1.cuda()); g.im2.copy_(im2.cuda())
            for _ in range(3):
                g.replay()
            ts = []
            for _ in range(15):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            out = g.flow_f.clone()
            if ref is None:
                ref = out
            print("%-20s %-34s median %.3f ms  min %.3f ms  mean|flow - first| %.3g" % (
                wl, name, ts[len(ts) // 2], ts[0], (out - ref).abs().mean().item()), flush=True)
            del g, eng
            torch.cuda.empty_cache()
lib.upf_debug_conv_win(1, 0, 0)
lib.upf_debug_conv_halo(1, (65 << 16) | (128 << 8))
