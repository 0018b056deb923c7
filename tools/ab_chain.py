"""A/B timing of the whole forward (CUDA-graph replay, L2 flushed per step): coarse levels as one launch per convolution
against persistent chains (conv_chain.cu) for several cluster sizes / grid sizes.  python tools/ab_chain.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from upflow_pytorch_b200 import _ext
from upflow_pytorch_b200 import engine as E
from upflow_pytorch_b200.engine import DecoderEngine
lib = _ext.load()
H, W, B = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "kitti_375x1242_b1"]
sd = bench.make_weights()
im1, im2 = bench.synth_inputs(B, H, W, 1234)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
# (name, chain on, cluster size | box rows << 8, clusters (0 = all co-resident), max pixels for a chain, conv_tc debug word)
variants = [("per-layer, 32-row boxes", 0, 8, 0, 4096, 0), ("per-layer, 128-row boxes", 0, 8, 0, 4096, 128 << 16),
            ("chain cs8 box32", 1, 8 | (32 << 8), 0, 4096, 0), ("chain cs8 box128", 1, 8 | (128 << 8), 0, 4096, 0),
            ("chain cs8 box16", 1, 8 | (16 << 8), 0, 4096, 0), ("chain cs4 box32", 1, 4 | (32 << 8), 0, 4096, 0),
            ("chain cs8 box32 <=2048 px", 1, 8 | (32 << 8), 0, 2048, 0), ("chain cs8 box32 <=512 px", 1, 8 | (32 << 8), 0, 512, 0)]
ref = None
for rep in range(2):
    for name, on, cs, ncl, maxpx, tcw in variants:
        lib.upf_debug_conv_chain(cs, ncl)
        lib.upf_debug_conv_tc(tcw)
        E.CHAIN_MAX_PIXELS = maxpx
        eng = DecoderEngine({k: v.cuda() for k, v in sd.items()}, precision="tf32")
        eng.chain = bool(on)
        with torch.no_grad():
            g = eng.capture(B, H, W)
        g.im1.copy_(im1.cuda()); g.im2.copy_(im2.cuda())
        for _ in range(5):
            g.replay()
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        out = g.flow_f.clone()
        if ref is None:
            ref = out
        print("%-26s median %.3f ms  min %.3f ms  (%.1f pairs/s)  launches %d  mean|flow diff vs first| %.3g" % (
            name, ts[len(ts) // 2], ts[0], B * 1e3 / ts[len(ts) // 2], g.launches, (out - ref).abs().mean().item()), flush=True)
lib.upf_debug_conv_chain(8, 0)
lib.upf_debug_conv_tc(0)
