"""A/B timing of the whole KITTI forward (CUDA-graph replay, L2 flushed per step): one resident CTA per SM with deep rings
(default) against two resident CTAs with ~108 KB rings for conv_halo / conv_win.  python tools/ab_two_cta.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from upflow_pytorch_b200 import _ext
from upflow_pytorch_b200.engine import DecoderEngine
lib = _ext.load()
H, W, B = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "kitti_375x1242_b1"]
sd = bench.make_weights()
im1, im2 = bench.synth_inputs(B, H, W, 1234)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
HALO = (65 << 16) | (128 << 8)
# (name, halo `enabled` word (bit 1 = two CTAs), win force_m word (16 = two CTAs))
# halo word: bit 1 = two CTAs per SM, bit 2 = weight multicast OFF, bit 3 = ONE MMA issuer; win word: 16 = ONE CTA per SM
# bits 4..7 of the halo word = TMA boxes per halo tile
variants = [("default", 1, 0), ("halo A in 2 boxes", 1 | (2 << 4), 0), ("halo one issuer", 9, 0), ("halo multicast off", 5, 0), ("win one CTA", 1, 16)]
ref = None
for rep in range(2):
    for name, hen, wforce in variants:
        lib.upf_debug_conv_win(1, 0, wforce)
        lib.upf_debug_conv_halo(hen, HALO)
        eng = DecoderEngine({k: v.cuda() for k, v in sd.items()}, precision="tf32")
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
        print("%-18s median %.3f ms  min %.3f ms  (%.1f pairs/s)  max|flow diff vs first| %.3g" % (
            name, ts[len(ts) // 2], ts[0], B * 1e3 / ts[len(ts) // 2], (out - ref).abs().max().item()), flush=True)
lib.upf_debug_conv_win(1, 0, 0)
lib.upf_debug_conv_halo(1, HALO)
