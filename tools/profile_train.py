"""Kernel-time breakdown of one training step (BASELINE config 4 shard: 4 pairs of 256x832) with torch.profiler.
    python tools/profile_train.py > gpurun_out/train_breakdown.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import upflow_pytorch_b200 as pkg
from upflow_pytorch_b200.train import Trainer

H, W, B = bench.WORKLOADS["train_256x832_b4"]
net = pkg.build_model(params={"if_use_boundary_warp": False, "multi_scale_distillation_weight": 0.01},
                      state_dict=bench.make_weights(), conv_precision="tf32").train()
tr = Trainer(net)
im1, im2 = bench.synth_inputs(B, H, W, 1234)
batch = {"im1": im1.cuda(), "im2": im2.cuda()}
for _ in range(3):
    tr.train_step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.train_step(batch)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[2] for r in rows)
print("# one training step, %d pairs of %dx%d: %.1f ms of kernel time" % (B, H, W, tot / 1e3))
for k, n, t in sorted(rows, key=lambda r: -r[2])[:45]:
    print("%6.2f%%  %8.2f ms  %5d  %s" % (100 * t / tot, t / 1e3, n, k[:110]))
