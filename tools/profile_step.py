"""Per-launch device-time table of one KITTI forward (CUDA events, see upflow_pytorch_b200/profiler.py).
    python tools/profile_step.py [workload] [precision] > gpurun_out/launch_table.txt"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from upflow_pytorch_b200 import profiler
from upflow_pytorch_b200.engine import DecoderEngine

workload = sys.argv[1] if len(sys.argv) > 1 else "kitti_375x1242_b1"
precision = sys.argv[2] if len(sys.argv) > 2 else "tf32"
H, W, B = bench.WORKLOADS[workload]
sd = bench.load_weights()[0]
eng = DecoderEngine({k: v.cuda() for k, v in sd.items()}, precision=precision)
eng.overlap = False   # single stream: clean per-launch times
im1, im2 = bench.synth_inputs(B, H, W, 1234)
im1, im2 = im1.cuda(), im2.cuda()
with torch.no_grad():
    for _ in range(3):
        eng.forward(im1, im2)
    with profiler.record() as rec:
        eng.forward(im1, im2)
recs = rec.records
tot = sum(r["ms"] for r in recs)
print("# %s %s: %d launches, %.3f ms summed device time (event-pair overhead %.1f us subtracted per launch)" % (
    workload, precision, len(recs), tot, rec.overhead_ms * 1e3))
print("%-4s %-16s %-34s %9s %9s %9s" % ("#", "kernel", "shape", "us", "GB/s", "TFLOP/s"))
for i, r in enumerate(recs):
    name = r.get("kernel") or r["name"]
    print("%-4d %-16s %-34s %9.1f %9.1f %9.2f" % (i, name, str(r["shape"]), r["ms"] * 1e3, r["bytes"] / (r["ms"] * 1e-3) / 1e9,
                                                   r["flops"] / (r["ms"] * 1e-3) / 1e12))
print("# by kernel family")
for k, v in sorted(rec.by_kernel().items(), key=lambda kv: -kv[1]["ms"]):
    print("# %-16s %3d launches %8.1f us  %5.1f%%  %8.2f TFLOP/s %8.1f GB/s" % (k, v["launches"], v["ms"] * 1e3, 100 * v["ms"] / tot,
          v["flops"] / (v["ms"] * 1e-3) / 1e12, v["bytes"] / (v["ms"] * 1e-3) / 1e9))
os.makedirs("gpurun_out", exist_ok=True)
json.dump([{k: v for k, v in r.items()} for r in recs], open("gpurun_out/launch_table_%s_%s.json" % (workload, precision), "w"))
