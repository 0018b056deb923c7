"""Which convolution calls of one (eager) training step cost what: CUDA events around every ops.k_conv (forward and
input-gradient convolutions) and ops.k_conv_wgrad call, summed by (kind, kernel family, shape).
CAVEAT: the step runs eagerly, so an event pair also spans the host's launch gaps (allocator, tensor-map encoding,
5 launches per weight gradient: ~0.1-0.25 ms per call at this problem size) -- the ranking of the LARGE entries is
usable, the absolute times are upper bounds; the graph-replayed step (bench.py) has none of these gaps.
    python tools/profile_train_convs.py > gpurun_out/train_convs.txt"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import upflow_pytorch_b200 as pkg
from upflow_pytorch_b200 import _ext, ops
from upflow_pytorch_b200.train import Trainer

H, W, B = bench.WORKLOADS["train_256x832_b4"]
net = pkg.build_model(params={"if_use_boundary_warp": False, "multi_scale_distillation_weight": 0.01},
                      state_dict=bench.make_weights(), conv_precision="tf32").train()
tr = Trainer(net, use_cuda_graph=False)
im1, im2 = bench.synth_inputs(B, H, W, 1234)
batch = {"im1": im1.cuda(), "im2": im2.cuda()}
for _ in range(2):
    tr.train_step(batch)
torch.cuda.synchronize()

records = []
orig_conv, orig_wgrad = ops.k_conv, ops.k_conv_wgrad


def timed(kind, fn, shape_of):
    def wrapper(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        fam = _ext.load().upf_last_kernel()
        e1.record()
        records.append((kind, fam.decode() if fam else "?", shape_of(*a, **k), e0, e1))
        return r
    return wrapper


def conv_shape(x, weight, bias, out, ksize, stride=1, dilation=1, slope=0.1, residual=None, precision=0):
    x, out = ops._as_slice(x), ops._as_slice(out)
    return "N%d %dx%d cin%d cout%d k%d s%d d%d%s" % (x.N, x.H, x.W, x.C, out.C, ksize, stride, dilation, " +res" if residual is not None else "")


def wgrad_shape(x, grad_out, ksize, stride=1, dilation=1, want_bias=True, tensor_cores=False, planar=None):
    x, g = ops._as_slice(x), ops._as_slice(grad_out)
    return "N%d %dx%d cin%d cout%d k%d s%d d%d %s%s" % (x.N, x.H, x.W, x.C, g.C, ksize, stride, dilation,
                                                        "tc" if tensor_cores and stride == 1 else "simt", " planar" if planar else "")


ops.k_conv = timed("conv", orig_conv, conv_shape)
ops.k_conv_wgrad = timed("wgrad", orig_wgrad, wgrad_shape)
tr.train_step(batch)
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for kind, fam, shape, e0, e1 in records:
    a = agg[(kind, fam, shape)]
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
by_kind = collections.defaultdict(float)
for (kind, fam, shape), v in agg.items():
    by_kind[kind] += v[1]
print("# %d timed calls, %.1f ms (eager, event-pair overhead ~5 us per call included): %s" % (
    len(records), tot, ", ".join("%s %.1f ms" % kv for kv in by_kind.items())))
for (kind, fam, shape), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print("%6.2f ms %3dx %7.1f us  %-5s %-22s %s" % (t, n, 1e3 * t / n, kind, fam, shape))
