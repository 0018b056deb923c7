"""Where the end-to-end step goes beyond the captured graph: graph replay alone, the drop-in call, the pipeline."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from upflow_pytorch_b200.pipeline import PipelinedInference
H, W, B = bench.WORKLOADS["kitti_375x1242_b1"]
net = bench.build_net(None, "tf32")[0].cuda()
im1, im2 = bench.synth_inputs(B, H, W, 1234)
im1_h, im2_h = im1.pin_memory(), im2.pin_memory()
im1_d, im2_d = im1.cuda(), im2.cuda()
K = 50
def timed(fn, name, sync_each=False):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        fn()
        if sync_each: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / K
    print("%-58s device %.3f ms/step   wall %.3f ms/step" % (name, e0.elapsed_time(e1) / K, wall), flush=True)
with torch.no_grad():
    net({"im1": im1_d, "im2": im2_d, "if_loss": False})
    g = list(net._graphs.values())[0]
    timed(lambda: g.replay(), "graph replay, back to back (no L2 flush)")
    timed(lambda: g(im1_d, im2_d), "input copies + graph replay")
    timed(lambda: net({"im1": im1_d, "im2": im2_d, "if_loss": False}), "net(dict), device inputs")
    t0 = time.perf_counter()
    for _ in range(200): net._get_engine()
    print("host: _get_engine %.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
    # host cost of one net() call (no GPU wait): enqueue only
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): net({"im1": im1_d, "im2": im2_d, "if_loss": False})
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print("host: enqueue of one net(dict) call %.1f us" % ((t1 - t0) / 20 * 1e6))
    pipe = PipelinedInference(net)
    def pstep(): pipe.submit(im1_h, im2_h)
    timed(pstep, "pipeline submit (pinned host in, host out)")
    pipe.flush()
