"""A/B: PipelinedInference with 1, 2, 3 compute lanes (graphs of consecutive pairs replaying concurrently).
   python tools/ab_lanes.py [workload] [steps] [lanes,lanes,...]
Prints pairs/s end to end (pinned host in, host out) per lane count and checks that every lane count returns
bit-identical flows for a sequence of different pairs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from upflow_pytorch_b200.pipeline import PipelinedInference

wl = sys.argv[1] if len(sys.argv) > 1 else "kitti_375x1242_b1"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
H, W, B = bench.WORKLOADS[wl]
if os.environ.get("UPF_WIN_DEBUG"):          # "mode,min_cin,force" for upf_debug_conv_win (triage runs)
    from upflow_pytorch_b200 import _ext
    _ext.load().upf_debug_conv_win(*[int(x) for x in os.environ["UPF_WIN_DEBUG"].split(",")])
net, sd, wdesc = bench.build_net(None, os.environ.get("UPF_PRECISION", "tf32"))
pairs = [tuple(t.pin_memory() for t in bench.synth_inputs(B, H, W, 1234 + i)) for i in range(4)]
ref = None
LANES = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [1, 2, 3, 1, 2]
for lanes in LANES:
    pipe = PipelinedInference(net, lanes=lanes)
    got = []
    for i in range(8):
        r = pipe.submit(*pairs[i % 4])
        if r is not None:
            got.append(r.clone())
    got += [r.clone() for r in pipe.drain()]
    assert len(got) == 8
    if ref is None:
        ref = got
    same = all(torch.equal(a, b) for a, b in zip(got, ref))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(K):
        pipe.submit(*pairs[i % 4])
    pipe.drain()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    print("%s lanes=%d  %.3f ms/pair  %.1f pairs/s  bit-identical=%s" % (wl, lanes, dt * 1e3, B / dt, same), flush=True)
