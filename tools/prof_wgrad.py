"""Kernel breakdown of one tensor-core weight-gradient call (torch.profiler): python tools/prof_wgrad.py cin cout [ks] [N h w]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from upflow_pytorch_b200 import ops
from upflow_pytorch_b200.ops import Slice
cin, cout = int(sys.argv[1]), int(sys.argv[2])
ks = int(sys.argv[3]) if len(sys.argv) > 3 else 3
N, h, w = (int(v) for v in sys.argv[4:7]) if len(sys.argv) > 6 else (8, 64, 208)
g = torch.Generator().manual_seed(0)
X = torch.randn(N, h, w, (cin + 31) // 32 * 32, generator=g).cuda()
G = torch.randn(N, h, w, (cout + 3) // 4 * 4, generator=g).cuda()
xs, gs = Slice(X, 0, cin), Slice(G, 0, cout)
import os as _os
from upflow_pytorch_b200 import _ext as _e
_e.load().upf_debug_conv_tc(int(_os.environ.get('UPF_TC_DEBUG', '0')))
if _os.environ.get('UPF_TAPS'): _e.load().upf_debug_wgrad_taps(int(_os.environ['UPF_TAPS']))
for _ in range(3):
    ops.k_conv_wgrad(xs, gs, ks, 1, 1, want_bias=True, tensor_cores=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        ops.k_conv_wgrad(xs, gs, ks, 1, 1, want_bias=True, tensor_cores=True)
    torch.cuda.synchronize()
print("wgrad %d->%d k%d at %dx%dx%d, 5 calls:" % (cin, cout, ks, N, h, w))
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
    print("  %8.1f us/call  x%d  %s" % (e.device_time_total / 5, e.count // 5, e.key[:90]))
