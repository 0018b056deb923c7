"""Run ONE kernel configuration a few times (target for `ncu --set full -k regex:...`).
    python tools/run_kernel.py corr|corr_planar|conv [args]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from upflow_pytorch_b200 import _ext, ops
from upflow_pytorch_b200.ops import Slice

what = sys.argv[1]
g = torch.Generator().manual_seed(0)
if what == "corr":
    N, C, h, w, d = 2, 32, 270, 480, 4
    f1 = torch.randn(N, h, w, C, generator=g).cuda()
    f2 = torch.randn(N, h, w, C, generator=g).cuda()
    out = torch.empty(N, h, w, 81, device="cuda")
    for _ in range(3):
        ops.k_corr(f1, f2, out, d, slope=0.1)
elif what == "corr_planar":
    N, C, h, w, d = 2, 32, 270, 480, int(sys.argv[2]) if len(sys.argv) > 2 else 4
    f1 = torch.randn(N, C, h, w, generator=g).cuda()
    f2 = torch.randn(N, C, h, w, generator=g).cuda()
    out = torch.empty(N, (2 * d + 1) ** 2, h, w, device="cuda")
    for _ in range(3):
        ops.k_corr_planar(f1, f2, out, d, slope=0.1)
elif what == "conv":
    # estimator conv2 at KITTI 1/4 res: X[0:256] -> 128 channels, both directions stacked
    N, h, w = 2, int(sys.argv[5]) if len(sys.argv) > 5 else 94, int(sys.argv[6]) if len(sys.argv) > 6 else 311
    cin, cout = int(sys.argv[2]) if len(sys.argv) > 2 else 256, int(sys.argv[3]) if len(sys.argv) > 3 else 128
    dil = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    X = torch.randn(N, h, w, 576, generator=g).cuda()
    wt = torch.randn(cout, cin, 3, 3, generator=g).cuda() * 0.02
    b = torch.zeros(cout).cuda()
    _, wtc = ops.pack_conv_weight(wt, tc=True)
    out = torch.empty(N, h, w, cout, device="cuda")
    for _ in range(3):
        ops.k_conv(Slice(X, 0, cin), wtc, b, out, 3, 1, dil, 0.1, None, _ext.CONV_TF32)
torch.cuda.synchronize()
print("done", what)
