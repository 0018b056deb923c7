import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upflow_pytorch_b200 import ops
d = int(sys.argv[1]) if len(sys.argv) > 1 else 4
N, C, H, W = 1, 8, 16, 64
f1 = torch.randn(N, C, H, W).cuda(); f2 = torch.randn(N, C, H, W).cuda()
out = torch.full((N, (2 * d + 1) ** 2, H, W), float("nan"), device="cuda")
ops.k_corr_planar(f1, f2, out, d, slope=1.0)
torch.cuda.synchronize()
a, b = ops.to_pixel_major(f1), ops.to_pixel_major(f2)
o2 = torch.empty(N, H, W, (2 * d + 1) ** 2, device="cuda")
ops.k_corr(a, b, o2, d, slope=1.0)
print("max diff", (out - o2.permute(0, 3, 1, 2)).abs().max().item(), "nan", torch.isnan(out).sum().item())
