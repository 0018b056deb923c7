"""Tensor-core weight gradient vs the fp32 SIMT one over the training step's convolution shapes at every pyramid level."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upflow_pytorch_b200 import ops
from upflow_pytorch_b200.ops import Slice
g = torch.Generator().manual_seed(0)
convs = [(3, 16, 3, 1), (16, 16, 3, 1), (16, 32, 3, 1), (32, 32, 3, 1), (32, 32, 1, 1), (64, 32, 1, 1), (64, 32, 3, 1), (32, 2, 3, 1),
         (565, 128, 3, 1), (128, 128, 3, 2), (128, 128, 3, 4), (128, 96, 3, 8), (96, 64, 3, 16), (96, 32, 1, 1), (128, 32, 1, 1), (196, 32, 1, 1),
         (115, 128, 3, 1), (243, 128, 3, 1), (371, 96, 3, 1), (467, 64, 3, 1), (531, 32, 3, 1), (563, 2, 3, 1), (160, 16, 3, 1), (184, 3, 3, 1)]
levels = [(4, 256, 832), (4, 128, 416), (8, 64, 208), (8, 32, 104), (8, 16, 52), (8, 8, 26), (8, 4, 13)]
worst = 0.0
for (N, h, w) in levels:
    for (cin, cout, ks, dil) in convs:
        if (h, w) == (256, 832) and cin > 16: continue
        if (h, w) == (128, 416) and cin > 32: continue
        X = torch.randn(N, h, w, (cin + 3) // 4 * 4, generator=g).cuda()
        G = torch.randn(N, h, w, (cout + 3) // 4 * 4, generator=g).cuda()
        xs, gs = Slice(X, 0, cin), Slice(G, 0, cout)
        try:
            gw, gb = ops.k_conv_wgrad(xs, gs, ks, 1, dil, want_bias=True, tensor_cores=True)
            torch.cuda.synchronize()
        except Exception as e:
            print("FAIL", (N, h, w), (cin, cout, ks, dil), repr(e)[:200], flush=True)
            sys.exit(1)
        rw, rb = ops.k_conv_wgrad(xs, gs, ks, 1, dil, want_bias=True, tensor_cores=False)
        rel = ((gw - rw).norm() / rw.norm().clamp_min(1e-20)).item()
        worst = max(worst, rel)
        if not rel < 5e-3:
            print("MISMATCH", (N, h, w), (cin, cout, ks, dil), rel, flush=True)
print("all shapes ran; worst relative L2 difference TC (TF32) vs SIMT (fp32): %.3g" % worst)
