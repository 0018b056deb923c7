import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upflow_pytorch_b200 import _ext, ops
from upflow_pytorch_b200.ops import Slice
lib = _ext.load()
g = torch.Generator().manual_seed(0)
probe = torch.zeros(8, dtype=torch.int64, device="cuda")
lib.upf_debug_probe(ctypes.c_void_p(probe.data_ptr()))
names = ["prod wait emptyA", "prod wait emptyB", "prod total", "mma wait fullA", "mma wait fullB", "mma total", "epi phase1 tmem->smem", "epi phase2 smem->gmem"]
for mode in (0, 2, 4, 6):
  lib.upf_debug_conv_halo(1, mode | (128 << 8))
  print('mode', mode)
  for (N, h, w, cin, cout) in ((2, 94, 311, 128, 128), (1, 94, 40, 128, 128)):
      X = torch.randn(N, h, w, 576, generator=g).cuda()
      wt = (torch.randn(cout, cin, 3, 3, generator=g) * 0.02).cuda()
      b = torch.zeros(cout).cuda()
      _, wtc = ops.pack_conv_weight(wt, tc=True)
      out = torch.empty(N, h, w, cout, device="cuda")
      for _ in range(3):
          ops.k_conv(Slice(X, 0, cin), wtc, b, out, 3, 1, 1, 0.1, None, _ext.CONV_TF32)
      torch.cuda.synchronize()
      p = probe.cpu().tolist()
      taps = ((cin + 31) // 32) * 9
      print("N%d %dx%d %d->%d (%d taps):" % (N, h, w, cin, cout, taps), ", ".join("%s %d" % (n, v) for n, v in zip(names, p)), "| cycles/tap %.0f" % (p[5] / taps))
