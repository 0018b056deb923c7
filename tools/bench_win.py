"""Validate conv_win.cu (linear-window tensor-core conv) against the per-tap kernel (conv_tc.cu) and the shared-halo
kernel (conv_halo.cu) on the same inputs, time the three, and print CTA 0's pipeline wait counters.
    python tools/bench_win.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from upflow_pytorch_b200 import _ext, ops
from upflow_pytorch_b200.ops import Slice
lib = _ext.load()
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
probe = torch.zeros(8, dtype=torch.int64, device="cuda")
lib.upf_debug_probe(ctypes.c_void_p(probe.data_ptr()))

# name -> (win enabled, force_m, halo enabled)
MODES = (("tap", (0, 0, 0)), ("halo", (0, 0, 1)), ("win9", (7, 0, 0)), ("win", (3, 0, 0)), ("win4w", (3, 32, 0)), ("winMc", (3, 64, 0)), ("winKb", (19, 0, 0)), ("win9 4w", (7, 32, 0)), ("win m1", (3, 1, 0)), ("win m2", (3, 2, 0)), ("win m4", (3, 4, 0)))
if os.environ.get("BW_MODES"):
    MODES = tuple(m for m in MODES if m[0] in os.environ["BW_MODES"].split(",") or m[0] == "tap")


def run(N, h, w, cin, cout, dil, ld=576, resid=False):
    X = torch.randn(N, h, w, ld, generator=g).cuda()
    wt = (torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5).cuda()
    b = (torch.randn(cout, generator=g) * 0.1).cuda()
    r = torch.randn(N, h, w, cout, generator=g).cuda() if resid else None
    _, wtc = ops.pack_conv_weight(wt, tc=True)
    res = {}
    for name, (wen, fm, hen) in MODES:
        lib.upf_debug_conv_win(wen, 0, fm)
        lib.upf_debug_conv_halo(hen, (1 << 16) | (128 << 8))
        out = torch.full((N, h, w, cout + 4), float("nan"), device="cuda")     # 16-byte aligned pitch, like the engine's buffers
        call = lambda: ops.k_conv(Slice(X, 0, cin), wtc, b, Slice(out, 0, cout), 3, 1, dil, 0.1, r, _ext.CONV_TF32)
        probe.zero_()
        lib.upf_debug_probe(ctypes.c_void_p(probe.data_ptr()))
        if os.environ.get('TW_TRACE'): print('  launching', name, flush=True)
        call()
        torch.cuda.synchronize()
        pr = probe.cpu().tolist()
        lib.upf_debug_probe(None)                  # timing runs use the probe-free kernel
        out.fill_(float("nan"))
        call()
        ts = []
        for i in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[name] = (out.clone(), 1e3 * min(ts), pr)
    ref = res["tap"][0]
    fl = 2.0 * N * h * w * 9 * cin * cout
    print("N%d %dx%d %d->%d dil%d%s:" % (N, h, w, cin, cout, dil, " +res" if resid else ""))
    for name, _ in MODES:
        o, us, pr = res[name]
        err = (o[..., :cout] - ref[..., :cout]).abs().max().item()
        untouched = bool(torch.isnan(o[..., cout:]).all())
        print("   %-7s %7.1f us %5.0f TF  err %.2g nan %d pad_ok %d | waits: emptyA %d emptyB %d fullA %d fullB %d | mma total %d epi %d" % (
            name, us, fl / us / 1e6, err, int(torch.isnan(o[..., :cout]).sum()), untouched, pr[0], pr[1], pr[3], pr[4], pr[5], pr[7]), flush=True)


SHAPES = ((2, 94, 311, 576, 128, 1), (2, 94, 311, 544, 32, 1), (2, 94, 311, 480, 64, 1), (2, 94, 311, 384, 96, 1), (2, 94, 311, 576, 2, 1),
          (2, 94, 311, 128, 128, 1), (2, 94, 311, 64, 32, 1), (2, 94, 311, 184, 3, 1), (2, 94, 311, 128, 128, 2), (2, 94, 311, 128, 128, 4),
          (2, 94, 311, 32, 2, 1, 576, True),
          (2, 47, 156, 576, 128, 1), (2, 47, 156, 256, 128, 1), (2, 47, 156, 544, 32, 1), (2, 47, 156, 128, 96, 4),
          (2, 188, 621, 32, 32, 1, 32), (2, 375, 1242, 16, 16, 1, 16), (2, 24, 78, 576, 128, 1), (1, 37, 61, 100, 50, 2, 100))
SHAPES += ((2, 94, 311, 96, 32, 1), (2, 94, 311, 128, 32, 1), (2, 94, 311, 160, 16, 1), (2, 47, 156, 480, 64, 1), (2, 47, 156, 160, 16, 1),
           (2, 188, 621, 16, 32, 1, 16), (2, 94, 311, 64, 32, 2), (2, 94, 311, 64, 48, 4), (1, 37, 61, 100, 50, 1, 100), (1, 33, 45, 40, 20, 1, 44, True))
if len(sys.argv) > 1:
    SHAPES = SHAPES[:int(sys.argv[1])]
if os.environ.get("BW_MAXCOUT"):
    SHAPES = tuple(a for a in SHAPES if a[4] <= int(os.environ["BW_MAXCOUT"]))
for args in SHAPES:
    run(*args)
lib.upf_debug_conv_win(1, 0, 0)
lib.upf_debug_conv_halo(1, (65 << 16) | (128 << 8))
