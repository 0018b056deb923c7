"""Per-shape A/B of this library's convolutions against cuDNN on the SAME box: every distinct
convolution of the KITTI forward at 1/4 and 1/8 resolution (N = 2: both flow directions stacked) and of the 256x832 b4
training step at the same two levels (N = 8), forward, input gradient (dgrad) and weight gradient (wgrad).
cuDNN side: torch.nn.functional.conv2d / torch.ops.aten.convolution_backward on channels_last tensors, TF32 allowed,
cudnn.benchmark = True (engine search during warm-up).  This side: upf_conv2d_fwd (dispatching conv_win / conv_halo /
conv_tc / the 1x1-expand path as the engine does), the same entry point on flipped-transposed weights for dgrad, and
upf_conv2d_wgrad_tc.  CUDA events, L2 flushed (256 MiB write) before every timed launch, best of 6.
    python tools/conv_vs_cudnn.py > profiles/r2_conv_vs_cudnn.md"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from upflow_pytorch_b200 import _ext, ops
from upflow_pytorch_b200.ops import Slice
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
g = torch.Generator().manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def best_us(fn, reps=6, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)

# (Cin, Cout, ksize, dilation) of the decoder convolutions (model/pwc_modules.py:250-286, 396-412; model/upflow.py:20-69)
EST = [(128, 128, 3, 1), (256, 128, 3, 1), (384, 96, 3, 1), (480, 64, 3, 1), (544, 32, 3, 1), (576, 2, 3, 1)]
CTX = [(576, 128, 3, 1), (128, 128, 3, 2), (128, 128, 3, 4), (128, 96, 3, 8), (96, 64, 3, 16), (64, 32, 3, 1), (32, 2, 3, 1)]
SGU = [(64, 32, 3, 1), (96, 32, 3, 1), (128, 32, 3, 1), (160, 16, 3, 1), (176, 8, 3, 1), (184, 3, 3, 1)]
ADP = [(32, 32, 1, 1), (64, 32, 1, 1)]
LEVELS = [("KITTI 1/4 (2x94x311)", 2, 94, 311, True), ("KITTI 1/8 (2x47x156)", 2, 47, 156, True),
          ("train 1/4 (8x64x208)", 8, 64, 208, False), ("train 1/8 (8x32x104)", 8, 32, 104, False)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    LEVELS = LEVELS[:1]; EST = EST[:2]; CTX = CTX[:1]; SGU = SGU[:1]; ADP = ADP[:1]
if len(sys.argv) > 1 and sys.argv[1] == "train":
    LEVELS = LEVELS[2:]

print("# Convolutions: this library vs cuDNN on the same B200 (tools/conv_vs_cudnn.py)\n")
print("cuDNN %s, torch %s, TF32 allowed, `cudnn.benchmark=True`, channels_last; CUDA events, L2 flushed before every launch, best of 6; µs.\n" % (torch.backends.cudnn.version(), torch.__version__))
wins = {"fwd": [0, 0], "dgrad": [0, 0], "wgrad": [0, 0]}
for lname, N, h, w, inference in LEVELS:
    print("## %s\n" % lname)
    print("| conv (Cin→Cout, k, dil) | fwd upf | fwd cuDNN | kernel | dgrad upf | dgrad cuDNN | wgrad upf | wgrad cuDNN |")
    print("|---|---|---|---|---|---|---|---|")
    for (cin, cout, ks, dil) in EST + CTX + SGU + ADP:
        ld = (cin + 31) // 32 * 32
        X = torch.randn(N, h, w, ld, generator=g).cuda()
        wt = (torch.randn(cout, cin, ks, ks, generator=g) * (2.0 / (cin * ks * ks)) ** 0.5).cuda()
        b = (torch.randn(cout, generator=g) * 0.1).cuda()
        ldo = (cout + 3) // 4 * 4
        out = torch.empty(N, h, w, ldo, device="cuda")
        _, wtc = ops.pack_conv_weight(wt, tc=True, tc_only=True)
        xs, os_ = Slice(X, 0, cin), Slice(out, 0, cout)
        f_upf = best_us(lambda: ops.k_conv(xs, wtc, b, os_, ks, 1, dil, 0.1, None, _ext.CONV_TF32))
        kern = ops.last_kernel()
        if ks == 3 and cout <= 8 and h * w >= 5000:
            # the engine's route for few-output-channel 3x3 convolutions: 1x1 expansion + tap combine
            we = ops.expand_taps_weight(wt)
            _, wetc = ops.pack_conv_weight(we, tc=True, tc_only=True)
            y = torch.empty(N, h, w, (9 * cout + 3) // 4 * 4, device="cuda")
            zb = torch.zeros(9 * cout, device="cuda")
            def expand():
                ops.k_conv(xs, wetc, zb, Slice(y, 0, 9 * cout), 1, 1, 1, 1.0, None, _ext.CONV_TF32)
                ops.k_tap_combine(Slice(y, 0, 9 * cout), b, os_, dil, 0.1, None)
            t2 = best_us(expand)
            if t2 < f_upf:
                f_upf, kern = t2, "1x1-expand + tap_combine"
        xc = X[..., :cin].permute(0, 3, 1, 2)            # NCHW view with channels_last strides (pitch ld)
        xcl = xc.contiguous(memory_format=torch.channels_last)
        wcl = wt.contiguous(memory_format=torch.channels_last)
        pad = dil * (ks - 1) // 2
        f_cud = best_us(lambda: F.leaky_relu(F.conv2d(xcl, wcl, b, 1, pad, dil), 0.1, inplace=True))
        row = "| %d→%d, %d, %d | %.1f | %.1f | %s |" % (cin, cout, ks, dil, f_upf, f_cud, kern)
        wins["fwd"][0] += f_upf <= f_cud; wins["fwd"][1] += 1
        if not inference:
            gout = torch.randn(N, h, w, ldo, generator=g).cuda()
            gs = Slice(gout, 0, cout)
            _, wft = ops.pack_conv_weight(wt, tc=True, flip_transpose=True, tc_only=True)
            gx = torch.empty(N, h, w, (cin + 3) // 4 * 4, device="cuda")
            zb = torch.zeros(cin, device="cuda")
            d_upf = best_us(lambda: ops.k_conv(gs, wft, zb, Slice(gx, 0, cin), ks, 1, dil, 1.0, None, _ext.CONV_TF32))
            w_upf = best_us(lambda: ops.k_conv_wgrad(xs, gs, ks, 1, dil, want_bias=True, tensor_cores=True))
            gcl = gout[..., :cout].permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
            conv_bwd = torch.ops.aten.convolution_backward
            d_cud = best_us(lambda: conv_bwd(gcl, xcl, wcl, [cout], [1, 1], [pad, pad], [dil, dil], False, [0, 0], 1, [True, False, False]))
            w_cud = best_us(lambda: conv_bwd(gcl, xcl, wcl, [cout], [1, 1], [pad, pad], [dil, dil], False, [0, 0], 1, [False, True, True]))
            row += " %.1f | %.1f | %.1f | %.1f |" % (d_upf, d_cud, w_upf, w_cud)
            wins["dgrad"][0] += d_upf <= d_cud; wins["dgrad"][1] += 1
            wins["wgrad"][0] += w_upf <= w_cud; wins["wgrad"][1] += 1
        else:
            row += " | | | |"
        print(row, flush=True)
    print()
print("Shapes at or ahead of cuDNN: " + ", ".join("%s %d/%d" % (k, v[0], v[1]) for k, v in wins.items()))
print("\n(the cuDNN forward includes its separate in-place LeakyReLU launch, which this library fuses into the epilogue; cuDNN's wgrad\nincludes the bias gradient, like `upf_conv2d_wgrad_tc`)")
