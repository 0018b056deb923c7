"""Diagnostic (not a test): per-level differences between the CUDA engine and the CPU port."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cpu_oracle as O, ref_port as P
from upflow_pytorch_b200.engine import DecoderEngine

def run(hw, wseed, precisions=("fp32", "tf32")):
    sd = P.det_state_dict(wseed)
    im1, im2 = O.synthetic_pair(*hw, seed=1234)
    t = time.time()
    with torch.no_grad():
        rf, rb, rflows = P.forward_2_frame(im1, im2, sd)
    print("== size", hw, "wseed", wseed, "cpu port %.2fs" % (time.time() - t), "mean|flow| %.3f" % rf.abs().mean().item())
    for pr in precisions:
        eng = DecoderEngine({k: v.cuda() for k, v in sd.items()}, precision=pr)
        f, b, flows = eng.forward(im1.cuda(), im2.cuda())
        torch.cuda.synchronize()
        for lv, ((a, c), (ra, rc)) in enumerate(zip(flows[::-1], rflows[::-1])):
            d = (a.cpu() - ra).abs()
            print("  %s level %d %s: max %.3g mean %.3g median %.3g  frac>1e-3 %.3f" % (pr, lv, tuple(a.shape[2:]), d.max().item(), d.mean().item(), d.median().item(), (d > 1e-3).float().mean().item()))
        d = (f.cpu() - rf).abs()
        print("  %s full: EPE %.4g max %.3g median %.3g" % (pr, O.epe(f.cpu(), rf), d.max().item(), d.median().item()))

if __name__ == "__main__":
    run((64, 96), 7)
    run((128, 192), 3)
    run((375, 1242), 3)
