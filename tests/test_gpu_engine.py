"""The fused two-frame decoder (engine.py) and the drop-in UPFlow_net against the
reference: golden end-to-end fixture (reference executed on CPU), the op-for-op
port at KITTI size, and module-by-module decode_level_res.  -m gpu.

Pointwise end-to-end equality with the reference is NOT attainable by any
re-associated implementation: WarpingLayer_no_div's `mask >= 1.0`
(model/pwc_modules.py:206) zeroes ~1.5 % of interior pixels on 1-ulp
differences, which pixels is re-randomised by any 1e-7 change of the flow, and
every flipped pixel moves the per-channel normalisation statistics of the whole
image.  Measured on the reference's own CPU path (oracle/ref_port.py, weights
det_state_dict(3), inputs scaled by 1+1e-6): mean EPE 0.056 px at 375x1242,
0.019 px at 128x192.  So the tests pin
  (1) every level BEFORE the first warp to rounding (fp32: 1e-6),
  (2) the full pipeline with the mask threshold relaxed to 0.9999 on BOTH sides
      ("robust mask" diagnostic) to 2e-4 px mean EPE in fp32 -- this is the
      statement that nothing but the mask discontinuity separates the two,
  (3) the reference-semantics pipeline to the reference's own noise floor.
"""
import pytest
import torch

from oracle import cpu_oracle as O
from oracle import ref_port as P

pytestmark = pytest.mark.gpu

NOISE_FLOOR_EPE = 0.15     # 375x1242, random-init weights: reference vs itself (1+1e-6) = 0.056 px


def _engine(precision, sd, **kw):
    from upflow_pytorch_b200.engine import DecoderEngine
    return DecoderEngine({k: v.cuda() for k, v in sd.items()}, precision=precision, **kw)


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_golden_e2e(golden, precision):
    g = golden("e2e")
    sd = P.det_state_dict(g["wseed"])
    im1, im2 = O.synthetic_pair(*g["hw"], seed=g["pair_seed"])
    eng = _engine(precision, sd)
    f, b, flows = eng.forward(im1.cuda(), im2.cuda())
    f, b = f.cpu(), b.cpu()
    flows = [[x.cpu(), y.cpu()] for x, y in flows]
    tol0 = 5e-6 if precision == "fp32" else 1e-3
    # the three coarsest levels of this fixture see no mask flip: rounding only
    for lv in (-1, -2):
        for d in (0, 1):
            assert (flows[lv][d] - g["flows"][lv][d]).abs().max().item() <= tol0
    epe_f, epe_b = O.epe(f, g["flow_f_out"]), O.epe(b, g["flow_b_out"])
    print("golden e2e", precision, "EPE fw/bw vs reference", epe_f, epe_b)
    assert epe_f <= 0.05 and epe_b <= 0.05


@pytest.mark.parametrize("hw", [(375, 1242), (128, 192)])
def test_kitti_size_vs_port(hw):
    """full size, random-init deterministic weights: CUDA vs the CPU port of the reference."""
    sd = P.det_state_dict(3)
    im1, im2 = O.synthetic_pair(*hw, seed=1234)
    with torch.no_grad():
        rf, rb, rflows = P.forward_2_frame(im1, im2, sd)
    for precision in ("fp32", "tf32"):
        eng = _engine(precision, sd)
        f, b, flows = eng.forward(im1.cuda(), im2.cuda())
        torch.cuda.synchronize()
        epe = O.epe(f.cpu(), rf)
        l0 = (flows[-1][0].cpu() - rflows[-1][0]).abs().max().item()
        print("size", hw, precision, "EPE vs port", epe, "level0 max diff", l0, "ref mean |flow|", rf.abs().mean().item())
        assert l0 <= (2e-6 if precision == "fp32" else 2e-3)
        assert epe <= NOISE_FLOOR_EPE
        # EPE against the known synthetic motion (-3,+2) is not meaningful for random weights; the
        # checkpointed comparison is reported by bench/epe_report.py when the checkpoint is present


@pytest.mark.parametrize("hw", [(375, 1242), (128, 192)])
def test_robust_mask_diagnostic_is_tight(hw):
    """mask threshold 0.9999 on both sides: the whole forward agrees to rounding (fp32) / TF32 noise."""
    sd = P.det_state_dict(3)
    im1, im2 = O.synthetic_pair(*hw, seed=1234)
    P.MASK_THRESHOLD = 0.9999
    try:
        with torch.no_grad():
            rf, rb, rflows = P.forward_2_frame(im1, im2, sd)
    finally:
        P.MASK_THRESHOLD = 1.0
    for precision, bound in (("fp32", 2e-4), ("tf32", 6e-3)):    # measured: fp32 4e-6 px, tf32 1.9e-3 px (random weights; 0.027 px with truncated operands)
        eng = _engine(precision, sd, mask_threshold=0.9999)
        f, b, flows = eng.forward(im1.cuda(), im2.cuda())
        epe_f, epe_b = O.epe(f.cpu(), rf), O.epe(b.cpu(), rb)
        mx = (f.cpu() - rf).abs().max().item()
        print("robust-mask", hw, precision, "EPE fw %.3g bw %.3g max|diff| %.3g" % (epe_f, epe_b, mx))
        assert epe_f <= bound and epe_b <= bound


def test_dropin_decode_level_golden(golden):
    """decode_level_res through the drop-in modules, teacher-forced with the reference's inputs (golden).
    flow_*_up (SGU output: the warp inside it sees bit-identical inputs) must be tight; the residuals pass through a
    warp by that (1e-6-perturbed) flow, i.e. through re-randomised mask flips on a 12x16 map."""
    import upflow_pytorch_b200 as pkg
    g = golden("decode_level")
    net = pkg.build_model(state_dict=P.det_state_dict(g["wseed"]), conv_precision="fp32")
    from model import pwc_modules
    pwc_modules.set_conv_precision("fp32")
    c = lambda k: g[k].cuda()
    with torch.no_grad():
        o = net.decode_level_res(level=g["level"], flow_1=c("flow_1"), flow_2=c("flow_2"), feature_1=c("x1"),
                                 feature_1_1x1=c("a1"), feature_2=c("x2"), feature_2_1x1=c("a2"), img_ori_1=None,
                                 img_ori_2=None)
    for got, key in zip(o, ("flow_1_up", "flow_2_up", "res_1", "res_2")):
        d = (got.cpu() - g[key]).abs()
        print("decode_level", key, "max", d.max().item(), "median", d.median().item())
        if key.startswith("flow"):
            assert d.max().item() <= 5e-5
        else:
            assert d.mean().item() <= 5e-2


def test_dropin_sgu_golden(golden):
    import upflow_pytorch_b200 as pkg
    s = golden("sgu")
    net = pkg.build_model(state_dict=P.det_state_dict(s["wseed"]), conv_precision="fp32")
    from model import pwc_modules
    pwc_modules.set_conv_precision("fp32")
    with torch.no_grad():
        _, up, iflow, imask = net.sgi_model(s["flow"].cuda(), s["f1"].cuda(), s["f2"].cuda())
        _, up2, iflow2, imask2 = net.sgi_model(s["flow"].cuda(), s["f1"].cuda(), s["f2"].cuda(),
                                               output_level_flow=s["output_level_flow"].cuda())
    assert (up.cpu() - s["flow_up"]).abs().max().item() <= 1e-4
    assert (iflow.cpu() - s["inter_flow"]).abs().max().item() <= 1e-4
    assert (imask.cpu() - s["inter_mask"]).abs().max().item() <= 1e-4
    assert (up2.cpu() - s["flow_up_out"]).abs().max().item() <= 1e-4
    assert (imask2.cpu() - s["inter_mask_out"]).abs().max().item() <= 1e-4


def test_dropin_net_forward_matches_reference_api(golden):
    """net(input_dict) -> output_dict with the reference's keys; flows within the noise floor of the golden run."""
    import upflow_pytorch_b200 as pkg
    g = golden("e2e")
    net = pkg.build_model(state_dict=P.det_state_dict(g["wseed"]), conv_precision="fp32")
    im1, im2 = O.synthetic_pair(*g["hw"], seed=g["pair_seed"])
    with torch.no_grad():
        out = net({"im1": im1.cuda(), "im2": im2.cuda(), "if_loss": False})
    assert set(out) == {"flow_f_out", "flow_b_out", "occ_fw", "occ_bw"}
    assert out["flow_f_out"].shape == g["flow_f_out"].shape
    assert O.epe(out["flow_f_out"].cpu(), g["flow_f_out"]) <= 0.05
    agree = (out["occ_fw"].cpu() == g["occ_fw"]).float().mean().item()
    assert agree >= 0.97
    with pytest.raises(RuntimeError):
        net.forward_2_frame_v3(im1, im2)          # CPU tensors: no fallback


def test_batch_and_repeatability():
    sd = P.det_state_dict(5)
    eng = _engine("fp32", sd)
    im1, im2 = O.synthetic_pair(96, 160, seed=7, batch=3)
    f, b, _ = eng.forward(im1.cuda(), im2.cuda())
    f1, b1 = f.clone(), b.clone()
    f, b, _ = eng.forward(im1.cuda(), im2.cuda())
    assert torch.equal(f, f1) and torch.equal(b, b1)             # deterministic, workspace reuse is clean
    # image 1 of the batch alone gives the same flow: per-image independence (multi-GPU sharding relies on it)
    fs, bs, _ = eng.forward(im1[1:2].cuda(), im2[1:2].cuda())
    assert (fs - f1[1:2]).abs().max().item() <= 1e-4
    # swapping the two frames swaps forward and backward flow
    fw, bw, _ = eng.forward(im2.cuda(), im1.cuda())
    assert (fw - b1).abs().max().item() <= 1e-4


@pytest.mark.parametrize("mode", ["obj", "all", "out"])
def test_fused_occlusion_check_vs_module(mode):
    """upf_occ_check (one launch inside the graph) against tools.occ_check_model of the drop-in (the reference's
    elementwise recipe, utils/tools.py:550-588, 641-677, around the library warp) on the same flows."""
    import upflow_pytorch_b200
    from upflow_pytorch_b200 import ops
    upflow_pytorch_b200.install_dropin()
    from utils.tools import tools
    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 37, 53
    ff, fb = torch.randn(B, 2, H, W, generator=g).cuda() * 4, torch.randn(B, 2, H, W, generator=g).cuda() * 4
    fb = -ff + fb * 0.05                                   # mostly consistent, so both classes occur
    ref_f, ref_b = tools.occ_check_model(occ_type='for_back_check', occ_alpha_1=0.1, occ_alpha_2=0.5, obj_out_all=mode)(ff, fb)
    stacked = torch.cat([ff, fb]).permute(0, 2, 3, 1).contiguous()
    occ = torch.empty(2 * B, H, W, 1, device="cuda")
    ops.k_occ_check(stacked, occ, 0.1, 0.5, mode)
    got_f, got_b = occ[:B].permute(0, 3, 1, 2), occ[B:].permute(0, 3, 1, 2)
    assert 0.02 < ref_f.mean().item() < 0.98
    assert (got_f != ref_f).float().mean().item() <= 1e-3 and (got_b != ref_b).float().mean().item() <= 1e-3


def test_tf32x3_is_fp32_class():
    """precision='tf32x3' (three TF32 tensor-core passes per convolution, hi*hi + lo*hi + hi*lo) against the strict
    fp32 SIMT engine and the CPU port, robust-mask diagnostic on all sides so that the `mask >= 1.0` flips do not
    hide the arithmetic: mean EPE <= 1e-3 px (the north star's bound), where plain TF32 sits at ~3e-2."""
    sd = P.det_state_dict(3)
    im1, im2 = O.synthetic_pair(128, 192, seed=1234)
    P.MASK_THRESHOLD = 0.9999
    try:
        with torch.no_grad():
            rf, rb, _ = P.forward_2_frame(im1, im2, sd)
    finally:
        P.MASK_THRESHOLD = 1.0
    res = {}
    for prec in ("fp32", "tf32x3", "tf32"):
        eng = _engine(prec, sd, mask_threshold=0.9999)
        f, b, _ = eng.forward(im1.cuda(), im2.cuda())
        res[prec] = (f.cpu().clone(), b.cpu().clone())
    for prec in ("fp32", "tf32x3", "tf32"):
        print("  %-7s mean EPE vs CPU port %.3g / %.3g px, vs fp32 engine %.3g px" % (
            prec, O.epe(res[prec][0], rf), O.epe(res[prec][1], rb), O.epe(res[prec][0], res["fp32"][0])))
    assert O.epe(res["tf32x3"][0], rf) <= 1e-3 and O.epe(res["tf32x3"][1], rb) <= 1e-3
    assert O.epe(res["tf32x3"][0], res["fp32"][0]) <= 1e-3


# ------------------------------------------------------------------ the shipped checkpoint (BASELINE config 2)
def _checkpoint_net(precision):
    """the drop-in UPFlow_net with scripts/upflow_kitti2015.pth (fixture copy) loaded the way test.py:31-38 does"""
    import os
    import upflow_pytorch_b200 as pkg
    GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    net = pkg.build_model(conv_precision=precision)
    net.load_model(os.path.join(GOLDEN, "upflow_kitti2015.pth"), if_relax=True, if_print=False)
    return net.cuda().eval()


@pytest.mark.parametrize("size", [0, 1])
def test_checkpoint_epe_vs_reference_at_full_size(golden, size):
    """North-star parity (EPE within 1e-3 px of the reference) with the SHIPPED weights at 375x1242 and 436x1024:
    every precision against the reference's own CPU flow.  Robust-mask diagnostic (threshold 0.9999 on both sides):
    fp32 and tf32x3 to rounding, tf32 (operands rounded to nearest) inside 1e-3.  Reference mask (`>= 1.0`): inside
    the reference's own 1e-6-perturbation noise floor (stored with the fixture) times a small factor."""
    c = golden("kitti_e2e")[size]
    im1, im2 = O.synthetic_pair(c["H"], c["W"], seed=c["seed"])
    for precision, bound in (("fp32", 1e-4), ("tf32x3", 1e-4), ("tf32", 1e-3)):
        net = _checkpoint_net(precision)
        eng = net._get_engine()
        res = {}
        for name, thr, key in (("robust", 0.9999, "flow_f_robust"), ("reference", 1.0, "flow_f_reference")):
            eng.mask = True if thr == 1.0 else thr
            with torch.no_grad():
                f, _, _ = eng.forward(im1.cuda(), im2.cuda())
            res[name] = O.epe(f.cpu(), c[key])
        eng.mask = True
        print("checkpoint %dx%d %-7s mean EPE vs reference: robust mask %.3g px, reference mask %.3g px (noise floor %.3g)"
              % (c["H"], c["W"], precision, res["robust"], res["reference"], c["noise_floor_px"]))
        assert res["robust"] <= bound, (precision, res)
        assert res["reference"] <= 3 * c["noise_floor_px"], (precision, res)


def test_checkpoint_through_the_public_api(golden):
    """net(input_dict) (CUDA-graph replay) with the checkpoint: the same flow as the eager engine, and the motion."""
    c = golden("kitti_e2e")[0]
    im1, im2 = O.synthetic_pair(c["H"], c["W"], seed=c["seed"])
    net = _checkpoint_net("tf32")
    with torch.no_grad():
        out = net({"im1": im1.cuda(), "im2": im2.cuda(), "if_loss": False})
        out2 = net({"im1": im1.cuda(), "im2": im2.cuda(), "if_loss": False})
    f = out["flow_f_out"].cpu()
    assert torch.equal(f, out2["flow_f_out"].cpu())
    assert O.epe(f, c["flow_f_reference"]) <= 3 * c["noise_floor_px"]
    assert abs(f[:, 0].mean().item() + 3) < 0.1 and abs(f[:, 1].mean().item() - 2) < 0.1


# ------------------------------------------------------------------ configurations other than test.py's
def test_config_modes_vs_reference_golden(golden):
    """UPFlow_net.config() class defaults (no normalisation, no SGU: model/upflow.py:311-323) and the other stable
    configurations run on the fused engine and match the reference's own outputs (oracle/make_golden_kitti.py)."""
    import contextlib
    import io
    import upflow_pytorch_b200 as pkg
    pkg.install_dropin()
    from model.upflow import UPFlow_net
    for c in golden("e2e_modes")["e2e"]:
        conf = UPFlow_net.config()
        with contextlib.redirect_stdout(io.StringIO()):
            conf.update(dict(c["params"]))
        net = conf()
        net.conv_precision = "fp32"
        sd = P.det_state_dict(c["wseed"])
        net.load_state_dict({k: v for k, v in sd.items() if k in net.state_dict()})
        net = net.cuda().eval()
        im1, im2 = O.synthetic_pair(*c["hw"], seed=c["pair_seed"])
        with torch.no_grad():
            out = net({"im1": im1.cuda(), "im2": im2.cuda(), "if_loss": False})
        e = O.epe(out["flow_f_out"].cpu(), c["flow_f_out"])
        agree = (out["occ_fw"].cpu() == c["occ_fw"]).float().mean().item()
        print("config", c["name"], "EPE vs reference %.3g px, occlusion agreement %.4f" % (e, agree))
        assert e <= 0.05 and agree >= 0.97, c["name"]


def test_pooled_moment_modes_at_the_operator(golden):
    """normalize_features' pooled modes (model/upflow.py:109-124) through statistics + upf_featnorm_combine + the fused
    correlation, against the reference's normalize_features followed by Corr_pyTorch.  The across-images mode divides
    by the std of the two VARIANCES (the reference's arithmetic), so values are large: relative tolerance."""
    from upflow_pytorch_b200 import ops
    for c in golden("e2e_modes")["ops"]:
        fa, fb = ops.to_pixel_major(c["fa"].cuda()), ops.to_pixel_major(c["fb"].cuda())
        N, H, W, C = fa.shape
        sa = torch.zeros(N, C, 2, dtype=torch.float64, device="cuda")
        sb, ca, cb = torch.zeros_like(sa), torch.zeros_like(sa), torch.zeros_like(sa)
        ops.k_stats(fa, sa)
        ops.k_stats(fb, sb)
        ops.k_norm_combine(sa, sb, ca, cb, N, C, H * W, 0, c["across_channels"], c["across_images"])
        out = torch.zeros(N, H, W, 81, device="cuda")
        ops.k_corr(fa, fb, out, 4, ca, cb, slope=1.0)
        got = out.permute(0, 3, 1, 2).cpu()
        rel = ((got - c["corr"]).abs().max() / c["corr"].abs().max()).item()
        na = torch.zeros_like(fa)
        ops.k_norm_apply(fa, ca, na)
        rel_n = ((na.permute(0, 3, 1, 2).cpu() - c["na"]).abs().max() / c["na"].abs().max()).item()
        print("pooled moments ch=%s img=%s: normalised rel err %.3g, correlation rel err %.3g" % (
            c["across_channels"], c["across_images"], rel_n, rel))
        assert rel_n <= 2e-5 and rel <= 5e-5


# ------------------------------------------------------------------ BASELINE configs 3 and 5 at their sizes
@pytest.mark.parametrize("case", [("sintel", 436, 1024, 8), ("hd", 1080, 1920, 2)])
def test_sintel_and_hd_sizes_vs_port(case):
    """436x1024 batch 8 and 1080x1920 batch 2 (BASELINE configs 3 and 5) against the CPU port, robust-mask diagnostic
    on both sides, shipped weights: fp32 and tf32x3 engines to rounding; tf32 (operands rounded to the nearest TF32)
    inside 1e-3 px at the Sintel size and inside 1e-2 px at 1080x1920 (measured 4.6e-3, tf32x3 1.5e-4: the checkpoint was trained at
    KITTI scale and its coarse levels amplify operand noise more at 17x30 than at 6x20 -- precision 'tf32x3' is the
    answer there).  Every image of the batch is its own pair (different seeds); the port runs them one by one."""
    import os
    GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    name, H, W, B = case
    sd = torch.load(os.path.join(GOLDEN, "upflow_kitti2015.pth"), weights_only=True)
    pairs = [O.synthetic_pair(H, W, seed=1234 + i) for i in range(B)]
    im1, im2 = torch.cat([p[0] for p in pairs]), torch.cat([p[1] for p in pairs])
    ref_idx = (0, B - 1) if B > 2 else tuple(range(B))       # the port costs 4-13 s per pair: first and last image
    P.MASK_THRESHOLD = 0.9999
    try:
        with torch.no_grad():
            refs = {i: P.forward_2_frame(im1[i:i + 1], im2[i:i + 1], sd)[0] for i in ref_idx}
    finally:
        P.MASK_THRESHOLD = 1.0
    for precision, bound in (("fp32", 1e-4), ("tf32x3", 1e-4 if name == "sintel" else 5e-4), ("tf32", 1e-3 if name == "sintel" else 1e-2)):
        eng = _engine(precision, sd, mask_threshold=0.9999)
        f, b, _ = eng.forward(im1.cuda(), im2.cuda())
        f = f.cpu()
        for i in ref_idx:
            e = O.epe(f[i:i + 1], refs[i])
            print(name, "%dx%d b%d" % (H, W, B), precision, "image", i, "mean EPE vs port %.3g px" % e)
            assert e <= bound, (name, precision, i, e)
        del eng
        torch.cuda.empty_cache()


def test_pipelined_inference_returns_every_flow_in_order(golden):
    """upflow_pytorch_b200.pipeline.PipelinedInference: pinned host pairs in, pinned host flows out one submit later,
    bit-identical to the plain call, for a sequence of DIFFERENT pairs (a stale input or result slot would show)."""
    from upflow_pytorch_b200.pipeline import PipelinedInference
    net = _checkpoint_net("tf32")
    pairs = [tuple(t.pin_memory() for t in O.synthetic_pair(96, 160, seed=50 + i)) for i in range(5)]
    with torch.no_grad():
        want = [net({"im1": a.cuda(), "im2": b.cuda(), "if_loss": False})["flow_f_out"].cpu() for a, b in pairs]
    pipe = PipelinedInference(net)
    got = []
    for a, b in pairs:
        r = pipe.submit(a, b)
        if r is not None:
            got.append(r.clone())
    got.append(pipe.flush().clone())
    assert pipe.flush() is None and len(got) == len(want)
    for g, w in zip(got, want):
        assert torch.equal(g, w)
    # two and three lanes: graphs of consecutive pairs replay concurrently on their own workspaces; same bits, same order
    for lanes in (2, 3):
        pipe = PipelinedInference(net, lanes=lanes)
        got = []
        for rep in range(2):
            for a, b in pairs:
                r = pipe.submit(a, b)
                assert (r is None) == (len(got) == 0 and pipe._k <= lanes)
                if r is not None:
                    got.append(r.clone())
        got += [r.clone() for r in pipe.drain()]
        assert pipe.drain() == [] and len(got) == 2 * len(want)
        for g, w in zip(got, want + want):
            assert torch.equal(g, w), lanes
    # the plain call still works after a pipeline put the engine back on lane 0
    with torch.no_grad():
        assert torch.equal(net({"im1": pairs[0][0].cuda(), "im2": pairs[0][1].cuda(), "if_loss": False})["flow_f_out"].cpu(), want[0])
