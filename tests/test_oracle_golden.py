"""The CPU oracle (oracle/cpu_oracle.py: first-principles restatement;
oracle/ref_port.py: op-for-op port) against the golden vectors produced by
executing the unmodified reference (oracle/make_golden.py).  Runs anywhere."""
import pytest
import torch

from oracle import cpu_oracle as O
from oracle import ref_port as P


@pytest.fixture(autouse=True)
def _one_thread():
    # the fixtures were generated single-threaded; oneDNN's conv reduction
    # order (hence the last bit) depends on the thread count
    n = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(n)


def _regen(seed, shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed))


def _corr_inputs(c):
    f1 = c["f1"] if c["f1"] is not None else _regen(c["seed"], c["shape"])
    f2 = c["f2"] if c["f2"] is not None else _regen(c["seed"] + 100, c["shape"])
    assert abs(f1.double().sum().item() - c["f1_sum"]) < 1e-9, "seeded input drifted"
    return f1, f2


def test_correlation(golden):
    for c in golden("corr"):
        f1, f2 = _corr_inputs(c)
        out = O.correlation(f1, f2, c["d"])
        assert out.shape == c["out"].shape
        # sum of C fp32 products, different association than torch.mean: 1e-5 abs
        assert (out - c["out"]).abs().max().item() <= 1e-5
        assert torch.equal(P.corr_unfold(f1, f2, c["d"]), c["out"])


def test_correlation_backward_matches_autograd(golden):
    c = golden("corr")[1]
    f1, f2 = _corr_inputs(c)
    f1 = f1.double().requires_grad_()
    f2 = f2.double().requires_grad_()
    go = torch.randn(c["out"].shape, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    P.corr_unfold(f1, f2, c["d"]).backward(go)
    g1, g2 = O.correlation_backward(f1.detach(), f2.detach(), go, c["d"])
    assert (g1 - f1.grad).abs().max().item() < 1e-12
    assert (g2 - f2.grad).abs().max().item() < 1e-12


def test_warp_mask_bit_exact_mask(golden):
    for w in golden("warp"):
        out = O.warp_mask(w["x"], w["flow"])
        ref = w["out"]
        # the validity mask must agree pixel for pixel (mask >= 1.0, pwc_modules.py:206)
        zr = (ref == 0).all(1)
        zo = (out == 0).all(1)
        assert torch.equal(zr, zo), w["kind"]
        assert (out - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())
        assert (O.torch_warp(w["x"], w["flow"]) - w["out_nomask"]).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())
        assert torch.equal(P.warp_mask(w["x"], w["flow"]), ref)


def test_normalize(golden):
    for n in golden("norm"):
        assert (O.normalize_features(n["f"]) - n["out"]).abs().max().item() <= 2e-6
        assert torch.equal(P.normalize(n["f"]), n["out"])


def test_upsample(golden):
    for u in golden("upsample"):
        h, w = u["hw"]
        if u["if_rate"]:
            out = O.upsample2d_flow_as(u["x"], h, w)
        else:
            out = O.resize_bilinear_ac(u["x"], h, w)
        assert (out - u["out"]).abs().max().item() <= 4e-6
        assert torch.equal(P.upsample2d_flow_as(u["x"], h, w, u["if_rate"]), u["out"])


def test_estimator_and_context(golden):
    e = golden("estimator")
    sd = P.det_state_dict(e["wseed"])
    x5, out = O.dense_block(e["x"], sd, "flow_estimators")
    assert (x5 - e["x5"]).abs().max().item() <= 2e-5
    assert (out - e["out"]).abs().max().item() <= 2e-5
    c = golden("context")
    assert (O.context_network(c["x"], sd) - c["out"]).abs().max().item() <= 2e-5
    assert torch.equal(P.dense(e["x"], sd, "flow_estimators")[1], e["out"])
    assert torch.equal(P.context(c["x"], sd), c["out"])


def test_sgu(golden):
    s = golden("sgu")
    sd = P.det_state_dict(s["wseed"])
    up = O.sgu_forward(s["flow"], s["f1"], s["f2"], sd)
    assert (up - s["flow_up"]).abs().max().item() <= 5e-5
    up2 = O.sgu_forward(s["flow"], s["f1"], s["f2"], sd, output_level_flow=s["output_level_flow"])
    assert (up2 - s["flow_up_out"]).abs().max().item() <= 5e-5
    # the blend alone, teacher-forced with the reference's inter_flow / mask
    b = O.sgu_blend(s["output_level_flow"], s["inter_flow_out"], s["inter_mask_out"])
    assert (b - s["flow_up_out"]).abs().max().item() <= 5e-6
    assert torch.equal(P.sgu(s["flow"], s["f1"], s["f2"], sd)[0], s["flow_up"])


def test_decode_level(golden):
    g = golden("decode_level")
    sd = P.det_state_dict(g["wseed"])
    o = P.decode_level(g["level"], g["flow_1"], g["flow_2"], g["x1"], g["a1"], g["x2"], g["a2"], sd)
    for got, key in zip(o, ("flow_1_up", "flow_2_up", "res_1", "res_2")):
        assert torch.equal(got, g[key]), key


def test_end_to_end_port_bit_identical(golden):
    g = golden("e2e")
    sd = P.det_state_dict(g["wseed"])
    im1, im2 = O.synthetic_pair(*g["hw"], seed=g["pair_seed"])
    with torch.no_grad():
        f, b, flows = P.forward_2_frame(im1, im2, sd)
    assert torch.equal(f, g["flow_f_out"])
    assert torch.equal(b, g["flow_b_out"])
    for (a, c), (ga, gc) in zip(flows, g["flows"]):
        assert torch.equal(a, ga) and torch.equal(c, gc)


def test_end_to_end_restatement_within_noise_floor(golden):
    """cpu_oracle re-associates sums, so pixels sitting on the mask>=1.0
    discontinuity may flip (SURVEY 7 'hard parts'): coarse levels agree to
    rounding, the full-resolution field within the measured noise floor."""
    g = golden("e2e")
    sd = P.det_state_dict(g["wseed"])
    im1, im2 = O.synthetic_pair(*g["hw"], seed=g["pair_seed"])
    f, b, flows = O.forward_2_frame(im1, im2, sd)
    assert O.epe(flows[-1][0], g["flows"][-1][0]) < 1e-5      # coarsest level: no warp yet
    assert O.epe(f, g["flow_f_out"]) < 0.1


def test_loss_side_ops_match_the_reference(golden):
    """tools.boundary_dilated_warp.warp_im and loss_functions.census_loss_torch of the drop-in (plain torch, run on CPU
    here) against the reference's own outputs (tests/golden/loss_ops.pt, oracle/make_golden_train.py)."""
    import upflow_pytorch_b200
    upflow_pytorch_b200.install_dropin()
    from utils.tools import tools
    from utils.loss import loss_functions
    g = golden("loss_ops")
    out = tools.boundary_dilated_warp.warp_im(g["I"], g["flow"], g["start"])
    assert out.shape == g["warp"].shape
    assert (out - g["warp"]).abs().max().item() <= 1e-6
    for key, charb, occ in (("census_occ", False, True), ("census_noocc", False, False), ("census_charb", True, True)):
        v = loss_functions.census_loss_torch(g["a"], g["b"], g["mask"], 0.4, charb, occ, True).item()
        assert abs(v - g[key]) <= 1e-5 * abs(g[key]), (key, v, g[key])


def test_port_config_modes_match_the_reference(golden):
    """the port under the configurations other than test.py's (model/upflow.py:311-323 class defaults: no
    normalisation, no SGU; pooled moments) against the reference's own outputs, and normalize_features' pooled modes
    at the operator (oracle/make_golden_kitti.py)."""
    g = golden("e2e_modes")
    for c in g["e2e"]:
        sd = P.det_state_dict(c["wseed"])
        im1, im2 = O.synthetic_pair(*c["hw"], seed=c["pair_seed"])
        p = c["params"]
        P.NORM_MODE = (p["if_norm_before_cost_volume"], p["norm_moments_across_channels"], p["norm_moments_across_images"])
        try:
            with torch.no_grad():
                f, b, _ = P.forward_2_frame(im1, im2, sd, use_sgu=p["if_sgu_upsample"])
        finally:
            P.NORM_MODE = (True, False, False)
        assert torch.equal(f, c["flow_f_out"]) and torch.equal(b, c["flow_b_out"]), c["name"]
    for c in g["ops"]:
        P.NORM_MODE = (True, c["across_channels"], c["across_images"])
        try:
            na, nb = P.normalize_pair(c["fa"], c["fb"])
        finally:
            P.NORM_MODE = (True, False, False)
        assert torch.equal(na, c["na"]) and torch.equal(nb, c["nb"])
        assert torch.equal(P.corr_unfold(na, nb, 4), c["corr"])


def test_kitti_size_golden_is_self_consistent(golden):
    """the 375x1242 pin with the shipped weights: fixture integrity (the full-size forward itself is re-run against the
    live reference by oracle/make_golden_kitti.py, which asserts port == reference bit for bit before it writes)."""
    import os
    GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sd = torch.load(os.path.join(GOLDEN, "upflow_kitti2015.pth"), weights_only=True)
    shapes = P.reference_param_shapes()
    assert set(sd) == set(shapes) and all(tuple(sd[k].shape) == shapes[k] for k in sd)
    assert sum(v.numel() for v in sd.values()) == 3494549
    for c in golden("kitti_e2e"):
        assert c["flow_f_reference"].shape == (1, 2, c["H"], c["W"]) == c["flow_f_robust"].shape
        # the checkpointed network recovers the synthetic (-3, +2) motion
        assert abs(c["mean_flow"][0] + 3) < 0.1 and abs(c["mean_flow"][1] - 2) < 0.1
        # what separates the two masks is the reference's own noise floor, not arithmetic
        assert c["noise_floor_robust_px"] < 1e-5 < c["noise_floor_px"] < 0.1
        assert O.epe(c["flow_f_reference"], c["flow_f_robust"]) < 0.1
    # a reduced-size forward of the port with these weights reproduces the motion too (seconds on one thread)
    im1, im2 = O.synthetic_pair(96, 160)
    with torch.no_grad():
        f = P.forward_2_frame(im1, im2, sd)[0]
    assert abs(f[:, 0].mean().item() + 3) < 0.3 and abs(f[:, 1].mean().item() - 2) < 0.3


def test_training_port_matches_the_reference_step(golden):
    """oracle/ref_port_train.py (the anchor bench.py times as the reference's GPU training path) against the loss terms
    and gradients of the reference's OWN training step (tests/golden/train_step.pt, oracle/make_golden_train.py)."""
    from oracle import ref_port_train as PT
    g = golden("train_step")
    sd = {k: v.clone().requires_grad_() for k, v in P.det_state_dict(g["wseed"]).items()}
    im1, im2 = O.synthetic_pair(*g["hw"], seed=g["pair_seed"], batch=g["batch"])
    c = g["conf"]
    out = PT.training_loss(im1, im2, sd, smooth1_weight=c["smooth_order_1_weight"], smooth2_weight=c["smooth_order_2_weight"],
                           photo_use_occ=c["photo_loss_use_occ"], msd_weight=c["multi_scale_distillation_weight"],
                           msd_occ=c["multi_scale_distillation_occ"])
    for k in ("photo_loss", "smooth_loss", "msd_loss", "loss"):
        assert abs(out[k].item() - g[k]) <= 1e-6 * max(1.0, abs(g[k])), (k, out[k].item(), g[k])
    out["loss"].backward()
    for n, ref in g["grads"].items():
        assert torch.allclose(sd[n].grad, ref, rtol=1e-4, atol=1e-7), n
    for n, ref in g["grad_norm"].items():
        assert abs(sd[n].grad.norm().item() - ref) <= 1e-4 * max(ref, 1e-6), n
