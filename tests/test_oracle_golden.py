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
