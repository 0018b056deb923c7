"""Pin the oracle by executing the UNMODIFIED reference (only where
/root/reference is mounted -- the authoring container; skipped on the GPU box)."""
import warnings

import pytest
import torch

from oracle import cpu_oracle as O
from oracle import ref_port as P
from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.have_reference(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    warnings.filterwarnings("ignore")
    mods = ref_shims.install()
    net = ref_shims.build_reference_net(checkpoint=True)
    return mods, net


def test_checkpoint_shapes_match_port_table(ref):
    _, net = ref
    sd = net.state_dict()
    shapes = P.reference_param_shapes()
    assert set(sd) == set(shapes)
    assert all(tuple(sd[k].shape) == shapes[k] for k in sd)
    assert sum(v.numel() for v in sd.values()) == 3494549


@pytest.mark.parametrize("hw", [(47, 156), (94, 311), (5, 7)])
def test_warp_mask_kitti_sizes(ref, hw):
    mods, _ = ref
    H, W = hw
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 4, H, W, generator=g)
    for fl in (torch.randn(1, 2, H, W, generator=g) * 3, torch.randint(-3, 4, (1, 2, H, W), generator=g).float()):
        r = mods.pwc_modules.WarpingLayer_no_div()(x, fl.clone())
        o = O.warp_mask(x, fl)
        assert torch.equal((r == 0).all(1), (o == 0).all(1))
        assert (r - o).abs().max().item() < 2e-6


def test_full_forward_port_is_bit_identical_with_checkpoint(ref):
    _, net = ref
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    im1, im2 = O.synthetic_pair(96, 160)
    with torch.no_grad():
        f, b, _ = net.forward_2_frame_v3(im1, im2)
        pf, pb, _ = P.forward_2_frame(im1, im2, sd)
    assert torch.equal(f, pf) and torch.equal(b, pb)
    # the checkpointed net recovers the synthetic (-3,+2) flow
    assert abs(f[:, 0].mean().item() + 3) < 0.3 and abs(f[:, 1].mean().item() - 2) < 0.3


def test_evaluation_metrics_match_the_reference():
    """upflow_pytorch_b200.evaluation against kitti_flow.Evaluation_bench.{flow_error_avg,outlier_pct} of the mounted
    reference (dataset/kitti_dataset.py:464-499)."""
    import importlib
    import torch
    from oracle import ref_shims
    if not ref_shims.have_reference():
        pytest.skip("reference not mounted")
    ref_shims.install()
    kd = importlib.import_module("dataset.kitti_dataset")
    from upflow_pytorch_b200 import evaluation as E
    g = torch.Generator().manual_seed(0)
    gt, pred = torch.randn(2, 2, 9, 11, generator=g) * 30, torch.randn(2, 2, 9, 11, generator=g) * 30
    pred = gt + (pred - gt) * 0.1
    mask = (torch.rand(2, 1, 9, 11, generator=g) > 0.4).float()
    bench = kd.kitti_flow.Evaluation_bench
    assert torch.equal(E.flow_error_avg(pred, gt, mask), bench.flow_error_avg(pred, gt, mask))
    assert torch.equal(E.outlier_pct(gt, pred, mask), bench.outlier_pct(gt, pred, mask))
