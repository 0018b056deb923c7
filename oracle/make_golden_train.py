"""TEST INFRASTRUCTURE ONLY -- tests/golden/train_step.pt: the reference's own training forward + backward
(UPFlow_net.forward with if_loss=True, model/upflow.py:370-491; loss.backward(), scripts/simple_train.py:140-146)
EXECUTED on CPU through oracle/ref_shims.py (training shim).  Stores the loss terms and, per parameter, the gradient
L2 norm plus the full gradient of a few small tensors.     python -m oracle.make_golden_train
"""
import os
import warnings

import torch

from oracle import ref_shims, ref_port
from oracle.cpu_oracle import synthetic_pair

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
KEEP = ("conv_1x1.4.0.weight", "context_networks.convs.6.0.weight", "flow_estimators.conv_last.0.bias",
        "sgi_model.dense_estimator_mask.conv_last.0.weight", "feature_pyramid_extractor.convs.0.0.0.bias")
CONF = {"if_norm_before_cost_volume": True, "norm_moments_across_channels": False, "norm_moments_across_images": False,
        "if_froze_pwc": False, "if_use_cor_pytorch": True, "if_sgu_upsample": True, "if_use_boundary_warp": False,
        "smooth_order_1_weight": 1, "smooth_order_2_weight": 0.5, "photo_loss_type": "abs_robust", "photo_loss_delta": 0.4,
        "photo_loss_use_occ": True, "photo_loss_census_weight": 0, "multi_scale_distillation_weight": 0.01,
        "multi_scale_distillation_style": "upup", "multi_scale_distillation_occ": True}


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(1)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
    ref_shims.install(patch_training=True)
    wseed, hw, pair_seed = 9, (64, 96), 4321
    net = ref_shims.build_reference_net(params=CONF, checkpoint=False)
    net.load_state_dict(ref_port.det_state_dict(wseed))
    net.train()
    im1, im2 = synthetic_pair(*hw, seed=pair_seed, batch=2)
    out = net({"im1": im1, "im2": im2, "if_loss": True})
    loss = out["photo_loss"].mean() + out["smooth_loss"].mean() + out["msd_loss"].mean()
    loss.backward()
    fx = dict(wseed=wseed, hw=hw, pair_seed=pair_seed, batch=2, conf=CONF,
              photo_loss=out["photo_loss"].item(), smooth_loss=out["smooth_loss"].item(), msd_loss=out["msd_loss"].item(),
              loss=loss.item(), grad_norm={n: p.grad.norm().item() for n, p in net.named_parameters()},
              grads={n: p.grad.clone() for n, p in net.named_parameters() if n in KEEP})
    torch.save(fx, os.path.join(OUT, "train_step.pt"))
    print({k: fx[k] for k in ("photo_loss", "smooth_loss", "msd_loss", "loss")}, len(fx["grad_norm"]), "gradients")

    # ---- scripts/simple_train.py's DEFAULT loss configuration: boundary-dilated warp on the un-cropped frames
    # (utils/tools.py:350-499) + the census term (utils/loss.py:51-91)
    conf2 = dict(CONF, if_use_boundary_warp=True, photo_loss_census_weight=1.0, if_sgu_upsample=False, photo_loss_use_occ=False)
    net = ref_shims.build_reference_net(params=conf2, checkpoint=False)
    net.load_state_dict(ref_port.det_state_dict(wseed), strict=False)
    net.train()
    raw1, raw2 = synthetic_pair(hw[0] + 16, hw[1] + 16, seed=pair_seed + 1, batch=2)
    start = torch.tensor([[8.0, 6.0], [3.0, 9.0]]).view(2, 2, 1, 1)          # (x, y) of the crop inside the raw frame
    crop = lambda t: torch.stack([t[b, :, int(start[b, 1]):int(start[b, 1]) + hw[0], int(start[b, 0]):int(start[b, 0]) + hw[1]]
                                  for b in range(2)])
    out = net({"im1": crop(raw1), "im2": crop(raw2), "im1_raw": raw1, "im2_raw": raw2, "start": start, "if_loss": True})
    loss = out["photo_loss"].mean() + out["smooth_loss"].mean() + out["census_loss"].mean() + out["msd_loss"].mean()
    loss.backward()
    fx2 = dict(wseed=wseed, hw=hw, pair_seed=pair_seed + 1, batch=2, conf=conf2, start=start,
               photo_loss=out["photo_loss"].item(), smooth_loss=out["smooth_loss"].item(), msd_loss=out["msd_loss"].item(),
               census_loss=out["census_loss"].item(), loss=loss.item(),
               grad_norm={n: p.grad.norm().item() for n, p in net.named_parameters() if p.grad is not None})
    torch.save(fx2, os.path.join(OUT, "train_step_boundary_census.pt"))
    print({k: fx2[k] for k in ("photo_loss", "smooth_loss", "census_loss", "msd_loss", "loss")}, len(fx2["grad_norm"]), "gradients")

    # ---- the two loss-side ops alone (CPU-checkable: the drop-in versions are plain torch)
    import importlib
    rtools = importlib.import_module("utils.tools").tools
    rloss = importlib.import_module("utils.loss").loss_functions
    g = torch.Generator().manual_seed(77)
    I = torch.rand(2, 3, 20, 24, generator=g)
    flow = torch.randn(2, 2, 12, 16, generator=g) * 3
    st = torch.tensor([[3.0, 2.0], [5.0, 4.0]]).view(2, 2, 1, 1)
    a, b = torch.rand(2, 3, 12, 16, generator=g), torch.rand(2, 3, 12, 16, generator=g)
    mask = (torch.rand(2, 1, 12, 16, generator=g) > 0.3).float()
    ops_fx = dict(I=I, flow=flow, start=st, warp=rtools.boundary_dilated_warp.warp_im(I, flow, st), a=a, b=b, mask=mask,
                  census_occ=rloss.census_loss_torch(a, b, mask, 0.4, False, True, True).item(),
                  census_noocc=rloss.census_loss_torch(a, b, mask, 0.4, False, False, True).item(),
                  census_charb=rloss.census_loss_torch(a, b, mask, 0.4, True, True, True).item())
    torch.save(ops_fx, os.path.join(OUT, "loss_ops.pt"))
    print("loss ops", ops_fx["census_occ"], ops_fx["census_noocc"], ops_fx["census_charb"])


if __name__ == "__main__":
    main()
