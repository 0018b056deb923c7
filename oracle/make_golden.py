"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt by EXECUTING the
unmodified reference (/root/reference, through oracle/ref_shims.py) on CPU.

Run here (the authoring container); the GPU box has no reference tree and only
reads the committed fixtures.   python -m oracle.make_golden

The reference ships no tests or golden vectors: these
fixtures ARE the pin.  Every fixture stores inputs, outputs and the reference
call that produced them; weights are ``ref_port.det_state_dict(seed)`` (keyed
by parameter name, reproducible anywhere) so only the seed is stored.
"""
import os
import warnings

import torch

from oracle import ref_shims, ref_port

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _g(seed):
    return torch.Generator().manual_seed(seed)


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(1)          # bit-stable reductions
    mods = ref_shims.install()
    os.makedirs(OUT, exist_ok=True)
    fx = {}

    # ---- a1/a2: correlation, Corr_pyTorch (utils/pytorch_correlation.py:27-50)
    corr = []
    for seed, shape, d in ((1234, (1, 32, 64, 64), 4),      # BASELINE config 1
                           (1, (2, 5, 7, 9), 4),             # ragged, C%4!=0, image smaller than the window
                           (2, (1, 16, 12, 39), 2),
                           (3, (1, 8, 13, 20), 6),
                           (4, (2, 196, 6, 20), 4)):         # KITTI L0 channels
        f1 = torch.randn(shape, generator=_g(seed))
        f2 = torch.randn(shape, generator=_g(seed + 100))
        out = mods.pytorch_correlation.Corr_pyTorch(d, 1, d, 1, 1)(f1, f2)
        keep_in = f1.numel() <= 40000
        corr.append(dict(seed=seed, shape=shape, d=d, f1=f1 if keep_in else None, f2=f2 if keep_in else None,
                         f1_sum=f1.double().sum().item(), out=out))
    fx["corr"] = corr

    # ---- a4: WarpingLayer_no_div (model/pwc_modules.py:184-207), tools.torch_warp
    warp = []
    wl = mods.pwc_modules.WarpingLayer_no_div()
    for seed, shape, kind in ((10, (2, 8, 13, 17), "randn3"), (11, (1, 4, 24, 78), "int"),
                              (12, (1, 3, 5, 7), "big"), (13, (2, 32, 12, 39), "small")):
        x = torch.randn(shape, generator=_g(seed))
        B, C, H, W = shape
        if kind == "randn3":
            fl = torch.randn(B, 2, H, W, generator=_g(seed + 1)) * 3
        elif kind == "int":
            fl = torch.randint(-3, 4, (B, 2, H, W), generator=_g(seed + 1)).float()
        elif kind == "big":
            fl = torch.randn(B, 2, H, W, generator=_g(seed + 1)) * 20
        else:
            fl = torch.randn(B, 2, H, W, generator=_g(seed + 1)) * 0.3
        warp.append(dict(x=x, flow=fl, kind=kind, out=wl(x, fl.clone()), out_nomask=mods.tools.torch_warp(x, fl.clone())))
    fx["warp"] = warp

    # ---- a5: normalize_features (model/upflow.py:94-137)
    norm = []
    for seed, shape in ((20, (2, 6, 9, 11)), (21, (1, 32, 24, 78))):
        f = torch.randn(shape, generator=_g(seed)) * 2.5 + 0.7
        out = mods.upflow.network_tools.normalize_features([f], normalize=True, center=True,
                                                           moments_across_channels=False,
                                                           moments_across_images=False)[0]
        norm.append(dict(f=f, out=out))
    fx["norm"] = norm

    # ---- a6: upsample2d_flow_as (model/pwc_modules.py:77-90)
    ups = []
    for seed, shape, hw, rate in ((30, (2, 2, 6, 20), (12, 39), True), (31, (1, 2, 5, 7), (20, 25), True),
                                  (32, (1, 1, 9, 11), (36, 41), False), (33, (1, 2, 4, 4), (4, 4), True)):
        x = torch.randn(shape, generator=_g(seed)) * 2
        tgt = torch.zeros(1, 1, *hw)
        ups.append(dict(x=x, hw=hw, if_rate=rate,
                        out=mods.pwc_modules.upsample2d_flow_as(x.clone(), tgt, mode="bilinear", if_rate=rate)))
    fx["upsample"] = ups

    # ---- modules with deterministic weights (seed only is stored)
    wseed = 7
    sd = ref_port.det_state_dict(wseed)
    net = ref_shims.build_reference_net(checkpoint=False)
    net.load_state_dict(sd)
    with torch.no_grad():
        # a8 FlowEstimatorDense_v2, a9 ContextNetwork_v2_ (pwc_modules.py:279-286, :401-412)
        x = torch.randn(1, 115, 10, 14, generator=_g(40))
        x5, res = net.flow_estimators(x)
        ctx_in = torch.randn(1, 565, 9, 12, generator=_g(41))
        fx["estimator"] = dict(wseed=wseed, x=x, x5=x5, out=res)
        fx["context"] = dict(wseed=wseed, x=ctx_in, out=net.context_networks(ctx_in))
        # a7 sgu_model.forward (model/upflow.py:71-89), same-res and output-level variants
        flow = torch.randn(1, 2, 12, 20, generator=_g(50)) * 2
        f1 = torch.randn(1, 32, 12, 20, generator=_g(51))
        f2 = torch.randn(1, 32, 12, 20, generator=_g(52))
        _, up, iflow, imask = net.sgi_model(flow.clone(), f1, f2)
        big = mods.pwc_modules.upsample2d_flow_as(flow.clone(), torch.zeros(1, 1, 47, 78), mode="bilinear", if_rate=True)
        _, up2, iflow2, imask2 = net.sgi_model(flow.clone(), f1, f2, output_level_flow=big.clone())
        fx["sgu"] = dict(wseed=wseed, flow=flow, f1=f1, f2=f2, flow_up=up, inter_flow=iflow, inter_mask=imask,
                         output_level_flow=big, flow_up_out=up2, inter_flow_out=iflow2, inter_mask_out=imask2)
        # a10 decode_level_res (model/upflow.py:535-573), level 2 (C=96)
        x1 = torch.randn(1, 96, 12, 16, generator=_g(60)).relu_()
        x2 = torch.randn(1, 96, 12, 16, generator=_g(61)).relu_()
        a1 = torch.randn(1, 32, 12, 16, generator=_g(62))
        a2 = torch.randn(1, 32, 12, 16, generator=_g(63))
        ff = torch.randn(1, 2, 6, 8, generator=_g(64))
        fb = torch.randn(1, 2, 6, 8, generator=_g(65))
        o = net.decode_level_res(level=2, flow_1=ff.clone(), flow_2=fb.clone(), feature_1=x1, feature_1_1x1=a1,
                                 feature_2=x2, feature_2_1x1=a2, img_ori_1=None, img_ori_2=None)
        fx["decode_level"] = dict(wseed=wseed, level=2, x1=x1, x2=x2, a1=a1, a2=a2, flow_1=ff, flow_2=fb,
                                  flow_1_up=o[0], flow_2_up=o[1], res_1=o[2], res_2=o[3])
        # end to end: UPFlow_net.forward (model/upflow.py:370-392) on the synthetic (-3,+2) pair
        from oracle.cpu_oracle import synthetic_pair
        im1, im2 = synthetic_pair(64, 96, seed=1234)
        out = net({"im1": im1, "im2": im2, "if_loss": False})
        _, _, flows = net.forward_2_frame_v3(im1, im2)
        fx["e2e"] = dict(wseed=wseed, hw=(64, 96), pair_seed=1234, flow_f_out=out["flow_f_out"],
                         flow_b_out=out["flow_b_out"], occ_fw=out["occ_fw"], occ_bw=out["occ_bw"],
                         flows=[[a.clone(), b.clone()] for a, b in flows])
    for k, v in fx.items():
        torch.save(v, os.path.join(OUT, k + ".pt"))
        print(k, os.path.getsize(os.path.join(OUT, k + ".pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
