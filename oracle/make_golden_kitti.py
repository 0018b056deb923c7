"""TEST INFRASTRUCTURE ONLY -- the KITTI-size end-to-end pin with the SHIPPED weights.

    python -m oracle.make_golden_kitti        (here, where /root/reference is mounted)

Writes
  tests/golden/upflow_kitti2015.pth   the state dict of /root/reference/scripts/upflow_kitti2015.pth (80 tensors,
                                      3,494,549 fp32 parameters, moved to the CPU; same file format, so
                                      `net.load_model(path, if_relax=True)` loads it like test.py:31-38 does).  Weights
                                      are data, not source; BASELINE config 2 names this checkpoint and it cannot
                                      travel to the GPU box any other way.
  tests/golden/kitti_e2e.pt           flows of the synthetic (-3,+2) pair at 375x1242 (cpu_oracle.synthetic_pair, seed
                                      1234; also 436x1024, seed 1234):
        flow_f_reference  UPFlow_net.forward_2_frame_v3 of the UNMODIFIED reference on CPU (`mask >= 1.0`)
        flow_f_robust     the op-for-op port (bit-identical to the reference, tests/test_oracle_vs_reference.py) with
                          the validity threshold relaxed to 0.9999 -- the diagnostic both sides use to compare
                          arithmetic without the reference's 1-ulp mask flips
        noise_floor_px    mean EPE between the reference and itself when the inputs are scaled by 1 + 1e-6
  Only forward flows are stored (fp32, 3.7 MB per 375x1242 field).
  tests/golden/e2e_modes.pt           UPFlow_net.forward (flows + occlusion masks) of the UNMODIFIED reference on the
                                      64x96 pair with det_state_dict(7) under the configurations the engine serves
                                      besides test.py's: the CLASS DEFAULTS of UPFlow_net.config (model/upflow.py:311-323:
                                      no normalisation, no SGU), and every pooled moment mode of normalize_features.
"""
import os
import warnings

import torch

from oracle import cpu_oracle as O
from oracle import ref_port as P
from oracle import ref_shims

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SIZES = ((375, 1242), (436, 1024))


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    net = ref_shims.build_reference_net(checkpoint=True)
    sd = {k: v.detach().float().cpu().clone() for k, v in net.state_dict().items()}
    assert len(sd) == 80 and sum(v.numel() for v in sd.values()) == 3494549
    torch.save(sd, os.path.join(OUT, "upflow_kitti2015.pth"))
    cases = []
    for H, W in SIZES:
        im1, im2 = O.synthetic_pair(H, W, seed=1234)
        with torch.no_grad():
            ref_f = net.forward_2_frame_v3(im1, im2)[0]
            port_f = P.forward_2_frame(im1, im2, sd)[0]
            assert torch.equal(ref_f, port_f), "the port is no longer bit-identical to the reference"
            pert_f = net.forward_2_frame_v3(im1 * (1 + 1e-6), im2 * (1 + 1e-6))[0]
            P.MASK_THRESHOLD = 0.9999
            try:
                rob_f = P.forward_2_frame(im1, im2, sd)[0]
                rob_pert = P.forward_2_frame(im1 * (1 + 1e-6), im2 * (1 + 1e-6), sd)[0]
            finally:
                P.MASK_THRESHOLD = 1.0
        c = dict(H=H, W=W, seed=1234, flow_f_reference=ref_f.clone(), flow_f_robust=rob_f.clone(),
                 noise_floor_px=O.epe(ref_f, pert_f), noise_floor_robust_px=O.epe(rob_f, rob_pert),
                 mean_flow=ref_f.mean(dim=(0, 2, 3)).tolist(), torch=torch.__version__)
        print("%dx%d: mean flow %s, reference noise floor %.4f px (robust mask: %.2e), robust vs reference %.4f px"
              % (H, W, c["mean_flow"], c["noise_floor_px"], c["noise_floor_robust_px"], O.epe(ref_f, rob_f)))
        cases.append(c)
    torch.save(cases, os.path.join(OUT, "kitti_e2e.pt"))

    # ---- configurations other than test.py's, small size, deterministic weights
    torch.set_num_threads(1)
    modes = []
    im1, im2 = O.synthetic_pair(64, 96, seed=1234)
    for name, params in (("class defaults", {"if_norm_before_cost_volume": False, "norm_moments_across_channels": True,
                                             "norm_moments_across_images": True, "if_sgu_upsample": False}),
                         ("no norm, sgu", {"if_norm_before_cost_volume": False}),
                         ("norm across channels", {"norm_moments_across_channels": True})):
        rnet = ref_shims.build_reference_net(params=params, checkpoint=False)
        rnet.load_state_dict(P.det_state_dict(7), strict=False)
        with torch.no_grad():
            out = rnet({"im1": im1, "im2": im2, "if_loss": False})
        cfg = dict(ref_shims.TEST_PARAMS, **params)
        modes.append(dict(name=name, params={k: cfg[k] for k in ("if_norm_before_cost_volume", "norm_moments_across_channels",
                                                                 "norm_moments_across_images", "if_sgu_upsample")},
                          wseed=7, hw=(64, 96), pair_seed=1234, flow_f_out=out["flow_f_out"].clone(),
                          flow_b_out=out["flow_b_out"].clone(), occ_fw=out["occ_fw"].clone(), occ_bw=out["occ_bw"].clone()))
        print(name, "mean flow", out["flow_f_out"].mean(dim=(0, 2, 3)).tolist())
    # moments_across_images makes the variance the VARIANCE OF THE TWO VARIANCES (model/upflow.py:121-124): the features
    # blow up by orders of magnitude and an end-to-end comparison measures chaos, so those modes are pinned at the
    # operator: normalize_features on a pair, then the correlation (Corr_pyTorch)
    mods = ref_shims.install()
    ops_cases = []
    g = torch.Generator().manual_seed(99)
    for across_ch, across_img in ((True, True), (True, False), (False, True)):
        fa = torch.randn(2, 32, 12, 20, generator=g) * 1.5 + 0.3
        fb = torch.randn(2, 32, 12, 20, generator=g) * 0.7 - 0.2
        na, nb = mods.upflow.network_tools.normalize_features((fa, fb), normalize=True, center=True,
                                                              moments_across_channels=across_ch, moments_across_images=across_img)
        corr = mods.pytorch_correlation.Corr_pyTorch(4, 1, 4, 1, 1)(na, nb)
        ops_cases.append(dict(across_channels=across_ch, across_images=across_img, fa=fa, fb=fb, na=na, nb=nb, corr=corr))
    torch.save(dict(e2e=modes, ops=ops_cases), os.path.join(OUT, "e2e_modes.pt"))


if __name__ == "__main__":
    main()
