"""upflow_pytorch_b200 -- B200-native (sm_100a) implementation of UPFlow's
per-pyramid-level decoder hot path behind the reference's Python API.

    import upflow_pytorch_b200 as upf
    upf.install_dropin()                  # puts `model` / `utils` ahead of the reference's on sys.path
    from model.upflow import UPFlow_net   # same names as coolbeam/UPFlow_pytorch

Layers: csrc/ (CUDA kernels + C ABI, include/upflow_b200.h) -> _ext.py (ctypes)
-> ops.py (tensor wrappers, autograd) -> engine.py (fused two-frame decoder)
-> dropin/{model,utils} (the reference's module names and signatures).
"""
import os
import sys

__version__ = "0.1.0"

DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


def install_dropin():
    """Make `import model.upflow`, `model.pwc_modules`, `model.correlation_package.correlation`,
    `utils.pytorch_correlation`, `utils.tools` resolve to this package's drop-in modules."""
    for name in list(sys.modules):
        if name.split(".")[0] in ("model", "utils"):
            mod = sys.modules[name]
            f = getattr(mod, "__file__", "") or ""
            if not f.startswith(DROPIN_DIR):
                del sys.modules[name]
    if DROPIN_DIR in sys.path:
        sys.path.remove(DROPIN_DIR)
    sys.path.insert(0, DROPIN_DIR)
    return DROPIN_DIR


def build_model(params=None, state_dict=None, device="cuda", conv_precision="tf32"):
    """UPFlow_net in test.py's configuration (test.py:22-38), eval mode, on `device`."""
    install_dropin()
    from model.upflow import UPFlow_net
    import contextlib
    import io
    conf = UPFlow_net.config()
    cfg = {"if_norm_before_cost_volume": True, "norm_moments_across_channels": False,
           "norm_moments_across_images": False, "if_froze_pwc": False, "if_use_cor_pytorch": False,
           "if_sgu_upsample": True}
    cfg.update(params or {})
    with contextlib.redirect_stdout(io.StringIO()):
        conf.update(cfg)
    net = conf()
    net.conv_precision = conv_precision
    if state_dict is not None:
        net.load_state_dict(state_dict)
    return net.to(device).eval()
