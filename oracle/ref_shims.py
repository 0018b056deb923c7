"""TEST INFRASTRUCTURE ONLY -- import shims so the UNMODIFIED reference at
/root/reference can be imported on the installed torch (2.11) in THIS container.

Only tests/, oracle/make_golden.py and bench.py's cpu_baseline leg may use this
module; the product package (upflow_pytorch_b200/) never imports it.  The GPU
box has no /root/reference: everything here degrades to ``have_reference() ==
False`` there and the parity tests fall back to the committed golden vectors
and the restatement in ``oracle/cpu_oracle.py``.

Shims:
 1. ``torch.utils.data.dataloader._DataLoaderIter`` (utils/tools.py:2) is gone
    -> alias to ``_BaseDataLoaderIter``.
 2. ``imageio``, ``png`` (utils/tools.py:8,25) and ``tensorflow``
    (dataset/kitti_dataset.py:10) are not installed -> stub modules carrying a
    ``__spec__``.
 3. ``correlation_cuda`` (model/correlation_package/correlation.py:4) is not
    built -> stub; the reference runs with ``if_use_cor_pytorch=True``.
 4. training only: ``upsample2d_flow_as`` multiplies ``chunk`` views in place
    (model/pwc_modules.py:86-88) which autograd rejects since torch 1.5; the
    same arithmetic is rebound out of place (opt-in, ``patch_training=True``).
 5. the checkpoint holds CUDA tensors and ``load_model`` calls ``torch.load``
    without ``map_location`` (utils/tools.py:117) -> wrapped while loading.
"""
import contextlib
import importlib
import importlib.machinery
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("UPFLOW_REFERENCE_ROOT", "/root/reference")
CHECKPOINT = os.path.join(REFERENCE_ROOT, "scripts", "upflow_kitti2015.pth")

# the inference configuration the reference ships (test.py:22-30), with the
# pure-PyTorch correlation selected because correlation_cuda cannot load.
TEST_PARAMS = {
    "if_norm_before_cost_volume": True,
    "norm_moments_across_channels": False,
    "norm_moments_across_images": False,
    "if_froze_pwc": False,
    "if_use_cor_pytorch": True,
    "if_sgu_upsample": True,
}


def have_reference():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "upflow.py"))


def _stub(name):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    sys.modules[name] = m
    return m


_installed = {}


def install(patch_training=False):
    """Put the reference on sys.path behind the shims; returns its modules.

    Idempotent.  The reference's top-level packages are called ``model`` and
    ``utils``; the product's drop-in mirror uses the same names, so a process
    must pick one of the two (tests that need both run the reference in a
    subprocess or purge ``sys.modules`` through ``purge()``)."""
    if not have_reference():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if _installed.get("mods") is not None and sys.modules.get("model.upflow") is not _installed["mods"].upflow:
        _installed.clear()          # somebody (the drop-in) replaced `model` since: import afresh
    if _installed.get("mods") is None:
        import torch.utils.data.dataloader as dl
        if not hasattr(dl, "_DataLoaderIter"):
            dl._DataLoaderIter = dl._BaseDataLoaderIter
        for name in ("imageio", "png", "tensorflow", "correlation_cuda"):
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
        # the product's drop-in mirror uses the same top-level names: evict it
        for k in list(sys.modules):
            if k.split(".")[0] in ("model", "utils"):
                f = getattr(sys.modules[k], "__file__", "") or ""
                if not f.startswith(REFERENCE_ROOT):
                    del sys.modules[k]
        sys.path[:] = [p for p in sys.path if not p.rstrip("/").endswith("upflow_pytorch_b200/dropin")]
        if REFERENCE_ROOT in sys.path:
            sys.path.remove(REFERENCE_ROOT)
        sys.path.insert(0, REFERENCE_ROOT)
        mods = types.SimpleNamespace()
        mods.upflow = importlib.import_module("model.upflow")
        mods.pwc_modules = importlib.import_module("model.pwc_modules")
        mods.pytorch_correlation = importlib.import_module("utils.pytorch_correlation")
        mods.tools = importlib.import_module("utils.tools").tools
        assert mods.upflow.__file__.startswith(REFERENCE_ROOT), mods.upflow.__file__
        _installed["mods"] = mods
    mods = _installed["mods"]
    if patch_training and not _installed.get("train"):
        import torch.nn.functional as F

        def upsample2d_flow_as(inputs, target_as, mode="bilinear", if_rate=False):
            # same arithmetic as model/pwc_modules.py:77-90, out of place
            _, _, h, w = target_as.size()
            res = F.interpolate(inputs, [h, w], mode=mode, align_corners=True)
            if if_rate:
                _, _, h_, w_ = inputs.size()
                u, v = res.chunk(2, dim=1)
                res = torch.cat([u * (w / w_), v * (h / h_)], dim=1)
            return res

        mods.upflow.upsample2d_flow_as = upsample2d_flow_as
        mods.pwc_modules.upsample2d_flow_as = upsample2d_flow_as
        _installed["train"] = True
    return mods


def purge():
    """Forget the reference's ``model``/``utils``/``dataset`` modules."""
    for k in list(sys.modules):
        if k.split(".")[0] in ("model", "utils", "dataset", "test"):
            del sys.modules[k]
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
    _installed.clear()


@contextlib.contextmanager
def cpu_checkpoint_load():
    real = torch.load

    def load(path, *a, **kw):
        kw.setdefault("map_location", "cpu")
        return real(path, *a, **kw)

    torch.load = load
    try:
        yield
    finally:
        torch.load = real


def build_reference_net(params=None, checkpoint=True, quiet=True):
    """The reference UPFlow_net, eval mode, on CPU, test.py's configuration."""
    mods = install()
    conf = mods.upflow.UPFlow_net.config()
    with _quiet(quiet):
        conf.update(dict(TEST_PARAMS, **(params or {})))
        net = conf()
        if checkpoint:
            with cpu_checkpoint_load():
                net.load_model(CHECKPOINT, if_relax=True, if_print=False)
    net.eval()
    return net


@contextlib.contextmanager
def _quiet(on):
    if not on:
        yield
        return
    with open(os.devnull, "w") as dn, contextlib.redirect_stdout(dn):
        yield
