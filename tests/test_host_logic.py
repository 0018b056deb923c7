"""CPU tier: host-side logic of the package -- drop-in API surface, state-dict
contract, the weight permutation that maps the reference's prepend-order dense
blocks onto append-only buffers, loud failure without CUDA."""
import os
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

from oracle import ref_port as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dropin():
    import upflow_pytorch_b200 as pkg
    pkg.install_dropin()
    import model.upflow as up
    import model.pwc_modules as pm
    assert up.__file__.startswith(pkg.DROPIN_DIR)
    return pkg, up, pm


def test_state_dict_contract(dropin):
    """80 tensors, the reference's names and shapes"""
    pkg, up, pm = dropin
    net = pkg.build_model(device="cpu")
    sd = net.state_dict()
    shapes = P.reference_param_shapes()
    assert set(sd) == set(shapes)
    assert all(tuple(sd[k].shape) == shapes[k] for k in sd)
    assert sum(v.numel() for v in sd.values()) == 3494549
    net.load_state_dict(P.det_state_dict(1))           # strict load works


def test_api_surface(dropin):
    pkg, up, pm = dropin
    for name in ("conv", "initialize_msra", "upsample2d_flow_as", "upsample_flow", "FlowEstimatorDense_v2",
                 "ContextNetwork_v2_", "WarpingLayer_no_div", "FeatureExtractor"):
        assert hasattr(pm, name)
    assert hasattr(up, "UPFlow_net") and hasattr(up.network_tools, "sgu_model")
    from model.correlation_package.correlation import Correlation, CorrelationFunction
    from utils.pytorch_correlation import Corr_pyTorch
    from utils.tools import tools
    Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1)
    Corr_pyTorch(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1)
    with pytest.raises(AssertionError):
        Corr_pyTorch(pad_size=3, max_displacement=4)          # same argument check as the reference
    with pytest.raises(NotImplementedError):
        Correlation(pad_size=20, kernel_size=3, max_displacement=20, stride1=1, stride2=2)
    conf = up.UPFlow_net.config()
    assert conf.if_sgu_upsample is False and conf.if_use_cor_pytorch is False       # reference defaults
    name = conf.get_name(print_now=False)
    assert "if_sgu_upsample|False_" in name
    assert isinstance(tools.occ_check_model(obj_out_all='obj'), object)


def test_no_cpu_fallback(dropin):
    pkg, up, pm = dropin
    net = pkg.build_model(device="cpu")
    x = torch.zeros(1, 3, 64, 64)
    with pytest.raises(RuntimeError):
        net({"im1": x, "im2": x, "if_loss": False})
    from upflow_pytorch_b200 import ops
    with pytest.raises(RuntimeError):
        ops.warp(torch.zeros(1, 4, 8, 8), torch.zeros(1, 2, 8, 8))
    if not torch.cuda.is_available():
        from upflow_pytorch_b200.engine import DecoderEngine
        with pytest.raises(RuntimeError):
            DecoderEngine(P.det_state_dict(0))


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "upflow_pytorch_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                if "import oracle" in src or "from oracle" in src or "/root/reference" in src:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_dense_block_weight_permutation():
    """Emulate the engine's append-only X buffer with torch convs on CPU using the PACKED weights and compare with
    the reference-order dense block: validates _dense_slots / pack_conv_weight (pure host logic)."""
    from upflow_pytorch_b200 import ops
    from upflow_pytorch_b200.engine import EST_CH, X_FLOW2, X_LD, X_OFF, _dense_slots
    sd = P.det_state_dict(11)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 115, 7, 9, generator=g)
    x5_ref, out_ref = P.dense(x, sd, "flow_estimators")
    flow2 = torch.randn(1, 2, 7, 9, generator=g)
    ctx_ref = P._conv(torch.cat([x5_ref, flow2], 1), sd, "context_networks.convs.0")

    X = torch.zeros(1, X_LD, 7, 9)
    X[:, :115] = x
    names = ("conv1", "conv2", "conv3", "conv4", "conv5", "conv_last")
    widths = (128,) + tuple(o + c for o, c in zip(X_OFF, EST_CH))

    def run(key, slots, width, relu=True):
        w, _ = ops.pack_conv_weight(sd[key + ".0.weight"], slots, width)          # [9, width, cout_pad]
        cout = sd[key + ".0.bias"].numel()
        wt = w[:, :, :cout].reshape(3, 3, width, cout).permute(3, 2, 0, 1).contiguous()
        y = F.conv2d(X[:, :width], wt, sd[key + ".0.bias"], padding=1)
        return F.leaky_relu(y, 0.1) if relu else y

    for k, name in enumerate(names):
        y = run("flow_estimators." + name, _dense_slots(115, range(115), X_OFF, EST_CH, k), widths[k], relu=k < 5)
        if k < 5:
            X[:, X_OFF[k]:X_OFF[k] + EST_CH[k]] = y
        else:
            assert (y - out_ref).abs().max().item() < 1e-5
    X[:, X_FLOW2:X_FLOW2 + 2] = flow2
    slots = _dense_slots(115, range(115), X_OFF, EST_CH, 5) + [X_FLOW2, X_FLOW2 + 1]
    y = run("context_networks.convs.0", slots, X_LD)
    assert (y - ctx_ref).abs().max().item() < 1e-5
    # x5 in reference order can be read back from the buffer
    x5 = torch.cat([X[:, X_OFF[k]:X_OFF[k] + EST_CH[k]] for k in (4, 3, 2, 1, 0)] + [X[:, :115]], 1)
    assert (x5 - x5_ref).abs().max().item() < 1e-5


def test_reference_test_py_constructs_against_dropin(tmp_path):
    """Drop-in acceptance on the host side: with our `model`/`utils` ahead of the reference's on
    sys.path, the reference's own test.py imports and builds Test_model (CPU, no forward).  Needs the reference
    tree, so it only runs in the authoring container."""
    from oracle import ref_shims
    if not ref_shims.have_reference():
        pytest.skip("reference tree not mounted")
    code = r'''
import sys, types, importlib.machinery, contextlib, io
sys.path.insert(0, %r)
import upflow_pytorch_b200 as pkg
for name in ("imageio", "png", "tensorflow"):
    m = types.ModuleType(name); m.__spec__ = importlib.machinery.ModuleSpec(name, None); sys.modules[name] = m
import torch.utils.data.dataloader as dl
dl._DataLoaderIter = dl._BaseDataLoaderIter
sys.path.insert(0, %r)            # the reference (provides test.py and dataset/)
pkg.install_dropin()              # ours first
# dataset.kitti_dataset needs a few names our utils.tools does not carry at import time: none expected
import test as ref_test
ref_test.if_cuda = False
import model.upflow
assert model.upflow.__file__.startswith(pkg.DROPIN_DIR), model.upflow.__file__
with contextlib.redirect_stdout(io.StringIO()):
    tm = ref_test.Test_model(pretrain_path=%r)
n = sum(p.numel() for p in tm.net_work.parameters())
assert n == 3494549, n
print("OK", type(tm.net_work).__module__)
''' % (ROOT, ref_shims.REFERENCE_ROOT, ref_shims.CHECKPOINT)
    r = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    assert "OK model.upflow" in r.stdout


def test_kitti_evaluation_helpers(tmp_path):
    """metrics of dataset/kitti_dataset.py:464-499 on a case worked by hand, and the 16-bit KITTI PNG round trip
    (write: utils/tools.py:1516-1525, read: dataset/kitti_dataset.py:130-149)."""
    import numpy as np
    import torch
    from upflow_pytorch_b200 import evaluation as E
    gt = torch.zeros(1, 2, 2, 3)
    gt[0, 0] = torch.tensor([[10.0, 0.0, 100.0], [0.0, 0.0, 0.0]])
    pred = gt.clone()
    pred[0, 0, 0, 0] += 4.0          # error 4 > max(3, 0.5): outlier
    pred[0, 1, 0, 2] += 4.0          # error 4 < max(3, 5): not an outlier
    pred[0, 0, 1, 1] += 2.0          # error 2: not an outlier
    pred[0, 0, 1, 2] += 50.0         # masked out
    mask = torch.ones(1, 1, 2, 3)
    mask[0, 0, 1, 2] = 0
    assert abs(E.flow_error_avg(pred, gt, mask).item() - 10.0 / 5) < 1e-5
    assert abs(E.outlier_pct(gt, pred, mask).item() - 100.0 / 5) < 1e-4
    epe_all, f1, epe_noc, epe_occ = E.evaluate(pred, gt, mask, gt, mask * 0 + torch.tensor([[1.0, 1, 0], [1, 0, 0]]))
    assert abs(epe_all - 2.0) < 1e-5 and abs(f1 - 20.0) < 1e-4
    assert abs(epe_noc - 4.0 / 3) < 1e-5 and abs(epe_occ - 6.0 / 2) < 1e-5
    rng = np.random.RandomState(0)
    flow = np.round(rng.randn(7, 9, 2) * 20 * 64) / 64.0          # representable at 1/64 px
    valid = (rng.rand(7, 9) > 0.3).astype(np.uint16)
    p = str(tmp_path / "flow.png")
    E.write_kitti_png_file(p, flow, valid)
    f2, m2 = E.read_png_flow(p)
    assert f2.shape == (2, 7, 9) and m2.shape == (1, 7, 9)
    assert np.array_equal(np.transpose(f2, [1, 2, 0]), flow) and np.array_equal(m2[0], valid.astype(np.uint8))


def test_bench_reference_arm_prints_one_json_line():
    """bench.py contract on the CPU tier: `--impl reference` runs the reference's CPU path (the port) and prints
    exactly one JSON line with the keys the driver reads."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "small_256x256_b1", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 0 and d["config"]["workload"] == "small_256x256_b1"


def test_bench_reference_arm_under_torchrun_prints_once():
    """The driver launches the reference arm like the product arm, torchrun included: with two ranks, rank 0 alone
    runs and prints the line, the other rank exits 0 without work."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--workload", "small_256x256_b1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, r.stdout[-800:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0


def test_pipelined_inference_bookkeeping_with_fake_streams(monkeypatch):
    """pipeline.PipelinedInference, host logic only (streams, events and pinning faked, tensors on the CPU): results come
    back in submission order `lanes` submits later, drain() hands out the rest, every pair is one call of the public model
    on lane k % lanes, and a result slot is not reused before the caller could have read it (lanes + 2 slots)."""
    import torch
    from upflow_pytorch_b200 import pipeline

    class FakeEvent:
        def record(self, stream=None): pass
        def synchronize(self): pass

    class FakeStream:
        def __init__(self, device=None): pass
        def wait_event(self, ev): pass

    class FakeCtx:
        def __init__(self, s): pass
        def __enter__(self): return self
        def __exit__(self, *a): return False

    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", FakeCtx)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None)

    class FakeNet:
        def __init__(self):
            self.lane, self.calls = 0, []
        def set_lane(self, k):
            self.lane = k
        def __call__(self, d):
            self.calls.append(self.lane)
            return {"flow_f_out": d["im1"][:, :2] + d["im2"][:, :2]}

    pairs = [(torch.full((1, 3, 4, 5), float(i)), torch.full((1, 3, 4, 5), 10.0 * i)) for i in range(9)]
    for lanes in (1, 2, 4):
        net = FakeNet()
        pipe = pipeline.PipelinedInference(net, device="cpu", lanes=lanes)
        got, held = [], None
        for i, (a, b) in enumerate(pairs):
            r = pipe.submit(a, b)
            assert (r is None) == (i < lanes)
            if held is not None:                      # the flow handed out one submit ago is still intact (valid until the next-but-one)
                assert torch.equal(held[0], torch.full((1, 2, 4, 5), 11.0 * held[1]))
            held = None
            if r is not None:
                got.append(r.clone())
                held = (r, i - lanes)
        rest = pipe.drain()
        assert len(rest) == lanes and pipe.drain() == [] and pipe.flush() is None
        got += [r.clone() for r in rest]
        assert [float(g[0, 0, 0, 0]) for g in got] == [11.0 * i for i in range(9)]
        assert net.calls == ([i % lanes for i in range(9)] if lanes > 1 else [0] * 9)
        assert net.lane == 0                          # drain() puts the model back on lane 0
