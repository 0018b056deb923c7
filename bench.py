#!/usr/bin/env python
"""Throughput of the UPFlow decoder hot path on B200 (BASELINE.json metric:
image-pairs/s at KITTI 1242x375, 6-level pyramid + SGU, batch 1 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = one two-frame forward (forward + backward flow, full pyramid, SGU)
of one image pair per GPU.  Prints ONE JSON line (rank 0).

  value      : pairs/s, inputs resident in HBM, CUDA-graph replay, device timed
  e2e        : the same through the public drop-in API `UPFlow_net(input_dict)`
               with PINNED HOST inputs: H2D of both frames and D2H of the flow
               inside the timed region, every step
  roofline   : the kernel with the largest share of the step, per-launch CUDA
               events (profiler.py); roofline_corr: the fused correlation kernel
               at the HD 1/4-resolution shape [2,32,270,480], d=4 (north star)
  cpu_baseline / --impl reference : the reference's CPU path (op-for-op port,
               oracle/ref_port.py -- the Python reference cannot travel to the
               GPU box) on the host cores, same workload, same weights
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {  # name -> (H, W, batch per GPU)
    "kitti_375x1242_b1": (375, 1242, 1),
    "sintel_436x1024_b8": (436, 1024, 8),
    "hd_1080x1920_b2": (1080, 1920, 2),
    "small_256x256_b1": (256, 256, 1),
    "train_256x832_b4": (256, 832, 4),        # BASELINE config 4: unsupervised training step, 4 pairs per GPU
}
METRIC = "image_pairs_per_sec"
UNIT = "pairs/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="MEASURED_PEAKS.json (measured)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="B200_PROFILING.md fallback")


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons of one GPU while the timed region runs (pynvml, 20 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_evt.set()
        if self.ok:
            self.join(timeout=1)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def make_weights(seed=1234):
    """Random-init weights of the reference architecture (its own MSRA init, model/pwc_modules.py:52-69), with small
    random biases; no checkpoint exists on the GPU box."""
    import upflow_pytorch_b200 as pkg
    torch.manual_seed(seed)
    net = pkg.build_model(device="cpu")
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(seed)
    for k in sd:
        if k.endswith("bias"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.02
        if "conv_last" in k or "context_networks.convs.6" in k:
            sd[k] = sd[k] * 0.1       # keep an untrained net's flows at a few pixels
    return sd


CHECKPOINT = os.path.join(ROOT, "tests", "golden", "upflow_kitti2015.pth")
GOLDEN_E2E = os.path.join(ROOT, "tests", "golden", "kitti_e2e.pt")


def load_weights():
    """BASELINE config 2 names the reference's shipped checkpoint (scripts/upflow_kitti2015.pth, loaded by
    test.py:31-38): its state dict travels to the GPU box as a committed fixture.  Falls back to random-init weights of
    the same architecture only if the fixture is missing.  Returns (state dict, description, checkpoint path or None)."""
    if os.path.exists(CHECKPOINT):
        return (torch.load(CHECKPOINT, weights_only=True), "upflow_kitti2015.pth (the reference's shipped checkpoint; "
                "tests/golden copy, loaded with net.load_model like test.py:31-38)", CHECKPOINT)
    return make_weights(), "random-init (MSRA, seed 1234): checkpoint fixture missing", None


def synth_inputs(B, H, W, seed):
    """Synthetic image pairs of the named resolution: a smooth random texture (bicubically upsampled
    uniform noise, range ~[-0.5, 0.5] like the KITTI preprocessing) and the same texture moved by (u, v) = (-3, +2)
    px.  seed 1234, B = 1 reproduces the pair of tests/golden/kitti_e2e.pt bit for bit."""
    g = torch.Generator().manual_seed(seed)
    lo = torch.rand(B, 3, H // 8 + 4, W // 8 + 4, generator=g)
    base = torch.nn.functional.interpolate(lo, size=(H + 16, W + 16), mode="bicubic", align_corners=False) - 0.5
    return base[:, :, 8:8 + H, 8:8 + W].contiguous(), base[:, :, 6:6 + H, 11:11 + W].contiguous()


def workload_config(workload, weights_desc):
    """The `config` object: what is measured, identical for the product arm and the reference arm."""
    H, W, B = WORKLOADS[workload]
    return {"workload": workload, "image": [H, W], "pairs_per_step_per_gpu": B, "pyramid_levels": 6, "decoder_levels": 5,
            "sgu": True, "directions": "forward + backward flow", "weights": weights_desc,
            "inputs": "synthetic textured pair with known motion (-3,+2) px, seed 1234 + rank",
            "l2": "GPU arm: L2 flushed between timed steps (a 256 MiB write before every step, outside the step's events; "
                  "`value_lanes`: a 160 MiB write per step inside the timed region; `e2e`: every step's inputs arrive from the host)"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    from oracle import ref_port as P
    H, W, B = WORKLOADS[args.workload]
    sd, wdesc, _ = load_weights()
    im1, im2 = synth_inputs(B, H, W, 1234)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    with torch.no_grad():
        for _ in range(max(1, args.warmup)):
            P.forward_2_frame(im1, im2, sd)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            P.forward_2_frame(im1, im2, sd)
        dt = time.perf_counter() - t0
    v = B * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(1, args.warmup), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, wdesc),
            "path": "oracle/ref_port.py: op-for-op CPU port of model/upflow.py forward_2_frame_v3 (F.conv2d / "
                    "F.grid_sample / unfold correlation), bit-identical to the reference on CPU",
            "host_processes": 1,
            "note": "ONE CPU process on all host cores whatever --gpus says (rank 0 runs, the other ranks exit): "
                    "compare with the product arm at N=1 only",
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d forwards of one %dx%d batch-%d pair after %d warm-up" % (args.steps, H, W, B, max(1, args.warmup))},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(json.dumps(line))


def build_net(params=None, precision="tf32", train=False):
    """the public drop-in model with the benchmark's weights, loaded the way test.py:31-38 does"""
    import upflow_pytorch_b200 as pkg
    sd, wdesc, ckpt = load_weights()
    net = pkg.build_model(params=params, state_dict=None if ckpt else sd, conv_precision=precision)
    if ckpt:
        net.load_model(ckpt, if_relax=True, if_print=False)
    return (net.train() if train else net.eval()), sd, wdesc


def measure_train(args, rank, world, local, dist, steps, warmup):
    """BASELINE config 4 (unsupervised training step on synthetic 832x256 pairs, 4 pairs per GPU = batch 32 over 8
    GPUs): one step = H2D of the local shard (pinned host), forward with the loss branch (photometric + edge-aware
    smoothness + multi-scale distillation), backward on this library's kernels, ONE all-reduce of the flat fp32
    gradient buffer over NCCL, Adam(amsgrad) update, D2H of the loss.  Returns the record (max over ranks)."""
    from upflow_pytorch_b200 import _ext
    from upflow_pytorch_b200.train import Trainer
    H, W, B = WORKLOADS["train_256x832_b4"]
    default_losses = args.train_losses == "all"
    params = {"if_use_boundary_warp": default_losses, "multi_scale_distillation_weight": 0.01}
    if default_losses:
        params["photo_loss_census_weight"] = 1.0
    net, sd, wdesc = build_net(params, args.precision, train=True)
    if args.torch_losses:
        from model.upflow import network_tools
        from utils.loss import loss_functions
        from utils.tools import tools
        network_tools.use_loss_kernels = loss_functions.use_loss_kernels = tools.boundary_dilated_warp.use_kernel = False
    tr = Trainer(net, use_cuda_graph=not args.no_train_graph)
    host = {}
    if default_losses:
        # the crop of a full KITTI frame, like dataset/kitti_dataset.py's random crop: the frames travel too
        RH, RW, y0, x0 = 375, 1242, 60, 200
        raw1, raw2 = synth_inputs(B, RH, RW, 1234 + rank)
        im1_h, im2_h = raw1[:, :, y0:y0 + H, x0:x0 + W].contiguous(), raw2[:, :, y0:y0 + H, x0:x0 + W].contiguous()
        host = {"im1_raw": raw1.pin_memory(), "im2_raw": raw2.pin_memory(),
                "start": torch.tensor([[x0, y0]] * B, dtype=torch.float32).reshape(B, 2, 1, 1).pin_memory()}
    else:
        im1_h, im2_h = synth_inputs(B, H, W, 1234 + rank)
    im1_h, im2_h = im1_h.pin_memory(), im2_h.pin_memory()
    host.update({"im1": im1_h, "im2": im2_h})
    h2d_bytes = sum(t.numel() * 4 for t in host.values())

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def step():
        loss = tr.train_step({k: t.cuda(non_blocking=True) for k, t in host.items()})
        return loss.item()                                   # D2H of the step's result

    for _ in range(warmup):
        step()
    sampler = ClockSampler(local)
    sync_all()
    sampler.start()
    n0 = _ext.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    sync_all()
    clocks = sampler.stop()
    launches = _ext.launch_count() - n0 + (tr.graph_launches * steps if tr.use_cuda_graph else 0)   # replayed kernels are not host launches
    ms = e0.elapsed_time(e1) / steps
    # the collective alone, and how much of it the step hides (Trainer issues it on a side stream under the Adam
    # update of ... nothing: it is 0.1 % of the step; reported as measured)
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    a0.record()
    for _ in range(5):
        nbytes = tr.grads.all_reduce_mean()
    a1.record()
    sync_all()
    ar_ms = a0.elapsed_time(a1) / 5
    if dist is not None:
        t = torch.tensor([ms, ar_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ar_ms = float(t[0].item()), float(t[1].item())
    v = world * B / (ms * 1e-3)
    rec = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
           "global_batch": world * B, "dtype": args.precision,
           "config": dict(workload_config("train_256x832_b4", wdesc),
                          step="forward + loss (%s) + backward + gradient all-reduce + Adam(amsgrad)" % (
                              "boundary-dilated-warp photo abs_robust on the 375x1242 frames, census 1.0, edge smooth, msd 0.01"
                              if default_losses else "photo abs_robust, edge smooth, msd 0.01"),
                          losses="torch expressions (A/B)" if args.torch_losses else "loss kernels (csrc/loss.cu)",
                          launch="eager" if args.no_train_graph else "zero-grad + forward + losses + backward replayed as one CUDA graph; all-reduce and Adam eager",
                          l2="per-step working set (> 1 GB of activations) exceeds the 126 MB L2",
                          parallelism="data parallel x%d, one all-reduce of %d fp32 gradients per step" % (world, tr.grads.numel)),
           "e2e": {"value": v, "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": h2d_bytes,
                   "d2h_bytes_per_step": 4, "api": "upflow_pytorch_b200.train.Trainer.train_step(batch) with pinned host tensors"},
           "allreduce": {"bytes": nbytes, "ms": ar_ms, "share_of_step": ar_ms / ms,
                         "comm": "NCCL all-reduce (sum) of one flat fp32 buffer" if world > 1 else "none (1 GPU)"},
           "gpu_launches": launches, "launches_per_step": launches // steps, "clocks": clocks, "final_loss": loss}
    del tr, net
    torch.cuda.empty_cache()
    return rec


def run_train(args, rank, world, local, dist):
    """--workload train_256x832_b4: the training record as the bench line."""
    rec = measure_train(args, rank, world, local, dist, args.steps, max(3, args.warmup))
    if rank == 0:
        ref = None
        if world == 1 and not args.no_cpu_baseline:
            ref = reference_gpu_train_step(args)
        line = dict(rec, higher_is_better=True, scaling="weak", vs_baseline=None, data="synthetic")
        if ref is not None:
            line["reference_gpu_path"] = ref
        emit(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def reference_gpu_train_step(args, steps=5):
    """The anchor of the training step: the op-for-op port of the reference (oracle/ref_port.py + ref_port_train.py:
    F.conv2d through cuDNN with TF32 allowed, F.grid_sample, unfold correlation, the loss branch as torch expressions)
    with torch autograd and torch.optim.Adam(amsgrad), eager, on the same GPU, same shard, host tensors in, loss out.
    Reported next to the result, never part of it."""
    try:
        from oracle import ref_port_train as PT
        H, W, B = WORKLOADS["train_256x832_b4"]
        sd, _, _ = load_weights()
        params = {k: v.clone().cuda().requires_grad_() for k, v in sd.items()}
        opt = torch.optim.Adam(list(params.values()), lr=1e-4, amsgrad=True, weight_decay=1e-4)
        im1_h, im2_h = (t.pin_memory() for t in synth_inputs(B, H, W, 1234))

        def step():
            opt.zero_grad(set_to_none=True)
            loss = PT.training_loss(im1_h.cuda(non_blocking=True), im2_h.cuda(non_blocking=True), params,
                                    msd_weight=0.01)["loss"]
            loss.backward()
            opt.step()
            return loss.item()
        for _ in range(2):
            step()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        r0.record()
        for _ in range(steps):
            loss = step()
        r1.record()
        torch.cuda.synchronize()
        ms = r0.elapsed_time(r1) / steps
        out = {"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "kind": "port", "final_loss": loss,
               "path": "oracle/ref_port.py + ref_port_train.py on cuda:0: eager PyTorch %s autograd, cuDNN (allow_tf32=%s), "
                       "torch.optim.Adam(amsgrad)" % (torch.__version__, torch.backends.cudnn.allow_tf32)}
        del params, opt
        torch.cuda.empty_cache()
        return out
    except Exception as exc:   # pragma: no cover - the comparison leg must never break the bench line
        return {"error": repr(exc)[:300]}


def epe_vs_reference(args, net, H, W, B):
    """Mean end-point error of the benchmarked precision against the REFERENCE's own CPU flow on the same pair, shipped
    weights (tests/golden/kitti_e2e.pt, made by oracle/make_golden_kitti.py from the unmodified reference): under the
    reference's `mask >= 1.0` (bounded below by its own noise floor, stored with the fixture) and under the robust-mask
    diagnostic (threshold 0.9999 on both sides), which is the number the 1e-3 px target is read on."""
    if not (os.path.exists(GOLDEN_E2E) and os.path.exists(CHECKPOINT)):
        return None
    case = [c for c in torch.load(GOLDEN_E2E, weights_only=False) if (c["H"], c["W"]) == (H, W)]
    if not case:
        return None
    c = case[0]
    im1, im2 = synth_inputs(1, H, W, c["seed"])
    eng = net._get_engine()
    out = {"weights": "upflow_kitti2015.pth", "precision": args.precision, "reference": "UPFlow_net.forward_2_frame_v3 of the "
           "unmodified reference on CPU (fp32), fixture tests/golden/kitti_e2e.pt", "noise_floor_of_the_reference_px": c["noise_floor_px"]}

    def epe(a, b):
        return torch.sqrt(((a - b) ** 2).sum(1)).mean().item()
    with torch.no_grad():
        for name, thr, key in (("robust_mask", 0.9999, "flow_f_robust"), ("real_mask", 1.0, "flow_f_reference")):
            eng.mask = True if thr == 1.0 else thr
            f, _, _ = eng.forward(im1.cuda(), im2.cuda())
            out[name] = epe(f.cpu(), c[key])
    eng.mask = True
    return out


def corr_roofline(pk, iters=20):
    """Fused correlation+LeakyReLU kernel alone at the HD 1/4-res shape: algorithmic bytes
    4*N*h*w*(2C+81) / CUDA-event time, L2 flushed before every launch."""
    from upflow_pytorch_b200 import ops
    N, C, h, w, d = 2, 32, 270, 480, 4
    g = torch.Generator().manual_seed(1)
    f1 = torch.randn(N, h, w, C, generator=g).cuda()
    f2 = torch.randn(N, h, w, C, generator=g).cuda()
    out = torch.empty(N, h, w, 81, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ops.k_corr(f1, f2, out, d, slope=0.1)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.k_corr(f1, f2, out, d, slope=0.1)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    nbytes = 4 * N * h * w * (2 * C + 81)
    ach = nbytes / (ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("corr_pipe_kernel", {}).get("dram_bytes_per_launch_avg")

    def timed(fn):
        for _ in range(3):
            fn()
        tt = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            tt.append(e0.elapsed_time(e1))
        return sum(tt) / len(tt) * 1e3
    # the same cost volume through the other entry points: the reference operator's own layout (NCHW in, NCHW out,
    # corr_planar.cu) and the fp16 / bf16 storage variants (half the bytes, fp32 arithmetic)
    variants = {}
    try:
        p1, p2 = f1.permute(0, 3, 1, 2).contiguous(), f2.permute(0, 3, 1, 2).contiguous()
        po = torch.empty(N, 81, h, w, device="cuda")
        us = timed(lambda: ops.k_corr_planar(p1, p2, po, d, slope=0.1))
        variants["planar_nchw_fp32"] = {"kernel": "corr_planar_kernel<4>", "us_per_launch": us, "algorithmic_bytes": nbytes,
                                        "frac": nbytes / us / 1e3 / pk["hbm"]}
        for name, dt in (("bf16_storage", torch.bfloat16), ("fp16_storage", torch.float16)):
            a, b, o = f1.to(dt), f2.to(dt), torch.empty(N, h, w, 81, device="cuda", dtype=dt)
            us = timed(lambda: ops.k_corr_lp(a, b, o, d, slope=0.1))
            variants[name] = {"kernel": "corr_fwd_kernel<4, %s>" % str(dt).split(".")[1], "us_per_launch": us,
                              "algorithmic_bytes": nbytes // 2, "frac": nbytes / 2 / us / 1e3 / pk["hbm"]}
    except Exception as exc:   # pragma: no cover - side measurements must never break the bench line
        variants["error"] = repr(exc)[:200]
    return {"kernel": "corr_pipe_kernel<4>", "shape": [N, C, h, w], "max_disp": d, "bound": "hbm", "achieved": ach,
            "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": traffic, "us_per_launch": ms * 1e3,
            "best_us": min(ts) * 1e3, "algorithmic_bytes": nbytes, "l2": "flushed before every launch", "variants": variants}


class _StdoutGuard:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    communicator creation), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to
    the saved descriptor."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        self.out = os.fdopen(os.dup(self.saved), "w")
        self._print = print
        return self

    def emit(self, text):
        self.out.write(text + "\n")
        self.out.flush()

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        self.out.close()
        return False


_GUARD = None


def emit(text):
    if _GUARD is not None:
        _GUARD.emit(text)
    else:
        print(text)


def main():
    global _GUARD
    with _StdoutGuard() as g:
        _GUARD = g
        try:
            _main()
        finally:
            _GUARD = None


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kitti_375x1242_b1", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32", "tf32x3"])
    ap.add_argument("--lanes", type=int, default=4,
                    help="e2e: pairs in flight in pipeline.PipelinedInference (1 = one pair at a time)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the `train` sub-record (BASELINE config 4) of the default line")
    ap.add_argument("--no-train-graph", action="store_true", help="training workload: eager launches instead of a CUDA graph")
    ap.add_argument("--train-losses", default="plain", choices=["plain", "all"],
                    help="training workload: 'plain' = photo + smooth + msd (the line BASELINE config 4 is measured on); "
                         "'all' = every loss term of model/upflow.py:394-491: the boundary-dilated warp on the un-cropped 375x1242 frames "
                         "(the model's default if_use_boundary_warp=True) + census weight 1")
    ap.add_argument("--torch-losses", action="store_true",
                    help="training workload: the loss branch as torch expressions instead of the loss kernels (A/B)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps > 60:
            args.steps = 60          # bounded CPU sample (~1-2.5 s per KITTI pair on the host cores)
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.workload.startswith("train"):
        if args.steps > 20:
            args.steps = 20
        run_train(args, rank, world, local, dist)
        return
    W_ = max(3, args.warmup)
    K = args.steps
    H, W, B = WORKLOADS[args.workload]
    pk = peaks()

    from upflow_pytorch_b200 import _ext, profiler
    from upflow_pytorch_b200.pipeline import PipelinedInference
    if os.environ.get("UPF_WIN_DEBUG"):          # triage runs only: "mode,min_cin,force" for upf_debug_conv_win
        _ext.load().upf_debug_conv_win(*[int(x) for x in os.environ["UPF_WIN_DEBUG"].split(",")])
    net, sd, wdesc = build_net(None, args.precision)                         # public API object (drop-in UPFlow_net)
    im1_h, im2_h = synth_inputs(B, H, W, 1234 + rank)
    im1_h, im2_h = im1_h.pin_memory(), im2_h.pin_memory()
    im1_d, im2_d = im1_h.cuda(), im2_h.cuda()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- value: graph replay, inputs resident in HBM
    eng = net._get_engine()
    with torch.no_grad():
        graphed = eng.capture(B, H, W)
    graphed.im1.copy_(im1_d)
    graphed.im2.copy_(im2_d)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")    # > 126 MB L2
    for _ in range(W_):
        graphed.replay()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    evs = []
    t_wall0 = time.perf_counter()
    for _ in range(K):
        flush.zero_()                                   # L2 flush, outside the per-step events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graphed.replay()
        e1.record()
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    step_ms = sum(a.elapsed_time(b) for a, b in evs) / K
    step_ms = max_over_ranks(step_ms)
    value = world * B / (step_ms * 1e-3)

    # the same K steps with `lanes` pairs in flight (what e2e below does, inputs resident): lane i owns a stream, a graph
    # and a workspace set; every step flushes L2 on its own stream INSIDE the timed region (counted against the result)
    conc = None
    if args.lanes > 1:
        with torch.no_grad():
            lane_graphs = []
            for i in range(args.lanes):
                eng.lane = i
                g = eng.capture(B, H, W)
                g.im1.copy_(im1_d)
                g.im2.copy_(im2_d)
                lane_graphs.append(g)
            eng.lane = 0
        streams = [torch.cuda.Stream() for _ in range(args.lanes)]
        flushes = [torch.empty(160 << 20, dtype=torch.uint8, device="cuda") for _ in range(args.lanes)]

        def run_lanes(n):
            main = torch.cuda.current_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            for st in streams:
                st.wait_event(e0)
            for k in range(n):
                i = k % args.lanes
                with torch.cuda.stream(streams[i]):
                    flushes[i].zero_()
                    lane_graphs[i].replay()
            for st in streams:
                main.wait_stream(st)
            e1.record(main)
            return e0, e1
        run_lanes(W_ * args.lanes)
        barrier()
        e0, e1 = run_lanes(K)
        barrier()
        conc_ms = max_over_ranks(e0.elapsed_time(e1) / K)
        conc = {"lanes": args.lanes, "value": world * B / (conc_ms * 1e-3), "unit": UNIT, "ms_per_step": conc_ms,
                "l2": "every step writes a 160 MiB buffer on its own stream before its replay, inside the timed region",
                "note": "K complete forwards, `lanes` of them in flight on separate streams, graphs and workspaces; "
                        "`value` above is the same forward one pair at a time"}
        del flushes

    # ---------------- e2e: public API, pinned host in, host out, every step (pipeline.py): the host->device copy of
    # pair k+1 and the device->host copy of flow k-1 run on copy streams under the forwards; every step still moves its
    # own inputs and its own result.  `lanes` pairs are in flight: consecutive pairs replay their graphs on separate
    # streams and workspaces, so one pair's latency-bound coarse levels run next to another pair's fine levels.  Measured
    # with one lane too (a pair at a time, the round-1 / round-2 arrangement) and both are printed.
    def run_e2e(lanes):
        pipe = PipelinedInference(net, lanes=lanes)
        for _ in range(max(W_, 2 * lanes)):
            pipe.submit(im1_h, im2_h)
        pipe.drain()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(pipe.h2d)                              # the first thing a timed step does is its H2D copy
        for _ in range(K):
            pipe.submit(im1_h, im2_h)
        flow_host = pipe.drain()[-1]                     # the last result is on the host when the clock stops
        e1.record(pipe.copy)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3 / K
        return max_over_ranks(max(e0.elapsed_time(e1) / K, wall_ms)), flow_host   # the slower of the device and host clocks

    e2e_1_ms, flow_host = run_e2e(1)
    e2e_ms, lanes = e2e_1_ms, 1
    if args.lanes > 1:
        e2e_ms, flow_host = run_e2e(args.lanes)
        lanes = args.lanes
    e2e = {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": 2 * im1_h.numel() * 4, "d2h_bytes_per_step": flow_host.numel() * 4,
           "lanes": lanes,
           "single_lane": {"value": world * B / (e2e_1_ms * 1e-3), "ms_per_step": e2e_1_ms},
           "api": "UPFlow_net(input_dict)['flow_f_out'] (drop-in model.upflow) behind upflow_pytorch_b200.pipeline."
                  "PipelinedInference(net, lanes=%d): pinned host tensors in, pinned host flow out, every pair its own "
                  "H2D copy, complete forward (one CUDA-graph replay) and D2H copy; %d pair(s) in flight on separate "
                  "streams and workspaces, copies on two more streams (K steps timed from the first H2D to the last "
                  "flow on the host; results bit-identical to the one-pair-at-a-time call, "
                  "tests/test_gpu_engine.py::test_pipelined_inference_returns_every_flow_in_order)" % (lanes, lanes)}
    # the same call without the pipeline (copy in, forward, copy out, wait), for the record
    out_h = torch.empty(B, 2, H, W).pin_memory()
    with torch.no_grad():
        def serial_step():
            o = net({"im1": im1_h.cuda(non_blocking=True), "im2": im2_h.cuda(non_blocking=True), "if_loss": False})
            out_h.copy_(o["flow_f_out"], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        for _ in range(3):
            serial_step()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(min(K, 20)):
            serial_step()
        s1.record()
        torch.cuda.synchronize()
    e2e["serial_ms_per_step"] = s0.elapsed_time(s1) / min(K, 20)

    # ---------------- the training step of BASELINE config 4 on the same ranks (4 pairs per GPU, NCCL all-reduce)
    train = None
    if args.workload == "kitti_375x1242_b1" and not args.no_train:
        train = measure_train(args, rank, world, local, dist, 10, 3)

    line = None
    if rank == 0:
        # ---------------- roofline: per-launch events over one eager forward (spin kernel lets the host run ahead)
        with torch.no_grad():
            eng.overlap = False          # one stream: per-launch times are not mixed with concurrent side-stream work
            eng.forward(im1_d, im2_d)
            with profiler.record() as rec:
                eng.forward(im1_d, im2_d)
            eng.overlap = True
        agg = rec.by_kernel()           # keyed by the kernel family that ran (upf_last_kernel)
        total_ms = sum(a["ms"] for a in agg.values())
        dom = max(agg, key=lambda k: agg[k]["ms"])
        a = agg[dom]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")     # dram bytes per launch from committed ncu captures
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(dom + "_kernel", {}).get("dram_bytes_per_launch_avg")
        if dom in ("conv_win", "conv_halo", "conv_tc"):
            tf32_peak = pk["bf16"] / 2.0
            ach = a["flops"] / (a["ms"] * 1e-3) / 1e12
            roof = {"kernel": dom + "_kernel", "bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": ach / tf32_peak, "traffic": traffic,
                    "peak_source": pk["source"] + ": bf16_tflops/2 (dense TF32 rate is half the bf16 rate)",
                    "launches_per_step": a["launches"], "ms_per_step": a["ms"], "share_of_step": a["ms"] / total_ms,
                    "algorithmic_flops_per_step": a["flops"],
                    "timing": "CUDA events around every launch of one eager forward (profiler.py), minus the measured "
                              "cost of an event pair (%.1f us), summed over this kernel's launches" % (rec.overhead_ms * 1e3)}
        elif dom.startswith("conv"):
            ach = a["flops"] / (a["ms"] * 1e-3) / 1e12
            roof = {"kernel": dom + "_kernel", "bound": "tensor", "achieved": ach, "peak": 74.4, "unit": "TFLOP/s",
                    "frac": ach / 74.4, "traffic": traffic, "peak_source": "fp32 SIMT 148 SM x 128 lanes x 2 x 1.965 GHz",
                    "launches_per_step": a["launches"], "ms_per_step": a["ms"], "share_of_step": a["ms"] / total_ms,
                    "algorithmic_flops_per_step": a["flops"]}
        else:
            ach = a["bytes"] / (a["ms"] * 1e-3) / 1e9
            roof = {"kernel": dom + "_kernel", "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": ach / pk["hbm"], "traffic": traffic, "peak_source": pk["source"],
                    "launches_per_step": a["launches"], "ms_per_step": a["ms"], "share_of_step": a["ms"] / total_ms}
        def _frac(k, v):
            # fraction of the roofline that bounds the family: TF32 tensor peak for the tcgen05 convolutions, HBM otherwise
            if v["ms"] <= 0:
                return None
            if k in ("conv_win", "conv_halo", "conv_tc"):
                return round(v["flops"] / (v["ms"] * 1e-3) / 1e12 / (pk["bf16"] / 2.0), 4)
            if k in ("conv_simt", "conv_c3"):
                return round(v["flops"] / (v["ms"] * 1e-3) / 1e12 / 74.4, 4)
            return round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / pk["hbm"], 4)
        breakdown = {k: {"launches": v["launches"], "ms": round(v["ms"], 4), "share": round(v["ms"] / total_ms, 4),
                         "GB/s": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else None,
                         "TFLOP/s": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2) if v["ms"] > 0 else None,
                         "roofline_frac": _frac(k, v)}
                     for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
        rc = corr_roofline(pk)

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref_port as P
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            a1, a2 = im1_h.clone(), im2_h.clone()
            reps = 3 if H * W <= 600 * 1300 else 1
            with torch.no_grad():
                P.forward_2_frame(a1, a2, sd)
                t0 = time.perf_counter()
                for _ in range(reps):
                    rf = P.forward_2_frame(a1, a2, sd)[0]
                dt = (time.perf_counter() - t0) / reps
                f, _, _ = eng.forward(im1_d, im2_d)
            from oracle import cpu_oracle as O
            cpu = {"value": B / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "%d forward(s) of the same %dx%d batch-%d pair after 1 warm-up, %d torch threads" % (reps, H, W, B, cores),
                   "epe_cuda_vs_cpu_port_px": O.epe(f.cpu(), rf)}
        epe = epe_vs_reference(args, net, H, W, B)

        # ---------------- the reference's GPU path (the >= 10x target is against it): the op-for-op
        # port of model/upflow.py (F.conv2d / cuDNN, F.grid_sample, unfold correlation) in eager PyTorch on this GPU,
        # same weights, same inputs, host tensors in and out like e2e.  Reported next to the result, never part of it.
        ref_gpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import ref_port as P
                sd_d = {k: v.cuda() for k, v in sd.items()}
                with torch.no_grad():
                    def ref_step():
                        rf = P.forward_2_frame(im1_h.cuda(non_blocking=True), im2_h.cuda(non_blocking=True), sd_d)[0]
                        out_h.copy_(rf, non_blocking=True)
                        torch.cuda.current_stream().synchronize()
                    for _ in range(3):
                        ref_step()
                    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    r0.record()
                    for _ in range(10):
                        ref_step()
                    r1.record()
                    torch.cuda.synchronize()
                rms = r0.elapsed_time(r1) / 10
                ref_gpu = {"value": B / (rms * 1e-3), "unit": UNIT, "ms_per_step": rms, "kind": "port",
                           "path": "oracle/ref_port.py on cuda:0 (eager PyTorch %s, cuDNN, allow_tf32=%s), pinned host in / host out"
                                   % (torch.__version__, torch.backends.cudnn.allow_tf32),
                           "e2e_speedup_over_it": e2e["value"] / (B / (rms * 1e-3))}
                del sd_d
                torch.cuda.empty_cache()
            except Exception as exc:   # pragma: no cover - the comparison leg must never break the bench line
                ref_gpu = {"error": repr(exc)[:200]}

        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W_,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic",
                "config": workload_config(args.workload, wdesc),
                "timing": {"l2": "flushed (256 MiB write) before every timed step; flush outside the step events",
                           "launch": "one CUDA graph replay per step, both flow directions stacked in one batch",
                           "parallelism": "batch-sharded x%d, no collective in inference" % world},
                "e2e": e2e, "gpu_launches": graphed.launches * K, "launches_per_step": graphed.launches,
                "clocks": clocks, "roofline": roof, "roofline_corr": rc, "kernel_breakdown_ms": breakdown,
                "wall_s_timed_region": wall}
        if conc is not None:
            line["value_lanes"] = conc
        if epe is not None:
            line["epe_vs_reference_px"] = epe
        if train is not None:
            if world == 1 and not args.no_cpu_baseline:
                train["reference_gpu_path"] = reference_gpu_train_step(args)
                if "value" in train["reference_gpu_path"]:
                    train["reference_gpu_path"]["speedup_over_it"] = train["value"] / train["reference_gpu_path"]["value"]
            line["train"] = train
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if ref_gpu is not None:
            line["reference_gpu_path"] = ref_gpu
        emit(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
