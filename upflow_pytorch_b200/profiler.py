"""Per-launch device timing of the library's kernels with CUDA events (used by
bench.py for the roofline object and by profiles/ summaries).

Wraps the `ops.k_*` launchers: every launch is bracketed by two events on the
launching stream.  A spin kernel is queued first so the host runs ahead and
the bracketed kernels execute back to back (event deltas then exclude launch
latency).  Algorithmic bytes / flops follow SURVEY.md section 8(d)."""
import contextlib

import torch

from . import _ext, ops


def _npix(s):
    return s.N * s.H * s.W


def _account(name, args, kwargs):
    S = ops._as_slice
    if name == "k_corr":
        f1, out = S(args[0]), S(args[2])
        n = _npix(f1)
        return dict(bytes=4 * n * (2 * f1.C + out.C), flops=2 * out.C * f1.C * n, shape=(f1.N, f1.C, f1.H, f1.W))
    if name == "k_warp":
        x, out = S(args[0]), S(args[2])
        return dict(bytes=4 * _npix(out) * (2 * x.C + 2), flops=8 * x.C * _npix(out), shape=(out.N, x.C, out.H, out.W))
    if name == "k_stats":
        x = S(args[0])
        return dict(bytes=4 * _npix(x) * x.C, flops=3 * _npix(x) * x.C, shape=(x.N, x.C, x.H, x.W))
    if name == "k_conv":
        x, out = S(args[0]), S(args[3])
        k = args[4]
        n = _npix(out)
        return dict(bytes=4 * (_npix(x) * x.C + n * out.C) + 4 * k * k * x.C * out.C, flops=2 * n * k * k * x.C * out.C,
                    shape=(x.N, x.C, x.H, x.W, out.C, k), tc=((kwargs.get("precision", args[9] if len(args) > 9 else 0) & 0xFF) == 1))
    if name == "k_conv_chain":
        layers = args[0]
        N, H, W = layers[0]._shape
        n = N * H * W
        return dict(bytes=sum(4 * n * (L.Cin + L.Cout) + 4 * L.ksize * L.ksize * L.Cin * L.Cout for L in layers),
                    flops=sum(2 * n * L.ksize * L.ksize * L.Cin * L.Cout for L in layers), shape=(N, H, W, "%d layers" % len(layers)),
                    tc=True)
    if name == "k_resize":
        a, out = S(args[0]), S(args[1])
        return dict(bytes=4 * a.C * (_npix(a) + _npix(out)), flops=8 * a.C * _npix(out), shape=(out.N, a.C, out.H, out.W))
    if name == "k_sgu_blend":
        inter, out = S(args[1]), S(args[2])
        return dict(bytes=4 * (4 * _npix(out) + 3 * _npix(inter)), flops=40 * _npix(out), shape=(out.N, out.H, out.W))
    if name == "k_occ_check":
        f = S(args[0])
        return dict(bytes=4 * _npix(f) * 7, flops=30 * _npix(f), shape=(f.N, f.H, f.W))
    if name == "k_tap_combine":
        y, out = S(args[0]), S(args[2])
        return dict(bytes=4 * _npix(out) * (y.C + out.C), flops=9 * _npix(out) * out.C, shape=(out.N, out.C, out.H, out.W))
    if name == "k_copy":
        a = S(args[1])                                    # the destination (the source may be None: zero fill)
        return dict(bytes=(8 if args[0] is not None else 4) * _npix(a) * a.C, flops=0, shape=(a.N, a.C, a.H, a.W))
    return dict(bytes=0, flops=0, shape=())


NAMES = ("k_corr", "k_warp", "k_stats", "k_conv", "k_conv_chain", "k_resize", "k_sgu_blend", "k_copy", "k_norm_apply", "k_tap_combine", "k_occ_check",
         "k_norm_combine")


def event_overhead_ms(n=200):
    """What a pair of events adds to one launch: n tiny launches bracketed one by one, minus the same n launches
    bracketed once, per launch.  (Measured on B200: ~4 us -- as much as a small kernel itself, so per-launch tables
    of a forward with 150 launches would otherwise over-state the families with many small launches.)"""
    a = torch.zeros(1, 1, 4, 4, device="cuda")
    b = torch.zeros_like(a)
    fn = ops.k_copy
    for _ in range(20):
        fn(a, b)
    torch.cuda._sleep(20_000_000)
    evs = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(a, b)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    each = sum(x.elapsed_time(y) for x, y in evs) / n
    torch.cuda._sleep(20_000_000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn(a, b)
    e1.record()
    torch.cuda.synchronize()
    together = e0.elapsed_time(e1) / n
    return max(0.0, each - together)


class Recorder:
    def __init__(self):
        self.records = []
        self.overhead_ms = 0.0

    def finalize(self):
        torch.cuda.synchronize()
        for r in self.records:
            raw = r.pop("e0").elapsed_time(r.pop("e1"))
            r["ms_raw"] = raw
            r["ms"] = max(raw - self.overhead_ms, 0.1 * raw)
        return self.records

    def by_kernel(self):
        """aggregate by the kernel family that actually ran (upf_last_kernel)"""
        agg = {}
        for r in self.records:
            name = r.get("kernel") or r["name"]
            a = agg.setdefault(name, dict(launches=0, ms=0.0, bytes=0, flops=0))
            a["launches"] += 1
            a["ms"] += r["ms"]
            a["bytes"] += r["bytes"]
            a["flops"] += r["flops"]
        return agg


@contextlib.contextmanager
def record(spin_cycles=20_000_000, subtract_event_overhead=True):
    rec = Recorder()
    if subtract_event_overhead:
        rec.overhead_ms = event_overhead_ms()
    saved = {n: getattr(ops, n) for n in NAMES}

    def wrap(name, fn):
        def inner(*args, **kwargs):
            meta = _account(name, args, kwargs)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(*args, **kwargs)
            e1.record()
            k = _ext.load().upf_last_kernel()
            meta.update(name=name, e0=e0, e1=e1, kernel=k.decode() if k else name)
            rec.records.append(meta)
        return inner

    for n, fn in saved.items():
        setattr(ops, n, wrap(n, fn))
    try:
        if spin_cycles:
            torch.cuda._sleep(int(spin_cycles))
        yield rec
    finally:
        for n, fn in saved.items():
            setattr(ops, n, fn)
        rec.finalize()
