#!/usr/bin/env python
"""bench.py -- headline benchmark of the HyGrid hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2half|c4]

Workload (BASELINE.json configs[1], "c2"): batched rect->hex bilinear resampling of 256 x 3 x 1024 x 1024
float32 synthetic images onto a 1024 x 1024 hex lattice (float32 result, HG_MATH_FAST).  One "step" = one
pass of the hot path over that batch.  Metric: hex Mpix/s = hex-lattice pixels produced per second
(N * H1 * W1, channels not multiplied).

Printed JSON line (rank 0):
  value      device-resident throughput: inputs already in HBM, K steps bracketed by barrier + synchronize,
             CUDA events on the launching stream, max over ranks.
  e2e        same metric through the C ABI host entry point hg_host_rect2hex with pinned HOST buffers:
             H2D of the step's inputs and D2H of its result are inside the timed region.
  roofline   dominant kernel (rect2hex_bilinear): algorithmic bytes (24 B per hex pixel: 3 channels x
             (4 B read + 4 B written)) / mean launch time (CUDA events around every launch) vs the measured
             HBM copy bandwidth in MEASURED_PEAKS.json (fallback 6650 GB/s, B200_PROFILING.md).
  cpu_baseline  the numpy oracle port of the reference (oracle/hygrid_oracle.py) timed on the host cores on a
             bounded sample of the same workload (rank 0, N=1 only).
  extra      the other BASELINE configs, measured in the same run, each with its own roofline block:
             c2_exact_f64  the drop-in numpy call's arithmetic (HG_MATH_EXACT, float64 result): device-resident and end to
                           end through hg_host_rect2hex with host buffers (36 B per hex pixel: 3 x (4 B read + 8 B written))
             c4            64 x 3 x 2160 x 3840 sharded over the ranks (64 / N images each, "strong"): rect->hex bilinear,
                           hex->rect linear (fast and exact) and the 5-level average-pool pyramid
             c3            4 x HexConv2d(64, 64, r=2) on 256 x 256, batch 128 per GPU, autocast bf16, fwd + bwd: the reference's
                           float32 tensors at the layer boundary, and bfloat16 activations end to end (out_dtype opt-in)
             c5            hex CNN training step, 64 images per GPU on 128 x 128 lattices, with the NCCL all-reduce of the
                           flat gradient bucket INSIDE the timed region (overlapped with backward)
Multi-GPU: one process per GPU (torchrun); the batch is sharded by image, no data-path collective
(resampling shards with no communication) -> "scaling": "weak" (256 images per GPU).

`--impl reference` times the reference's CPU implementation of the path (the numpy oracle port -- the
reference is pure Python and cannot travel to the GPU box) with all host cores (one process per core over
a bounded sample of images per step).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "hex Mpix/s (rect->hex bilinear resample)"
UNIT = "hex Mpix/s"

WORKLOADS = {
    # name: (images per GPU, channels, H, W, h1, w1)
    "c2": (256, 3, 1024, 1024, 1024, 1024),
    "c2half": (256, 3, 1024, 1024, 512, 512),
    "c4": (64, 3, 2160, 3840, 2160, 3840),
}


def workload_name(wl):
    n, c, h, w, h1, w1 = WORKLOADS[wl]
    return f"rect->hex bilinear, {n}x{c}x{h}x{w} float32 -> {h1}x{w1} hex lattice (BASELINE configs[1])" if wl == "c2" \
        else f"rect->hex bilinear, {n}x{c}x{h}x{w} float32 -> {h1}x{w1}"


def algorithmic_bytes(wl, images):
    """s*C*N*(min(H*W, taps*H1*W1) + H1*W1), s = 4 bytes, 4 taps."""
    _, c, h, w, h1, w1 = WORKLOADS[wl]
    return 4 * c * images * (min(h * w, 4 * h1 * w1) + h1 * w1)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(wl):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(wl)
    except Exception:
        return None


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every ~2 ms on a thread, so that even a 20 ms timed region
    holds samples of its own (B200_PROFILING.md clocks line; nvidia-smi -lms 100 cannot resolve it).  `window(t0, t1)`
    only ever reports samples taken inside [t0, t1]; an empty window is reported as such, never padded with others."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.max_mhz, self.err = index, [], False, None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.index
            if vis:
                try:
                    idx = int(vis.split(",")[self.index])
                except Exception:
                    idx = self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self.stop_flag:
                    try:
                        mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                        try:
                            bits = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                        except Exception:
                            bits = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        self.rows.append((time.perf_counter(), float(mhz), int(bits)))
                    except Exception as e:                      # noqa: BLE001
                        self.err = repr(e)
                        return
                    time.sleep(0.002)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
        except Exception as e:                                  # noqa: BLE001
            self.err = repr(e)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)

    def window(self, t0, t1):
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        sm = sorted(r[1] for r in rows)
        reasons = sorted({name for _, _, bits in rows for name, mask in self.REASONS if bits & mask})
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
               "samples": len(sm), "window_ms": (t1 - t0) * 1e3, "source": "NVML, 2 ms period, samples inside the window only"}
        if self.err:
            out["error"] = self.err
        return out


# ----------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference)
# ----------------------------------------------------------------------------------------------------
def _cpu_one(args):
    import numpy as np
    from oracle import hygrid_oracle as O
    seed, c, h, w, h1, w1 = args
    img = (np.random.default_rng(seed).random((c, h, w), dtype=np.float32) * 255).astype(np.float32)
    t = time.perf_counter()
    out = O.rect_to_hex_resample(img, (h1, w1), "bilinear")
    dt = time.perf_counter() - t
    return dt, float(out[0, h1 // 2, w1 // 2])


def host_info():
    """CPU model and library versions next to every CPU number."""
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except Exception:
        pass
    import numpy as np
    return {"cpu_model": model, "os_cpu_count": os.cpu_count(), "numpy": np.__version__}


def cpu_baseline(wl, images=8):
    """Single-process numpy port on `images` images of the workload (about 5-30 s of CPU work)."""
    _, c, h, w, h1, w1 = WORKLOADS[wl]
    t = time.perf_counter()
    for i in range(images):
        _cpu_one((i, c, h, w, h1, w1))
    dt = time.perf_counter() - t
    inner = images * h1 * w1 / dt / 1e6
    return {"value": inner, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{images} images of {c}x{h}x{w} float32 -> {h1}x{w1}, oracle/hygrid_oracle.rect_to_hex_resample "
                      f"(numpy restatement of geometry_np.py:358-519), single process, {dt:.1f} s", "host": host_info()}


def run_reference(args):
    """--impl reference: the reference's CPU path (numpy oracle port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    wl = args.workload
    n_img, c, h, w, h1, w1 = WORKLOADS[wl]
    cores = os.cpu_count() or 1
    per_step = cores            # one image per core per step: bounded sample of the 256-image batch
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        jobs = [(i, c, h, w, h1, w1) for i in range(per_step)]
        for _ in range(args.warmup):
            pool.map(_cpu_one, jobs)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_one, jobs)
        dt = time.perf_counter() - t0
    value = args.steps * per_step * h1 * w1 / dt / 1e6
    sample = (f"{per_step} images of {c}x{h}x{w} float32 per step (of the {n_img}-image batch), one process per core, "
              f"numpy oracle port of geometry_np.rect_to_hex_resample")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(wl), "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "host": host_info()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def bind_near_gpu(local):
    """Pin this rank to the CPUs of its GPU's NUMA node (sysfs local_cpulist) so that the pinned host buffers
    of the end-to-end leg are allocated next to the PCIe root the GPU hangs off; returns a short description."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{dev}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            return f"rank bound to {len(use)} CPUs local to GPU {dev}"
        return f"GPU {dev}: local CPUs = all allowed CPUs ({len(allowed)})"
    except Exception as e:                                    # no sysfs / no permission: keep the default placement
        return f"no NUMA binding ({type(e).__name__})"


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def bf16_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops_sustained"]), "measured sustained (MEASURED_PEAKS.json)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md)"


class Ctx:
    """What every leg of the GPU arm needs: rank / world, device, the barrier and a max-over-ranks reduction."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa = bind_near_gpu(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.peak, self.peak_src = measured_peak()
        self.graphs = []

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def timed(self, fn, steps, warmup=3):
        """(ms per step over the bracketed region, mean ms per step from per-step events), max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        t0.record()
        for a, b in ev:
            a.record(); fn(); b.record()
        t1.record()
        self.barrier()
        total = t0.elapsed_time(t1) / steps
        per = sum(a.elapsed_time(b) for a, b in ev) / steps
        return self.max_over_ranks(total, per)

    def hbm(self, nbytes, ms):
        a = nbytes / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": self.peak, "unit": "GB/s", "frac": a / self.peak, "traffic": None,
                "peak_source": self.peak_src, "algorithmic_bytes_per_launch": nbytes, "kernel_ms": ms}


def headline(cx, sampler):
    """C2 (BASELINE configs[1]): device-resident `value`, roofline of the dominant kernel, fast-float32 `e2e`."""
    import ctypes as C
    import numpy as np
    torch, args = cx.torch, cx.args
    from HyGrid import _native as nv
    from HyGrid import functional as Fn
    wl = args.workload
    n_img, c, h, w, h1, w1 = WORKLOADS[wl]
    math_mode = args.math
    torch.manual_seed(1234 + cx.rank)
    x = torch.rand(n_img, c, h, w, device=cx.dev, dtype=torch.float32) * 255
    y = torch.empty(n_img, c, h1, w1, device=cx.dev, dtype=torch.float32)

    def step():
        Fn.rect_to_hex(x, (h1, w1), "bilinear", out_dtype=torch.float32, math=math_mode, out=y)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    cx.barrier()
    time.sleep(0.05)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nv.reset_launch_count()
    cx.barrier()
    w0 = time.perf_counter()
    t_start.record()
    for a, b in ev:
        a.record(); step(); b.record()
    t_end.record()
    cx.barrier()
    w1_ = time.perf_counter()
    launches = nv.launch_count()
    kernel_name = nv.last_launch()
    total_ms = t_start.elapsed_time(t_end)
    kernel_ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)

    # ---- end to end through the C ABI host entry point (pinned host buffers, H2D + D2H timed) --------
    e2e_steps = max(1, args.e2e_steps)
    hx = torch.empty((n_img, c, h, w), dtype=torch.float32).pin_memory()
    hy = torch.empty((n_img, c, h1, w1), dtype=torch.float32).pin_memory()
    hx.copy_(x)
    xs, ys = (np.ascontiguousarray(v) for v in Fn.coordinate_tables("rect2hex", h, w, h1, w1, "np", cx.dev)[2:])

    def e2e_step():
        nv.call("hg_host_rect2hex", C.c_void_p(hx.data_ptr()), C.c_void_p(hy.data_ptr()), C.c_void_p(xs.ctypes.data),
                C.c_void_p(ys.ctypes.data), n_img * c, h, w, h1, w1, nv.F32, nv.F32, 1,
                nv.MATH_FAST if math_mode == "fast" else nv.MATH_EXACT, cx.local)

    e2e_step()
    cx.barrier()
    nv.reset_launch_count()
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    cx.barrier()
    e1 = time.perf_counter()
    e2e_launches = nv.launch_count()
    ok = bool(torch.equal(hy[:2].to(cx.dev), y[:2]))       # the host path must reproduce the device path
    total_ms, kernel_ms, e2e_ms = cx.max_over_ranks(total_ms, kernel_ms, (e1 - e0) * 1e3)
    line = None
    if cx.rank == 0:
        world = cx.world
        pix = n_img * h1 * w1
        value = world * pix * args.steps / (total_ms * 1e-3) / 1e6
        e2e_value = world * pix * e2e_steps / (e2e_ms * 1e-3) / 1e6
        achieved = algorithmic_bytes(wl, n_img) / (kernel_ms * 1e-3) / 1e9
        clocks = sampler.window(w0, w1_)
        clocks["e2e"] = {k: v for k, v in sampler.window(e0, e1).items() if k in ("sm_mhz", "reasons", "samples", "window_ms")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(wl), "images_per_gpu": n_img, "global_images": n_img * world,
                       "math": math_mode, "out_dtype": "float32", "parallelism": f"image-sharded x{world}, no collective",
                       "l2": "inputs (3.2 GB) and outputs (3.2 GB) exceed the 126 MB L2; no flush needed", "numa": cx.numa},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": cx.peak, "unit": "GB/s", "frac": achieved / cx.peak,
                         "traffic": ncu_traffic(wl), "peak_source": cx.peak_src,
                         "kernel": kernel_name, "algorithmic_bytes_per_launch": algorithmic_bytes(wl, n_img),
                         "kernel_ms": kernel_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": hx.numel() * 4, "d2h_bytes_per_step": hy.numel() * 4,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "api": "hg_host_rect2hex (C ABI, pinned host buffers)",
                    "matches_device_path": ok, "gpu_launches": int(e2e_launches)},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
    del x, y, hx, hy
    torch.cuda.empty_cache()
    return line


# ---- extra legs: every entry is a dict with its own roofline; a failing leg records its error and the run goes on ----
def extra_c2_exact_f64(cx):
    """The arithmetic and result type of the drop-in call geometry_np.rect_to_hex_resample(img, None, 'bilinear')
    (geometry_np.py:358-519): float64 blend in the reference's operation order, float64 result -- device-resident and
    through hg_host_rect2hex with pinned host buffers (6.4 GB come back per step instead of 3.2)."""
    import ctypes as C
    import numpy as np
    torch = cx.torch
    from HyGrid import _native as nv
    from HyGrid import functional as Fn
    n_img, c, h, w, h1, w1 = WORKLOADS["c2"]
    torch.manual_seed(99 + cx.rank)
    x = torch.rand(n_img, c, h, w, device=cx.dev) * 255
    y = torch.empty(n_img, c, h1, w1, device=cx.dev, dtype=torch.float64)
    total, per = cx.timed(lambda: Fn.rect_to_hex(x, None, "bilinear", out_dtype=torch.float64, math="exact", out=y), 10)
    nbytes = c * n_img * (4 * h * w + 8 * h1 * w1)
    pix = n_img * h1 * w1
    out = {"workload": "C2 with HG_MATH_EXACT and a float64 result (bit-identical to geometry_np.rect_to_hex_resample)",
           "value": cx.world * pix / (total * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": total, "dtype": "f64",
           "roofline": cx.hbm(nbytes, per)}
    steps = max(1, min(cx.args.e2e_steps, 10))
    hx = torch.empty((n_img, c, h, w), dtype=torch.float32).pin_memory()
    hy = torch.empty((n_img, c, h1, w1), dtype=torch.float64).pin_memory()
    hx.copy_(x)
    xs, ys = (np.ascontiguousarray(v) for v in Fn.coordinate_tables("rect2hex", h, w, h1, w1, "np", cx.dev)[2:])

    def e2e_step():
        nv.call("hg_host_rect2hex", C.c_void_p(hx.data_ptr()), C.c_void_p(hy.data_ptr()), C.c_void_p(xs.ctypes.data),
                C.c_void_p(ys.ctypes.data), n_img * c, h, w, h1, w1, nv.F32, nv.F64, 1, nv.MATH_EXACT, cx.local)
    e2e_step()
    cx.barrier()
    e0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    cx.barrier()
    (ms,) = cx.max_over_ranks((time.perf_counter() - e0) * 1e3 / steps)
    out["e2e"] = {"value": cx.world * pix / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms, "steps": steps,
                  "h2d_bytes_per_step": hx.numel() * 4, "d2h_bytes_per_step": hy.numel() * 8,
                  "api": "hg_host_rect2hex(HG_F32 -> HG_F64, HG_MATH_EXACT): what HyGrid.geometry_np.rect_to_hex_resample calls",
                  "matches_device_path": bool(torch.equal(hy[:2].to(cx.dev), y[:2]))}
    return out


def extra_c4(cx):
    """BASELINE configs[3]: 64 x 3 x 2160 x 3840, sharded over the ranks (strong scaling, no collective)."""
    torch = cx.torch
    from HyGrid import functional as Fn
    from HyGrid import HexFrames as hf
    from HyGrid.distributed import shard_range
    lo, hi = shard_range(64, cx.rank, cx.world)
    n, c, h, w = hi - lo, 3, 2160, 3840
    torch.manual_seed(400 + cx.rank)
    x = torch.rand(n, c, h, w, device=cx.dev)
    y = torch.empty_like(x)
    px, total_px = h * w, 64 * h * w
    legs = {}

    def leg(name, fn, nbytes, steps=10):
        total, per = cx.timed(fn, steps)
        legs[name] = {"value": total_px / (total * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": total, "roofline": cx.hbm(nbytes, per)}
    leg("rect_to_hex_bilinear_fast", lambda: Fn.rect_to_hex(x, None, "bilinear", out_dtype=torch.float32, math="fast", out=y), 8 * n * c * px)
    leg("hex_to_rect_linear_fast", lambda: Fn.hex_to_rect(x, None, "linear", out_dtype=torch.float32, math="fast", twin="np", out=y), 8 * n * c * px)
    leg("hex_to_rect_linear_exact", lambda: Fn.hex_to_rect(x, None, "linear", out_dtype=torch.float32, math="exact", twin="np", out=y), 8 * n * c * px)
    pool = hf.HexPool2d("average", 2, 2)
    nb, hh, ww = 0, h, w
    for _ in range(5):
        hn, wn = (hh - 2) // 2 + 1, (ww - 1) // 2
        nb += 4 * n * c * (hh * ww + hn * wn)
        hh, ww = hn, wn

    def pyramid():
        cur = x
        for _ in range(5):
            cur = pool(cur)
        return cur
    leg("pool_pyramid_5_levels_average", pyramid, nb)
    return {"workload": f"C4: 64x3x2160x3840 float32 over {cx.world} GPU(s), {n} images on rank 0", "scaling": "strong", "legs": legs}


def extra_c3(cx):
    """BASELINE configs[2]: 4 x HexConv2d(64, 64, r=2, padding=1), 256 x 256 lattice, batch 128 per GPU, autocast bf16,
    loss = y.float().sum(), forward + backward through the modules.  SURVEY.md 8d per layer-step: 6.44 GB (bf16
    activations: fwd, dgrad, wgrad each read / write 2 x 1.07 GB) and 1.443e12 FLOP."""
    torch = cx.torch
    from HyGrid import HexFrames as hf
    N, layers, HW = 128, 4, 256
    tf_peak, tf_src = bf16_peak()
    out = {"workload": f"C3: {layers} x HexConv2d(64,64,r=2) on {HW}x{HW}, batch {N} per GPU, autocast bf16, fwd+bwd", "modes": {}}
    pix = N * HW * HW
    alg_bytes = layers * 3 * 2 * (64 * pix * 2)          # per step, bf16 activations
    flops = (layers * 3 - 1) * 2 * 7 * 64 * 64 * pix      # the first layer's data gradient is not needed (x is a leaf input)
    for mode, odt in (("float32 layer boundary (reference semantics)", torch.float32), ("bfloat16 activations end to end", torch.bfloat16)):
        torch.manual_seed(300)
        net = torch.nn.Sequential(*[hf.HexConv2d(64, 64, 0, 2, stride=1, padding=1) for _ in range(layers)]).to(cx.dev)
        for l in net:
            l.out_dtype = odt
        x = torch.randn(N, 64, HW, HW, device=cx.dev).to(odt)

        def step():
            for p in net.parameters():
                p.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = net(x)
            y.float().sum().backward()
        total, _ = cx.timed(step, 5)
        a_hbm = alg_bytes / (total * 1e-3) / 1e9
        a_tf = flops / (total * 1e-3) / 1e12
        out["modes"][mode] = {
            "ms_per_step": total, "value": cx.world * pix / (total * 1e-3) / 1e6, "unit": UNIT,
            "hex_mpix_per_s_per_layer_step": cx.world * pix * layers / (total * 1e-3) / 1e6,
            "roofline": {"bound": "hbm", "achieved": a_hbm, "peak": cx.peak, "unit": "GB/s", "frac": a_hbm / cx.peak, "traffic": None,
                         "algorithmic_bytes_per_step": alg_bytes, "peak_source": cx.peak_src},
            "tensor": {"bound": "tensor", "achieved": a_tf, "peak": tf_peak, "unit": "TFLOP/s", "frac": a_tf / tf_peak,
                       "flop_per_step": flops, "peak_source": tf_src}}
        del net, x
        torch.cuda.empty_cache()
    return out


def extra_c5(cx):
    """BASELINE configs[4]: hex CNN training step, 64 images per GPU on 128 x 128 hex lattices, autocast bf16, SGD; the
    flat gradient bucket is all-reduced over NCCL inside the timed region, group by group while backward still runs."""
    torch = cx.torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from hexcnn import HexCNN
    from HyGrid import _native as nv
    from HyGrid.distributed import FlatGradBucket
    B, HW = 64, 128
    torch.manual_seed(0)                                   # identical initial weights on every rank
    model = HexCNN().to(cx.dev)
    g = torch.Generator(device="cpu").manual_seed(500 + cx.rank)
    x = torch.randn(B, 3, HW, HW, generator=g).to(cx.dev)
    t = torch.randint(0, 10, (B,), generator=g).to(cx.dev)
    out = {"workload": f"C5: hex CNN (3->32->64->128, BN+ReLU, max pool, GAP, Linear) train step, {B} img/GPU x {cx.world} GPU, "
                       f"{HW}x{HW}, autocast bf16, SGD momentum", "scaling": "weak", "variants": {}}
    # groups by layer: {c1: conv kernel + BN affine}, {c2: ...}, {c3: ... + the classifier}; the last group is complete --
    # and its all-reduce in flight -- after the first third of the backward pass
    for name, kw in (("overlapped (3 layer groups, all-reduce issued from backward)", dict(groups=[3, 3, 5], overlap=True)),
                     ("blocking (one all-reduce after backward)", dict(groups=1, overlap=False))):
        bucket = FlatGradBucket(model.parameters(), **kw)
        opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, fused=True)

        def step():
            bucket.zero_()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = torch.nn.functional.cross_entropy(model(x).float(), t)
            loss.backward()
            bucket.finish()
            opt.step()
            return loss
        nv.reset_launch_count()
        total, _ = cx.timed(step, 20, warmup=5)
        launches = nv.launch_count() // 25
        check = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
        lo_, hi_ = check.clone(), check.clone()
        if cx.world > 1:
            cx.dist.all_reduce(lo_, op=cx.dist.ReduceOp.MIN)
            cx.dist.all_reduce(hi_, op=cx.dist.ReduceOp.MAX)
        out["variants"][name] = {"ms_per_step": total, "images_per_s": cx.world * B / (total * 1e-3),
                                 "value": cx.world * B * HW * HW / (total * 1e-3) / 1e6, "unit": UNIT,
                                 "grad_bucket_bytes": bucket.nbytes, "group_bytes": list(bucket.group_bytes()),
                                 "library_launches_per_step": int(launches),
                                 "ranks_in_sync": bool(torch.allclose(lo_, hi_, rtol=0, atol=1e-6 * float(hi_.abs()) + 1e-9)),
                                 "collective": "NCCL all-reduce inside the timed region" if cx.world > 1 else "single GPU: no collective"}
        bucket.detach()
    if cx.args.c5_graph:
        # The eager step is host-bound (about 85 launches of a few microseconds against 1.1 ms of GPU work): every rank replays the
        # WHOLE step -- forward, backward, the NCCL all-reduce of the flat bucket, SGD -- as one captured CUDA graph.  Two captures:
        # one all-reduce after backward, and the three layer groups reduced from inside backward (parallel branches of the graph:
        # only the first layer's 3 KB group is left to reduce when backward ends).
        for gname, kw in (("whole step replayed as one CUDA graph (all-reduce captured)", dict(groups=1, overlap=False)),
                          ("whole step as one CUDA graph, 3 layer groups all-reduced from inside backward", dict(groups=[3, 3, 5], overlap=True))):
            if cx.world == 1 and kw["overlap"]:
                continue                                   # no collective on one GPU: the two captures are the same graph
            try:
                bucket = FlatGradBucket(model.parameters(), **kw)
                opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, fused=True)

                def step():
                    bucket.zero_()
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        loss = torch.nn.functional.cross_entropy(model(x).float(), t)
                    loss.backward()
                    bucket.finish()
                    opt.step()
                    return loss
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    for _ in range(3):
                        step()
                torch.cuda.current_stream().wait_stream(side)
                cx.barrier()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    step()
                cx.barrier()
                total, _ = cx.timed(graph.replay, 50, warmup=5)
                check = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
                lo_, hi_ = check.clone(), check.clone()
                if cx.world > 1:
                    cx.dist.all_reduce(lo_, op=cx.dist.ReduceOp.MIN)
                    cx.dist.all_reduce(hi_, op=cx.dist.ReduceOp.MAX)
                out["variants"][gname] = {
                    "ms_per_step": total, "images_per_s": cx.world * B / (total * 1e-3), "value": cx.world * B * HW * HW / (total * 1e-3) / 1e6,
                    "unit": UNIT, "ranks_in_sync": bool(torch.allclose(lo_, hi_, rtol=0, atol=1e-6 * float(hi_.abs()) + 1e-9)),
                    "collective": "NCCL all-reduce inside the graph" if cx.world > 1 else "single GPU: no collective"}
                cx.graphs.append(graph)                    # kept alive until the process leaves (see run_ours)
                bucket.detach()
            except Exception as e:                          # noqa: BLE001
                out["variants"][gname] = {"error": f"{type(e).__name__}: {e}"[:300]}
    return out


def run_ours(args):
    cx = Ctx(args)
    from HyGrid import _native as nv
    nv.lib()
    sampler = ClockSampler(cx.local)
    if cx.rank == 0:
        sampler.start()
    line = headline(cx, sampler)
    extra = {}
    if not args.no_extra:
        for name, fn in (("c2_exact_f64", extra_c2_exact_f64), ("c4", extra_c4), ("c3", extra_c3), ("c5", extra_c5)):
            if args.extra and name not in args.extra.split(","):
                continue
            t0 = time.perf_counter()
            try:
                extra[name] = fn(cx)
            except Exception as e:                          # noqa: BLE001 -- a failing leg must not take the headline down
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
                cx.torch.cuda.empty_cache()
            if cx.rank == 0 and isinstance(extra[name], dict):
                extra[name]["leg_wall_s"] = round(time.perf_counter() - t0, 2)
                extra[name]["clocks"] = {k: v for k, v in sampler.window(t0, time.perf_counter()).items()
                                         if k in ("sm_mhz", "sm_max_mhz", "reasons", "samples")}
    sampler.stop()
    if cx.rank == 0:
        line["extra"] = extra
        if cx.world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_images)
        print(json.dumps(line), flush=True)
    nv.lib().hg_host_release()
    if cx.world > 1:
        cx.dist.barrier()
        if cx.graphs:
            # tearing the process group down while a captured graph still holds NCCL kernels hung an 8-GPU run in round 1
            # (after the result line was out): the line above is flushed, leave without the collective teardown
            cx.torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)
        cx.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--math", default="fast", choices=["fast", "exact"])
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs (extra block)")
    ap.add_argument("--no-c5-graph", dest="c5_graph", action="store_false", help="skip the CUDA-graph variant of the C5 step")
    ap.add_argument("--extra", default="", help="comma list of extra legs to run (c2_exact_f64,c4,c3,c5); default all")
    ap.add_argument("--cpu-images", type=int, default=48)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
