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
import subprocess
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
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [v.strip() for v in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= t1)] or [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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
def run_ours(args):
    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_near_gpu(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from HyGrid import _native as nv
    from HyGrid import functional as Fn
    nv.lib()

    wl = args.workload
    n_img, c, h, w, h1, w1 = WORKLOADS[wl]
    math_mode = args.math
    torch.manual_seed(1234 + rank)
    x = torch.rand(n_img, c, h, w, device=dev, dtype=torch.float32) * 255
    y = torch.empty(n_img, c, h1, w1, device=dev, dtype=torch.float32)

    def step():
        Fn.rect_to_hex(x, (h1, w1), "bilinear", out_dtype=torch.float32, math=math_mode, out=y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nv.reset_launch_count()
    barrier()
    w0 = time.perf_counter()
    t_start.record()
    for a, b in ev:
        a.record(); step(); b.record()
    t_end.record()
    barrier()
    w1_ = time.perf_counter()
    launches = nv.launch_count()
    total_ms = t_start.elapsed_time(t_end)
    kernel_ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)

    # ---- end to end through the C ABI host entry point (pinned host buffers, H2D + D2H timed) --------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    hx = torch.empty((n_img, c, h, w), dtype=torch.float32).pin_memory()
    hy = torch.empty((n_img, c, h1, w1), dtype=torch.float32).pin_memory()
    hx.copy_(x)
    xs, ys = (np.ascontiguousarray(v) for v in Fn.coordinate_tables("rect2hex", h, w, h1, w1, "np", dev)[2:])

    def e2e_step():
        nv.call("hg_host_rect2hex", C.c_void_p(hx.data_ptr()), C.c_void_p(hy.data_ptr()), C.c_void_p(xs.ctypes.data),
                C.c_void_p(ys.ctypes.data), n_img * c, h, w, h1, w1, nv.F32, nv.F32, 1,
                nv.MATH_FAST if math_mode == "fast" else nv.MATH_EXACT, local)

    e2e_step()
    barrier()
    nv.reset_launch_count()
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - e0
    e2e_launches = nv.launch_count()
    ok = bool(torch.equal(hy[:2].to(dev), y[:2]))       # the host path must reproduce the device path
    clocks = sampler.stop(w0, w1_) if rank == 0 else None

    times = torch.tensor([total_ms, kernel_ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms, e2e_ms = (float(v) for v in times.tolist())

    if rank == 0:
        pix = n_img * h1 * w1
        value = world * pix * args.steps / (total_ms * 1e-3) / 1e6
        e2e_value = world * pix * e2e_steps / (e2e_ms * 1e-3) / 1e6
        peak, peak_src = measured_peak()
        achieved = algorithmic_bytes(wl, n_img) / (kernel_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(wl), "images_per_gpu": n_img, "global_images": n_img * world,
                       "math": math_mode, "out_dtype": "float32", "parallelism": f"image-sharded x{world}, no collective",
                       "l2": "inputs (3.2 GB) and outputs (3.2 GB) exceed the 126 MB L2; no flush needed", "numa": numa},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(wl), "peak_source": peak_src, "kernel": "rect2hex_bilinear_ws_kernel" if math_mode == "fast" else "rect2hex_bilinear_tma_kernel",
                         "algorithmic_bytes_per_launch": algorithmic_bytes(wl, n_img), "kernel_ms": kernel_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": hx.numel() * 4, "d2h_bytes_per_step": hy.numel() * 4,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps, "api": "hg_host_rect2hex (C ABI, pinned host buffers)",
                    "matches_device_path": ok, "gpu_launches": int(e2e_launches)},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(wl, args.cpu_images)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    nv.lib().hg_host_release()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--math", default="fast", choices=["fast", "exact"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-images", type=int, default=48)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
