#!/usr/bin/env python
"""Headline benchmark: sliced single-amplitude contraction of the 53-qubit depth-14
Sycamore-like random circuit (BASELINE.json configs[3], the config the metric is quoted on).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c4]

One *step* = the hot path over one batch of slices: ``--slices-per-step`` slices of the
plan (default 16, fixed for every N: strong scaling), split over the N ranks as
``compute_slices`` does, each rank running its slices through the fused
permute+contract kernels, followed by the single NCCL all-reduce of the partials.
``value`` = algorithmic TFLOP/s (8 flops per complex MAC) of the
whole job with the leaf tensors already resident in HBM; ``e2e`` = the same through
the public call with the leaves in pinned host memory (packed H2D inside the timed
region) and the amplitude read back to the host.  The JSON line also extrapolates the
wall time of the full contraction (all slices) from the measured slice rate.

``--impl reference`` times the reference's CPU path -- restated: the same kind of plan
executed node by node with ``numpy.tensordot`` (what quimb/cotengra do for numpy
arrays; neither is installable here) -- on a bounded sample of slices.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--slices-per-step", type=int, default=16)
    ap.add_argument("--no-mps", action="store_true", help="skip the MPS gates/s measurement (N=1 only)")
    ap.add_argument("--cpu-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--plan-seconds", type=float, default=60.0)
    return ap.parse_args()


# ----------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region (recipe's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc = device, None
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.proc.wait()
        self.file.flush()
        rows = [r.strip().split(", ") for r in open(self.file.name) if r.strip()]
        os.unlink(self.file.name)
        sm = [float(r[1]) for r in rows if len(r) >= 9]
        reasons = set()
        for r in rows:
            if len(r) < 9:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), r[5:9]):
                if val.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(rows[0][2]) if rows and len(rows[0]) >= 3 else None,
                "power_w_max": max((float(r[3]) for r in rows if len(r) >= 9), default=None),
                "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(kernel, plan):
    """DRAM bytes per launch of this node, if an ncu capture of it is on file (profiles/ncu_traffic.json:
    entries keyed by kernel name and the node's algorithmic flops and bytes)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            table = json.load(f)
    except (OSError, ValueError):
        return None
    for e in table.get("entries", []):
        if (e["kernel"] == kernel and abs(e["algorithmic_flops"] - plan.flops) < 1 and
                abs(e["algorithmic_bytes"] - plan.bytes) < 1):
            return e["dram_bytes"]
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def get_plan(wl, net, name, plan_seconds, rank=0, size=1, log=None):
    """Committed plan if present, else a time-bounded search voted across ranks."""
    from qibotn_b200 import dist, plans
    from qibotn_b200.planner import find_path

    info = plans.load_plan(name, net.modes, net.out)
    if info is not None:
        return info, "plans/%s.json" % name
    info = find_path(net.modes, net.out, samples=8, seed=rank, target_log2_width=31,
                     time_budget_s=plan_seconds)
    if size > 1:
        info, _ = dist.vote_best(info.opt_cost, info)
    return info, "searched %.0fs" % plan_seconds


# ------------------------------------------------------------------ CPU baseline / reference arm
def cpu_plan(wl, net, name, seconds):
    from qibotn_b200 import plans
    from qibotn_b200.planner import find_path

    info = plans.load_plan(name + "_cpu", net.modes, net.out)
    if info is None:
        info = find_path(net.modes, net.out, samples=4, seed=0, target_log2_width=24,
                         time_budget_s=seconds)
    return info


def use_all_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=len(os.sched_getaffinity(0)))
    except Exception:            # threadpoolctl missing: keep whatever the BLAS picked
        pass
    return len(os.sched_getaffinity(0))


def cpu_sample(wl, net, info, seconds, max_slices=64):
    """Run slices of ``info`` with numpy.tensordot for about ``seconds``; returns
    (TFLOP/s, slices done, wall seconds, flops per slice)."""
    from oracle import contract as oc
    from qibotn_b200 import executor

    low = executor.lower(net.modes, net.out, info, wl.dtype)       # accounting only
    dt = np.complex64 if wl.dtype == "complex64" else np.complex128
    ops = [np.asarray(t, dtype=dt) for t in net.tensors]
    per_slice = low.flops_per_slice if info.num_slices > 1 else low.flops_hoisted
    t0 = time.perf_counter()
    done = 0
    while done < min(info.num_slices, max_slices):
        oc.contract_ssa(ops, net.modes, net.out, info.path, sliced=info.sliced_labels,
                        slice_id=done, dtype=dt)
        done += 1
        if time.perf_counter() - t0 > seconds:
            break
    wall = time.perf_counter() - t0
    flops = per_slice * done + (low.flops_hoisted if info.num_slices > 1 else 0) * done
    return flops / wall / 1e12, done, wall, per_slice


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from qibotn_b200 import workloads

    wl = workloads.build(a.workload)
    net = wl.network()
    info = cpu_plan(wl, net, a.workload, a.plan_seconds)
    cores = use_all_cores()
    per_step = max(2.0, min(30.0, 150.0 / max(1, a.steps + a.warmup)))
    for _ in range(min(a.warmup, 1)):
        cpu_sample(wl, net, info, 1.0, max_slices=1)
    rates, walls, slices = [], [], []
    for _ in range(a.steps):
        r, done, wall, per_slice = cpu_sample(wl, net, info, per_step)
        rates.append(r * wall); walls.append(wall); slices.append(done)
    value = sum(rates) / sum(walls)
    full_s = info.flops / (value * 1e12)
    sample = (f"{sum(slices)} of {info.num_slices} slices of a width-2^{info.log2_width} plan "
              f"(10^{math.log10(info.flops):.2f} flops in all), numpy.tensordot per tree node")
    line = {
        "impl": "reference", "metric": "tn_contraction_tflops", "value": value, "unit": "TFLOP/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * sum(walls) / a.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": {"workload": wl.description, "plan": "cpu plan, width 2^%d" % info.log2_width},
        "full_contraction_s_extrapolated": full_s,
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm
def measure_fp32_peak(torch):
    """Live FP32 FMA rate of this GPU through the library's own complex-FMA tile kernel is
    not a peak; use a dense torch.matmul fp32 (no TF32) as the pipe reference instead."""
    torch.backends.cuda.matmul.allow_tf32 = False
    n = 8192
    x = torch.randn(n, n, device="cuda", dtype=torch.float32)
    y = torch.randn(n, n, device="cuda", dtype=torch.float32)
    for _ in range(2):
        x @ y
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        x @ y
    e1.record()
    torch.cuda.synchronize()
    return 3 * 2 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12


def mps_metric(n=64, steps=20, chi=256):
    """The second half of BASELINE.json's metric: MPS gates/s on config 3 (64-site TFIM Trotter
    brickwork, 20 steps, bond <= 256, QR off, SVD abs_cutoff 1e-12, complex128), one GPU."""
    import torch

    from qibotn_b200 import circuits, mps as gmps

    algo = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12, "max_extent": chi}}
    gmps.QiboCircuitToMPS(circuits.tfim_trotter(n, 2), algo, dtype="complex128")          # warm-up
    circ = circuits.tfim_trotter(n, steps)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    conv = gmps.QiboCircuitToMPS(circ, algo, dtype="complex128")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gates = len(circ.queue)
    return {"workload": f"{n}-site TFIM Trotter, {steps} steps, bond<={chi}, complex128", "gates": gates,
            "seconds": dt, "gates_per_s": gates / dt, "max_bond": conv.stats["max_bond"],
            "two_site_updates": conv.stats["two_site"], "svd_batches": conv.stats["svd_batches"],
            "svd_sweeps": conv.stats["svd_sweeps"], "svd_seconds": conv.stats.get("svd_seconds"),
            "svd_credited_tflops": conv.stats.get("svd_credited_flops", 0.0) / max(conv.stats.get("svd_seconds", 0.0), 1e-9) / 1e12,
            "svd_note": "credited flops = 4(14 p q^2 + 8 q^3) per p x q block (SURVEY 8d); LQ-preconditioned one-sided "
                        "Jacobi, FP64 FMA pipe (rotation-bound, not GEMM-shaped)"}


def run_b200(a):
    import ctypes as C

    import torch

    from qibotn_b200 import dist, executor, lib, workloads

    rank, size, device = dist.initialize()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    wl = workloads.build(a.workload)
    net = wl.network()
    info, plan_src = get_plan(wl, net, a.workload, a.plan_seconds, rank, size)
    t0 = time.perf_counter()
    low = executor.lower(net.modes, net.out, info, wl.dtype)
    lower_s = time.perf_counter() - t0
    runner = executor.ContractionRunner(low, use_graph=True)
    h = lib.load()

    batch = min(info.num_slices, a.slices_per_step) if info.num_slices > 1 else 1
    mine = list(dist.compute_slices(batch, rank, size))
    flops_step = low.flops_hoisted + batch * low.flops_per_slice
    bytes_step = low.bytes_hoisted + batch * low.bytes_per_slice
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step(e2e):
        if e2e:
            runner.upload(net.tensors)
        out = runner.run(mine)
        if size > 1:
            out = dist.reduce_result(out, method="NCCL")
        if e2e:
            return out.cpu()
        return out

    def timed(n, e2e):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        dist.barrier(); torch.cuda.synchronize()
        wall0 = time.perf_counter()
        ev[0].record()
        for _ in range(n):
            res = step(e2e)
        ev[1].record()
        torch.cuda.synchronize(); dist.barrier()
        wall = time.perf_counter() - wall0
        ms = ev[0].elapsed_time(ev[1])
        if e2e:
            ms = max(ms, wall * 1e3)            # the D2H read ends on the host
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if size > 1:
            import torch.distributed as td
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item()), res

    runner.upload(net.tensors)
    torch.cuda.synchronize()
    for _ in range(a.warmup):
        step(False)
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    ms_dev, res = timed(a.steps, False)
    launches = runner.launches * a.steps
    ms_e2e, res_e2e = timed(a.steps, True)
    clocks = sampler.stop() if rank == 0 else None
    amp = complex(res.reshape(-1)[0].item())

    if rank != 0:
        return
    peaks, peak_src = load_peaks()
    tfs = flops_step * a.steps / (ms_dev * 1e-3) / 1e12
    tfs_e2e = flops_step * a.steps / (ms_e2e * 1e-3) / 1e12
    slice_s = (ms_dev * 1e-3 / a.steps) / max(1, len(mine))
    full_s = slice_s * info.num_slices / size

    # dominant kernel: the node with the most work per step, timed alone (CUDA events on the
    # launch stream, L2 flushed between launches)
    node = max(low.nodes, key=lambda n: n.plan.flops * (batch if n.depends else 1))
    st = torch.cuda.current_stream().cuda_stream
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.check(h.qtn_pair_run(C.byref(node.plan), runner._ptr(node.a), runner._ptr(node.b),
                                 runner._ptr(node.c), runner.ws.data_ptr(),
                                 runner.slice_id.data_ptr(), st))
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    k_ms = float(np.mean(ts[1:]))
    k_tfs = node.plan.flops / (k_ms * 1e-3) / 1e12
    intensity = node.plan.flops / node.plan.bytes
    names = {0: "gett_tile_kernel<float>", 1: "gett_reduce_kernel<float>", 2: "gett_tc_persistent_kernel",
             3: "gett_kred_kernel<float>"}
    if node.plan.kernel == lib.KERNEL_TC and node.plan.tc_pair:
        names[2] = "gett_tc_pair_kernel"
    if node.plan.kernel == lib.KERNEL_TC:
        # tcgen05 kind::tf32, 3 MMAs per product (hi*hi + hi*lo + lo*hi): the tensor pipe does 3x
        # the algorithmic flops at the tf32 rate (half the bf16 rate measured in MEASURED_PEAKS.json)
        roof = {"bound": "tensor", "achieved": k_tfs, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s"}
        extra = {"mma": "tcgen05.mma kind::tf32, 3xTF32 split, complex as 4 real GEMMs",
                 "tensor_flops_per_algorithmic_flop": 3.0,
                 "tf32_peak_assumed_tflops": peaks["bf16_tflops"] / 2,
                 "tensor_pipe_frac_tf32": 3.0 * k_tfs / (peaks["bf16_tflops"] / 2),
                 "hbm_gbs_at_this_speed": node.plan.bytes / (k_ms * 1e-3) / 1e9,
                 "traffic_note": "traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch from an ncu "
                                 "--set full capture of this kernel on this node's (m, n, k) shape, launched "
                                 "stand-alone (profiles/ncu_traffic.json); null when that shape was not captured",
                 "note": "frac is algorithmic complex64 flops over the measured bf16 cuBLAS burst; "
                         "tensor_pipe_frac_tf32 is the share of the tf32 pipe the 3x split keeps busy"}
    else:
        fp32_peak = measure_fp32_peak(torch)
        hbm_bound = intensity < fp32_peak * 1e12 / (peaks["hbm_gbs"] * 1e9)
        if hbm_bound:
            roof = {"bound": "hbm", "achieved": node.plan.bytes / (k_ms * 1e-3) / 1e9,
                    "peak": peaks["hbm_gbs"], "unit": "GB/s"}
        else:
            roof = {"bound": "tensor", "achieved": k_tfs, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s"}
        extra = {"fp32_pipe_tflops_measured": fp32_peak, "fp32_pipe_frac": k_tfs / fp32_peak,
                 "note": "complex64 FFMA kernel on the FP32 pipe; tensor peak is the bf16 cuBLAS burst"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof.update({"traffic": ncu_traffic(names.get(node.plan.kernel, "?"), node.plan), "peak_source": peak_src + " (MEASURED_PEAKS.json, burst)",
                 "kernel": names.get(node.plan.kernel, "?"),
                 "kernel_ms": k_ms, "kernel_flops": node.plan.flops, "kernel_bytes": node.plan.bytes,
                 "tile_log2": [node.plan.tmb, node.plan.tnb, node.plan.tkb],
                 "share_of_step_flops": node.plan.flops * (batch if node.depends else 1) / flops_step})
    roof.update(extra)
    kinds = {}
    for nd in low.nodes:
        if nd.depends or info.num_slices == 1:
            kinds[names.get(nd.plan.kernel, "?")] = kinds.get(names.get(nd.plan.kernel, "?"), 0.0) + nd.plan.flops
    roof["flops_share_by_kernel"] = {k: v / sum(kinds.values()) for k, v in kinds.items()}

    line = {
        "metric": "tn_contraction_tflops", "value": tfs, "unit": "TFLOP/s", "n_gpus": size,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_dev / a.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": wl.dtype,
        "data": "synthetic",
        "config": {"workload": wl.description, "slices_per_step": batch,
                   "num_slices": info.num_slices, "log2_width": info.log2_width,
                   "plan_log10_flops": math.log10(info.flops), "plan": plan_src,
                   "flops_per_step": flops_step, "algorithmic_bytes_per_step": bytes_step,
                   "l2": "intermediates up to 2^%d elements (>> 126 MB L2)" % info.log2_width,
                   "parallelism": "slices over %d GPU(s), one NCCL all-reduce" % size},
        "full_contraction_s_extrapolated": full_s,
        "amplitude_partial": [amp.real, amp.imag],
        "e2e": {"value": tfs_e2e, "unit": "TFLOP/s", "ms_per_step": ms_e2e / a.steps,
                "h2d_bytes_per_step": low.leaf_bytes, "d2h_bytes_per_step": 8 if wl.dtype == "complex64" else 16},
        "gpu_launches": launches, "lower_s": lower_s, "clocks": clocks, "roofline": roof,
    }
    if size == 1 and not a.no_mps:
        line["mps"] = mps_metric()
    if size == 1 and not a.no_cpu_baseline:
        cinfo = cpu_plan(wl, net, a.workload, a.plan_seconds)
        use_all_cores()
        r, done, wall, _ = cpu_sample(wl, net, cinfo, a.cpu_seconds)
        line["cpu_baseline"] = {
            "value": r, "unit": "TFLOP/s", "cores": len(os.sched_getaffinity(0)), "kind": "port",
            "sample": f"{done} of {cinfo.num_slices} slices of a width-2^{cinfo.log2_width} plan "
                      f"(10^{math.log10(cinfo.flops):.2f} flops), numpy.tensordot per node, {wall:.1f} s",
            "full_contraction_s_extrapolated": cinfo.flops / (r * 1e12)}
    print(json.dumps(line), flush=True)


def _shutdown():
    try:
        import torch.distributed as td
        if td.is_available() and td.is_initialized():
            td.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    args = parse()
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)
    finally:
        _shutdown()
