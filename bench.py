#!/usr/bin/env python
"""Headline benchmark: the sliced single-amplitude contraction of the 53-qubit depth-14 Sycamore-like
random circuit (BASELINE.json configs[3], the config the metric is quoted on; it fits one GPU), plus the
MPS gates/s half of the metric (configs[2]) on one GPU.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--slices-per-step S]

One *step* = ONE WHOLE AMPLITUDE: all 64 slices of the committed plan (``plans/c4.json``), split over the
N ranks as the reference's ``compute_slices`` does (eval.py:197-205), each rank running its slices
through the fused permute+contract kernels, then the single NCCL all-reduce of the partials
(``--slices-per-step`` < 64 restricts the device-resident arm to the first S slices for quick runs).

* ``value``  -- algorithmic TFLOP/s (8 flops per complex MAC) of the whole job with the leaf
  tensors resident in HBM, timed with CUDA events, max over ranks.
* ``e2e``    -- the same amplitude through the reference-facing call a user makes,
  ``qibotn_b200.eval.amplitude_tn(circuit, dtype, bitstring)`` (what ``CuTensorNet.execute_circuit`` runs
  underneath, backends/cutensornet.py:71-169): every step converts the circuit's gates to tensors on the host,
  looks the plan up, uploads the packed leaves from pinned memory, contracts every slice and reads the amplitude
  back.  The FIRST call additionally loads the committed plan, lowers it (64-node kernel plans + arena),
  allocates the arenas and captures the CUDA graphs: reported apart as ``first_call_s``.
* ``roofline`` -- the node of the contraction tree that takes the most device time, timed alone with CUDA
  events (L2 flushed in between): HBM-bound or tensor-bound by its arithmetic intensity against the measured
  peaks; ``nodes`` lists the largest nodes with their own bound and fraction.
* ``cpu_baseline`` -- the CPU restatement of the reference's quimb path (numpy.tensordot per tree node) on a
  bounded sample of slices of a CPU-sized plan, all host cores.
* ``mps`` -- config 3 (64-site TFIM Trotter, bond <= 256, complex128): gates/s over the whole circuit and over
  the saturated layers, the CPU oracle's rate on the same saturated state, cuSOLVER / cuBLAS stand-ins
  (``torch.linalg.svd``, ``torch.matmul``) on the same batch shapes, and the SVD's credited FLOP rate
  against the measured FP64 FMA-pipe peak.

``--impl reference`` times that CPU arm alone (rank 0 only); it imports nothing that maps the CUDA library.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
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
    ap.add_argument("--slices-per-step", type=int, default=0, help="0 = every slice of the plan (one whole amplitude)")
    ap.add_argument("--no-mps", action="store_true", help="skip the MPS gates/s measurement (N=1 only)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end arm (quick kernel runs)")
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


def ncu_traffic(kernel, flops, nbytes):
    """DRAM bytes per launch of this node, if an ncu capture of it is on file (profiles/ncu_traffic.json:
    entries keyed by kernel name and the node's algorithmic flops and bytes)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            table = json.load(f)
    except (OSError, ValueError):
        return None
    for e in table.get("entries", []):
        if (e["kernel"] == kernel and abs(e["algorithmic_flops"] - flops) < 1 and
                abs(e["algorithmic_bytes"] - nbytes) < 1):
            return e["dram_bytes"]
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------ CPU baseline / reference arm
# (nothing below imports qibotn_b200.lib / executor: the reference arm must not map the CUDA library)
def cpu_plan(net, name, seconds):
    from qibotn_b200 import plans
    from qibotn_b200.planner import find_path

    info = plans.load_plan(name + "_cpu", net.modes, net.out)
    if info is None:
        info = find_path(net.modes, net.out, samples=4, seed=0, target_log2_width=24, time_budget_s=seconds)
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
    (TFLOP/s, slices done, wall seconds).  Flops are the planner's count for the whole plan
    (8 per complex MAC, every tree node, every slice) divided by the number of slices."""
    from oracle import contract as oc

    dt = np.complex64 if wl.dtype == "complex64" else np.complex128
    ops = [np.asarray(t, dtype=dt) for t in net.tensors]
    per_slice = info.flops / info.num_slices
    t0 = time.perf_counter()
    done = 0
    while done < min(info.num_slices, max_slices):
        oc.contract_ssa(ops, net.modes, net.out, info.path, sliced=info.sliced_labels, slice_id=done, dtype=dt)
        done += 1
        if time.perf_counter() - t0 > seconds:
            break
    wall = time.perf_counter() - t0
    return per_slice * done / wall / 1e12, done, wall


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from qibotn_b200 import workloads

    wl = workloads.build(a.workload)
    net = wl.network()
    info = cpu_plan(net, a.workload, a.plan_seconds)
    cores = use_all_cores()
    per_step = max(2.0, min(30.0, 150.0 / max(1, a.steps + a.warmup)))
    for _ in range(min(a.warmup, 1)):
        cpu_sample(wl, net, info, 1.0, max_slices=1)
    work, walls, slices = [], [], []
    for _ in range(a.steps):
        r, done, wall = cpu_sample(wl, net, info, per_step)
        work.append(r * wall); walls.append(wall); slices.append(done)
    value = sum(work) / sum(walls)
    sample = (f"{sum(slices)} of {info.num_slices} slices of a width-2^{info.log2_width} plan "
              f"(10^{math.log10(info.flops):.2f} flops in all), numpy.tensordot per tree node")
    line = {
        "impl": "reference", "metric": "tn_contraction_tflops", "value": value, "unit": "TFLOP/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * sum(walls) / a.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": {"workload": wl.description, "plan": "cpu plan, width 2^%d" % info.log2_width,
                   "same_contraction_tree_as_gpu_arm": False,
                   "note": "a 2^31-wide slice (16 GiB intermediates) does not run on the host in bounded time; the "
                           "CPU arm contracts the same network with a 2^24-wide plan"},
        "full_contraction_s_extrapolated": info.flops / (value * 1e12),
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ our arm
def _event_ms(torch, fn, reps=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measure_fp16_tensor_peak(torch):
    """cuBLAS fp16 GEMM 8192^3, best of 5: the pipe the 3xFP16 split runs on (same rate as bf16)."""
    n = 8192
    x = torch.randn(n, n, device="cuda", dtype=torch.float16)
    y = torch.randn(n, n, device="cuda", dtype=torch.float16)
    for _ in range(2):
        x @ y
    best = min(_event_ms(torch, lambda: x @ y) for _ in range(5))
    return 2 * n ** 3 / (best * 1e-3) / 1e12


def measure_fma_peak(torch, lib, h, is_double):
    import ctypes as C

    scratch = torch.empty(148 * 8 * 256 * 8 * 2, dtype=torch.uint8, device="cuda")
    out = C.c_double(0.0)
    lib.check(h.qtn_microbench_fma(int(is_double), 4096, scratch.data_ptr(), scratch.numel(), C.byref(out),
                                   torch.cuda.current_stream().cuda_stream))
    return out.value


def mps_metric(cpu_seconds, n=64, steps=20, chi=256):
    """The second half of BASELINE.json's metric: MPS gates/s on config 3 (64-site TFIM Trotter brickwork,
    20 steps, bond <= 256, SVD abs_cutoff 1e-12, complex128), one GPU."""
    import torch

    from oracle import mps as omps
    from qibotn_b200 import circuits, lib, mps as gmps
    from qibotn_b200.network import QiboCircuitToEinsum

    h = lib.load()
    algo = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12, "max_extent": chi}}
    gmps.QiboCircuitToMPS(circuits.tfim_c3(n, 2), algo, dtype="complex128")          # warm-up
    circ = circuits.tfim_c3(n, steps)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    conv = gmps.QiboCircuitToMPS(circ, algo, dtype="complex128")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gates = len(circ.queue)
    helper = gmps.MPSContractionHelper(n)
    norm = float(helper.contract_norm(conv.mps_tensors).cpu())
    bonds = [int(t.shape[2]) for t in conv.mps_tensors[:-1]]
    svd_s = conv.stats.get("svd_seconds", 0.0)
    credited = conv.stats.get("svd_credited_flops", 0.0)
    out = {"workload": f"{n}-site TFIM Trotter (J dt, h dt) = ({circuits.C3_ANGLES['j_dt']}, {circuits.C3_ANGLES['h_dt']}), "
                       f"{steps} steps, bond<={chi}, abs_cutoff 1e-12, complex128",
           "gates": gates, "seconds": dt, "gates_per_s": gates / dt, "max_bond": max(bonds),
           "bonds_at_max": sum(b == chi for b in bonds), "norm": norm, "discarded_weight": 1.0 - norm,
           "two_site_updates": conv.stats["two_site"], "svd_batches": conv.stats["svd_batches"],
           "svd_sweeps": conv.stats["svd_sweeps"], "svd_seconds": svd_s,
           "svd_credited_tflops": credited / max(svd_s, 1e-9) / 1e12}
    # --- saturated regime: two more Trotter steps on the final (bond-256) state
    layer = QiboCircuitToEinsum(circuits.tfim_c3(n, 2), dtype="complex128").gate_tensors[n:]   # without the H layer
    eng = conv.engine
    host_sites = [t.cpu().numpy() for t in eng.sites]
    ev0 = len(eng._svd_events)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for g, q in layer:
        eng.apply_gate(g, q)
    eng.flush()
    torch.cuda.synchronize()
    sat = time.perf_counter() - t0
    sat_svd = sum(x.elapsed_time(y) for x, y in eng._svd_events[ev0:]) * 1e-3
    two_site = sum(1 for _, q in layer if len(q) == 2)
    sat_credited = two_site * 4.0 * (14.0 * 512 * 512 ** 2 + 8.0 * 512 ** 3)       # upper bound: every block 512 x 512
    out["saturated"] = {"gates": len(layer), "seconds": sat, "gates_per_s": len(layer) / sat, "svd_seconds": sat_svd,
                        "note": "two further Trotter steps on the final state (all bulk bonds at 256)"}
    # --- CPU oracle (numpy.linalg.svd gate loop, the restatement of mps_utils.py:42-95) on the SAME saturated state
    cores = use_all_cores()
    mps_cpu = [s.copy() for s in host_sites]
    t0 = time.perf_counter()
    done = 0
    for g, q in layer:
        omps.apply_gate(mps_cpu, np.asarray(g), tuple(q), algo)
        done += 1
        if time.perf_counter() - t0 > cpu_seconds:
            break
    cpu_wall = time.perf_counter() - t0
    out["cpu_baseline"] = {"value": done / cpu_wall, "unit": "gates/s", "cores": cores, "kind": "port",
                           "sample": f"first {done} gates of one Trotter step applied to the saturated (bond-256) state, "
                                     f"numpy einsum + numpy.linalg.svd per gate, {cpu_wall:.1f} s",
                           "gpu_over_cpu_saturated": (len(layer) / sat) / (done / cpu_wall)}
    # --- library stand-ins on the shapes of one saturated layer (31 blocks of 512 x 512 complex128)
    x = torch.randn(31, 512, 512, dtype=torch.complex128, device="cuda")
    a_ = torch.randn(31, 512, 256, dtype=torch.complex128, device="cuda")
    b_ = torch.randn(31, 256, 512, dtype=torch.complex128, device="cuda")
    torch.linalg.svd(x[:2], full_matrices=False)
    svd_ms = _event_ms(torch, lambda: torch.linalg.svd(x, full_matrices=False))
    torch.bmm(a_, b_)
    mm_ms = _event_ms(torch, lambda: torch.bmm(a_, b_), reps=3)
    layer_svd_ms = 1e3 * sat_svd / max(1, len(eng._svd_events) - ev0)
    out["library_stand_ins"] = {
        "torch_linalg_svd_31x512x512_c128_ms": svd_ms, "ours_svd_batch_ms_saturated": layer_svd_ms,
        "torch_bmm_31x(512x256x512)_c128_ms": mm_ms,
        "note": "cuSOLVER gesvd(j) / cuBLAS zgemm through torch, the libraries the reference reaches through "
                "cuTensorNet's contract_decompose (mps_utils.py:78-84); same batch shape as one saturated layer"}
    # --- roofline of the decomposition: credited SVD flops (SURVEY 8d convention) against the measured FP64 FMA pipe
    fp64_peak = measure_fma_peak(torch, lib, h, True)
    out["roofline"] = {"bound": "fp64_fma", "achieved": credited / max(svd_s, 1e-9) / 1e12, "peak": fp64_peak,
                       "unit": "TFLOP/s", "frac": credited / max(svd_s, 1e-9) / 1e12 / fp64_peak,
                       "achieved_saturated": sat_credited / max(sat_svd, 1e-9) / 1e12,
                       "peak_source": "qtn_microbench_fma(double): 16 independent DFMA chains per thread, this run",
                       "note": "credited flops = 4(14 p q^2 + 8 q^3) per p x q block; the Jacobi sweeps execute more "
                               "(rotations of every row pair per sweep), which is not credited"}
    return out


def run_b200(a):
    import ctypes as C

    import torch

    from qibotn_b200 import dist, engine, executor, lib, plans, workloads
    from qibotn_b200 import eval as ev

    rank, size, device = dist.initialize()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    wl = workloads.build(a.workload)
    t0 = time.perf_counter()
    net = wl.network()
    build_s = time.perf_counter() - t0
    info = plans.load_plan(a.workload, net.modes, net.out)
    plan_src = "plans/%s.json" % a.workload
    if info is None:
        from qibotn_b200.planner import find_path
        info = find_path(net.modes, net.out, samples=8, seed=rank, target_log2_width=31, time_budget_s=a.plan_seconds)
        if size > 1:
            info, _ = dist.vote_best(info.opt_cost, info)
        plan_src = "searched %.0fs" % a.plan_seconds
    t0 = time.perf_counter()
    low = executor.lower(net.modes, net.out, info, wl.dtype)
    lower_s = time.perf_counter() - t0
    runner = executor.ContractionRunner(low, use_graph=True)
    h = lib.load()

    batch = info.num_slices if a.slices_per_step <= 0 else min(info.num_slices, a.slices_per_step)
    mine = list(dist.compute_slices(batch, rank, size))
    flops_step = low.flops_hoisted + batch * low.flops_per_slice
    bytes_step = low.bytes_hoisted + batch * low.bytes_per_slice
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        out = runner.run(mine)
        if size > 1:
            out = dist.reduce_result(out, method="NCCL")
        return out

    def timed(n, fn, host_end=False):
        """K steps bracketed by barrier + synchronize; device time by CUDA events, MAX over ranks."""
        ev_ = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        dist.barrier(); torch.cuda.synchronize()
        wall0 = time.perf_counter()
        ev_[0].record()
        for _ in range(n):
            res = fn()
        ev_[1].record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        dist.barrier()
        mine_ms = ev_[0].elapsed_time(ev_[1])
        if host_end:
            mine_ms = max(mine_ms, wall * 1e3)            # the result is read on the host: the host clock closes the step
        t = torch.tensor([mine_ms], dtype=torch.float64, device="cuda")
        every = [mine_ms]
        if size > 1:
            import torch.distributed as td
            gathered = [torch.zeros_like(t) for _ in range(size)]
            td.all_gather(gathered, t)
            every = [float(x.item()) for x in gathered]
        return max(every), every, res

    runner.upload(net.tensors)
    torch.cuda.synchronize()
    for _ in range(a.warmup):
        step()
    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    ms_dev, ms_ranks, res = timed(a.steps, step)
    launches = runner.launches * a.steps
    clocks = sampler.stop() if rank == 0 else None
    amp = complex(res.reshape(-1)[0].item())
    hoisted_ms = _event_ms(torch, runner._run_hoisted)

    # ---- per-node device times of one slice (rank 0): the dominant node and the table around it
    roof = None
    if rank == 0:
        peaks, peak_src = load_peaks()
        fp16_peak = measure_fp16_tensor_peak(torch)
        # algorithmic flops the tensor pipe can deliver: 3 fp16 MMA flops per flop of the 3xFP16 split
        ridge = (peaks["bf16_tflops"] / 3.0) * 1e12 / (peaks["hbm_gbs"] * 1e9)
        rows = []
        names = {0: "gett_tile_kernel<float>", 1: "gett_reduce_kernel<float>", 2: "gett_tc_persistent_kernel",
                 3: "gett_kred_kernel<float>"}
        for node, ms in runner.time_nodes(0, reps=4, flush=flush):
            P = node.plan
            kname = names.get(P.kernel, "?")
            if P.kernel == lib.KERNEL_TC and P.tc_pair:
                kname = "gett_tc_pair2_kernel" if P.tc_pair == 2 else "gett_tc_pair_kernel"
            ai = P.flops / P.bytes
            hbm_bound = ai < ridge
            ra, rb, rc = (len(low.tensors[t].labels) for t in (node.a, node.b, node.c))
            k = (ra + rb - rc) // 2
            row = {"m": max(ra, rb) - k, "n": min(ra, rb) - k, "k": k, "kernel": kname, "split": "3xFP16" if P.tc_kind else
                   ("3xTF32" if P.kernel == lib.KERNEL_TC else "none"), "tile_log2": [P.tmb, P.tnb, P.tkb], "ms": ms,
                   "flops": P.flops, "bytes": P.bytes, "intensity": ai, "TFs": P.flops / ms / 1e9, "GBs": P.bytes / ms / 1e6,
                   "bound": "hbm" if hbm_bound else "tensor",
                   "frac": (P.bytes / ms / 1e6 / peaks["hbm_gbs"]) if hbm_bound else (P.flops / ms / 1e9 / peaks["bf16_tflops"])}
            rows.append(row)
        rows.sort(key=lambda r: -r["ms"])
        top = rows[0]
        slice_ms = sum(r["ms"] for r in rows)
        if top["bound"] == "hbm":
            roof = {"bound": "hbm", "achieved": top["GBs"], "peak": peaks["hbm_gbs"], "unit": "GB/s"}
        else:
            roof = {"bound": "tensor", "achieved": top["TFs"], "peak": peaks["bf16_tflops"], "unit": "TFLOP/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof.update({
            "traffic": ncu_traffic(top["kernel"], top["flops"], top["bytes"]), "peak_source": peak_src + ", burst",
            "kernel": top["kernel"], "kernel_ms": top["ms"], "kernel_flops": top["flops"], "kernel_bytes": top["bytes"],
            "node": {k_: top[k_] for k_ in ("m", "n", "k", "tile_log2", "split", "intensity")},
            "share_of_slice_time": top["ms"] / slice_ms,
            "ridge_flop_per_byte": ridge,
            "mma": "tcgen05.mma cta_group::2 kind::f16, scaled 3xFP16 split (12 MMAs per 16 contracted complex values), "
                   "complex product as 2 concatenated real GEMMs per half tile",
            "tensor_flops_per_algorithmic_flop": 3.0,
            "fp16_tensor_peak_measured_tflops": fp16_peak,
            "tensor_pipe_frac_of_measured_fp16": 3.0 * top["TFs"] / fp16_peak,
            "sum_of_node_ms_per_slice": slice_ms,
            "nodes": [{k_: (round(v, 4) if isinstance(v, float) else v) for k_, v in r.items() if k_ not in ("flops", "bytes")}
                      for r in rows[:10]],
            "note": "a node is HBM-bound when its algorithmic flop/byte is below the ridge (measured bf16 burst / 3) / "
                    "measured HBM copy rate; frac is against MEASURED_PEAKS.json for the bound that applies; traffic = "
                    "dram__bytes_read + dram__bytes_write per launch from an ncu --set full capture of this node shape "
                    "(profiles/ncu_traffic.json), null when that shape was not captured"})
    del flush

    # ---- end to end through the public call, host buffers in, amplitude out
    e2e = None
    if not a.no_e2e:
        runner = None
        torch.cuda.empty_cache()
        distributed = size > 1

        def e2e_step():
            out = ev.amplitude_tn(wl.circuit, wl.dtype, wl.bitstring, n_samples=8, distributed=distributed)
            arr = out[0] if distributed else out
            return arr.get()

        engine.clear_caches()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        first = e2e_step()
        torch.cuda.synchronize()
        first_s = time.perf_counter() - t0
        for _ in range(min(a.warmup, 2)):
            e2e_step()
        ms_e2e, ms_e2e_ranks, res_e2e = timed(a.steps, e2e_step, host_end=True)
        low2 = next(iter(engine._PLAN_CACHE.values()))[1]
        flops_e2e = low2.flops_hoisted + info.num_slices * low2.flops_per_slice
        e2e = {"value": flops_e2e * a.steps / (ms_e2e * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": ms_e2e / a.steps,
               "h2d_bytes_per_step": int(low2.leaf_bytes), "d2h_bytes_per_step": 8 if wl.dtype == "complex64" else 16,
               "call": "qibotn_b200.eval.amplitude_tn(circuit, dtype, bitstring%s)" % (", distributed=True" if distributed else ""),
               "slices_per_step": info.num_slices, "first_call_s": first_s,
               "first_call_includes": "circuit -> tensors, plan lookup (plans/%s.json by structure), lowering, arena "
                                      "allocation, CUDA-graph capture, contraction" % a.workload,
               "amplitude": [float(np.real(res_e2e.reshape(-1)[0])), float(np.imag(res_e2e.reshape(-1)[0]))],
               "ms_per_rank": [m / a.steps for m in ms_e2e_ranks]}

    if rank != 0:
        return
    tfs = flops_step * a.steps / (ms_dev * 1e-3) / 1e12
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
        "contraction_wall_s": (ms_dev * 1e-3 / a.steps) * info.num_slices / batch,
        "contraction_wall_s_is": "measured (one step = the whole amplitude)" if batch == info.num_slices else "extrapolated from %d slices" % batch,
        "amplitude": [amp.real, amp.imag],
        "ms_per_rank": [m / a.steps for m in ms_ranks], "hoisted_ms": hoisted_ms,
        "host_s": {"circuit_to_network": build_s, "lowering": lower_s},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if size == 1 and not a.no_mps:
        line["mps"] = mps_metric(a.cpu_seconds)
    if size == 1 and not a.no_cpu_baseline:
        cinfo = cpu_plan(net, a.workload, a.plan_seconds)
        use_all_cores()
        r, done, wall = cpu_sample(wl, net, cinfo, a.cpu_seconds)
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
