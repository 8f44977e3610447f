"""BASELINE config 2 on one GPU: 30-qubit random brickwork circuit, depth 12, <Z...Z> as ONE closed sandwich
contraction (complex64).  Plans on the host (time-bounded search, or plans/c2.json when committed), lowers, runs,
and prints the wall time, the algorithmic GB/s and TFLOP/s, and where the time goes per kernel class
(per-node CUDA-event times, L2 flushed between launches).

    python scripts/c2_bench.py [--plan-seconds 60] [--save-plan]"""
import argparse
import json
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

from qibotn_b200 import executor, lib, plans, workloads  # noqa: E402
from qibotn_b200.planner import find_path  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--plan-seconds", type=float, default=60.0)
ap.add_argument("--save-plan", action="store_true")
ap.add_argument("--dtype", default=None)
a = ap.parse_args()

wl = workloads.build(a.workload)
dtype = a.dtype or wl.dtype
net = wl.network()
t0 = time.perf_counter()
info = plans.load_plan(a.workload, net.modes, net.out)
src = f"plans/{a.workload}.json"
if info is None:
    info = find_path(net.modes, net.out, samples=16, seed=0, target_log2_width=31 if dtype == "complex64" else 30,
                     time_budget_s=a.plan_seconds)
    src = f"searched {a.plan_seconds:.0f}s"
    if a.save_plan:
        plans.save_plan(a.workload, info, net.modes, net.out)
plan_s = time.perf_counter() - t0
low = executor.lower(net.modes, net.out, info, dtype)
runner = executor.ContractionRunner(low, use_graph=True)
runner.upload(net.tensors)
torch.cuda.synchronize()
for _ in range(2):
    out = runner.run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 5
for _ in range(reps):
    out = runner.run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = low.flops_hoisted + info.num_slices * low.flops_per_slice
nbytes = low.bytes_hoisted + info.num_slices * low.bytes_per_slice
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
KN = {0: "gett_tile_kernel", 1: "gett_reduce_kernel", 2: "gett_tc", 3: "gett_kred_kernel"}
cls = {}
rows = []
if info.num_slices == 1:
    # unsliced: every node is "hoisted"; time them in order
    st = torch.cuda.current_stream().cuda_stream
    import ctypes as C
    runner._measure_leaves()
    if runner._amax_used["hoist"]:
        runner._amax["hoist"].zero_()
    slot = runner._amax_slot
    for node in low.nodes:
        ts = []
        for _ in range(3):
            flush.zero_()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            lib.check(runner.lib.qtn_pair_run_scaled(C.byref(node.plan), runner._ptr(node.a), runner._ptr(node.b),
                                                     runner._ptr(node.c), runner.ws.data_ptr(), runner.slice_id.data_ptr(),
                                                     slot.get(node.a), slot.get(node.b), slot.get(node.c), st))
            a1.record(); torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
        rows.append((node, sum(ts[1:]) / 2))
else:
    rows = runner.time_nodes(0, reps=3, flush=flush)
for node, t in rows:
    P = node.plan
    name = KN[P.kernel] + (f"_pair{P.tc_pair}" if P.kernel == 2 else "")
    c = cls.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0.0, "flops": 0.0})
    c["launches"] += 1; c["ms"] += t; c["bytes"] += P.bytes; c["flops"] += P.flops
top = sorted(rows, key=lambda r: -r[1])[:8]
print(json.dumps({
    "workload": wl.description, "dtype": dtype, "plan": src, "plan_s": plan_s, "num_slices": info.num_slices,
    "log2_width": info.log2_width, "log10_flops": __import__("math").log10(info.flops), "ms": ms,
    "TFs": flops / ms / 1e9, "GBs": nbytes / ms / 1e6, "value": [float(out.reshape(-1)[0].real), float(out.reshape(-1)[0].imag)],
    "launches": runner.launches,
    "by_kernel": {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "GBs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1),
                      "TFs": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 2)} for k, v in cls.items()},
    "top_nodes": [{"kernel": KN[n.plan.kernel], "tile": [n.plan.tmb, n.plan.tnb, n.plan.tkb], "ms": round(t, 4),
                   "GBs": round(n.plan.bytes / t / 1e6, 1), "TFs": round(n.plan.flops / t / 1e9, 2),
                   "intensity": round(n.plan.flops / n.plan.bytes, 2)} for n, t in top]}))
