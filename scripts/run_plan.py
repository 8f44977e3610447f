"""Run a stored plan of a named workload in full (all slices) on the current GPU and print the
amplitude, the wall time and the achieved TFLOP/s.  Used to check path independence at full size:
two different plans of the same network must give the same number.

    python scripts/run_plan.py c4 --plan c4 [--plan c4_v1 ...] [--dtype complex64]
"""
import argparse
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qibotn_b200 import executor, plans, workloads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--plan", action="append", required=True)
ap.add_argument("--max-slices", type=int, default=0)
ap.add_argument("--dtype", default=None, help="override the workload dtype (complex64 | complex128)")
ap.add_argument("--no-tc", action="store_true", help="FFMA/DFMA kernels only")
ap.add_argument("--tc-terms", type=int, default=3)
ap.add_argument("--accum-split", type=int, default=1)
ap.add_argument("--tc-pair", type=int, default=None, help="CTA-pair kernel for tiles of >= 2^N columns (0 = off)")
ap.add_argument("--lib", default=None, help="another build of libqibotn_b200.so (A/B runs)")
a = ap.parse_args()
wl = workloads.build(a.workload)
if a.dtype:
    wl.dtype = a.dtype
from qibotn_b200 import lib  # noqa: E402
if a.lib:
    lib.LIB_PATH = os.path.abspath(a.lib)
if a.tc_pair is not None:
    lib.set_option("tc_pair", a.tc_pair)
if a.no_tc:
    lib.set_option("tc_enable", 0)
lib.set_option("tc_terms", a.tc_terms)
lib.set_option("tc_accum_split", a.accum_split)
net = wl.network()
for name in a.plan:
    info = plans.load_plan(name, net.modes, net.out)
    if info is None:
        print(json.dumps({"plan": name, "dtype": wl.dtype, "tensor_cores": not a.no_tc, "tc_terms": a.tc_terms, "accum_split": a.accum_split, "error": "missing or made for another network"}), flush=True)
        continue
    low = executor.lower(net.modes, net.out, info, wl.dtype)
    runner = executor.ContractionRunner(low, use_graph=True)
    runner.upload(net.tensors)
    ns = info.num_slices if a.max_slices <= 0 else min(info.num_slices, a.max_slices)
    runner.run(range(min(2, ns)))                       # warm-up (graph capture)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = runner.run(range(ns))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    amp = complex(out.reshape(-1)[0].item())
    flops = low.flops_hoisted + ns * low.flops_per_slice
    print(json.dumps({"plan": name, "dtype": wl.dtype, "tensor_cores": not a.no_tc, "tc_terms": a.tc_terms, "accum_split": a.accum_split, "tc_pair": lib.get_option("tc_pair") if a.lib is None or "old" not in a.lib else 0, "lib": a.lib, "log10_flops": math.log10(info.flops), "num_slices": info.num_slices,
                      "slices_run": ns, "log2_width": info.log2_width, "seconds": dt, "tflops": flops / dt / 1e12,
                      "amplitude": [amp.real, amp.imag], "slice_arena_gb": low.slice_bytes / 2**30,
                      "launches_per_slice": low.launches_per_slice}), flush=True)
    del runner
    torch.cuda.empty_cache()
