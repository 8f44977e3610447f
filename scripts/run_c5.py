"""BASELINE config 5's workload at small depth: 40-qubit QAOA MaxCut on a 3-regular graph, sum of the 60 ZZ
expectations, one light-cone-reduced contraction per term, complex64 and complex128, first call and cached call.
    python scripts/run_c5.py [--p 3] [--plan-store plans/c5_p3]
Config 5 as named (p = 4, 10^15.7 flops) is scripts/run_c5_p4.py with the plans of scripts/plan_c5.py --p 4."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qibotn_b200 import circuits, engine  # noqa: E402
from qibotn_b200 import eval as ev  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--p", type=int, default=3)
ap.add_argument("--n", type=int, default=40)
ap.add_argument("--samples", type=int, default=4)
ap.add_argument("--dtypes", default="complex64,complex128")
ap.add_argument("--distributed", action="store_true",
                help="under torchrun: (term, slice) units sharded over the ranks, one NCCL all-reduce (expectation_tn_nccl)")
ap.add_argument("--plan-store", default=None, help="directory of plans made offline by scripts/plan_c5.py")
a = ap.parse_args()
if a.plan_store:
    engine.PLAN_STORE = a.plan_store
rank = 0
if a.distributed:
    from qibotn_b200 import dist
    rank, world, _ = dist.initialize()
circ, edges = circuits.qaoa_maxcut(a.n, a.p, seed=0)
obs = circuits.zz_terms(edges)
out = {}
ev.PLAN_SAMPLES = a.samples
dtypes = a.dtypes.split(",")
for dtype in dtypes:
    ev.EXPECTATION_ROUTE = "network"
    engine.clear_caches()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    def run():
        if a.distributed:
            return float(ev.expectation_tn_nccl(circ, dtype, obs, a.samples)[0].real.get().reshape(-1)[0])
        return float(ev.expectation_tn(circ, dtype, obs).real.get().reshape(-1)[0])
    val = run()
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    flops = sum(low.flops_hoisted + inf.num_slices * low.flops_per_slice for inf, low in engine._PLAN_CACHE.values())
    t0 = time.perf_counter()
    val2 = run()                                                                       # plans cached
    torch.cuda.synchronize()
    cached = time.perf_counter() - t0
    out[dtype] = {"value": val, "seconds_first_call": first, "seconds_cached_plans": cached,
                  "repeat_equal": val == val2, "log10_flops": __import__("math").log10(max(flops, 1.0)),
                  "tflops_cached_plans": flops / cached / 1e12}
out["workload"] = f"{a.n}-qubit QAOA MaxCut p={a.p}, {len(edges)} ZZ terms"
if a.distributed:
    out["ranks"] = world
    out["note"] = "flops / tflops are this rank's share of the (term, slice) units"
if len(dtypes) == 2:
    out["rel_diff_c64_c128"] = abs(out["complex64"]["value"] - out["complex128"]["value"]) / max(1e-3, abs(out["complex128"]["value"]))
if rank == 0:
    print(json.dumps(out), flush=True)
if a.distributed:
    dist.barrier()
