"""Plan the light-cone networks of BASELINE config 5 (40-qubit QAOA MaxCut, 60 ZZ terms) offline, on CPU
cores, into an on-disk plan store that scripts/run_c5.py --plan-store reads on the GPU box.
    python scripts/plan_c5.py --p 3 [--samples 8] [--store plans/c5_p3]
Prints one line per term (tensors, log10 flops, width, slices, seconds) and a JSON summary."""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qibotn_b200 import circuits, engine  # noqa: E402
from qibotn_b200 import eval as ev  # noqa: E402
from qibotn_b200.planner import find_path  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--p", type=int, default=3)
ap.add_argument("--n", type=int, default=40)
ap.add_argument("--samples", type=int, default=8)
ap.add_argument("--dtypes", default="complex64,complex128")
ap.add_argument("--store", default=None)
a = ap.parse_args()
engine.PLAN_STORE = a.store or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            "plans", f"c5_p{a.p}")
circ, edges = circuits.qaoa_maxcut(a.n, a.p, seed=0)
obs = circuits.zz_terms(edges)
summary = {}
for dtype in a.dtypes.split(","):
    width = engine.default_width(dtype)
    total, worst_w, t_all = 0.0, 0, time.perf_counter()
    for k, net in enumerate(ev._term_networks(circ, dtype, obs)):
        t0 = time.perf_counter()
        info = engine._stored_plan(net, width, 1)
        if info is None:
            info = find_path(net.modes, net.out, samples=a.samples, seed=0, target_log2_width=width, min_slices=1)
            engine._store_plan(net, width, 1, info)
        total += info.flops
        worst_w = max(worst_w, info.log2_width)
        print(dtype, k, len(net.tensors), "log10 flops %.2f" % math.log10(max(info.flops, 1.0)), "width", info.log2_width,
              "slices", info.num_slices, "%.1fs" % (time.perf_counter() - t0), flush=True)
    summary[dtype] = {"log10_flops_total": math.log10(total), "max_log2_width": worst_w,
                      "seconds": time.perf_counter() - t_all}
print(json.dumps({"workload": f"{a.n}-qubit QAOA MaxCut p={a.p}, {len(edges)} ZZ terms", "store": engine.PLAN_STORE,
                  "plans": summary}), flush=True)
