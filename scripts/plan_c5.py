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
ap.add_argument("--procs", type=int, default=1, help="terms planned in parallel")
ap.add_argument("--seconds", type=float, default=None, help="time budget of one term's search")
ap.add_argument("--refine", type=int, default=3)
a = ap.parse_args()
engine.PLAN_STORE = a.store or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                            "plans", f"c5_p{a.p}")
circ, edges = circuits.qaoa_maxcut(a.n, a.p, seed=0)
obs = circuits.zz_terms(edges)
summary = {}


def plan_term(job):
    dtype, k, net, width = job
    t0 = time.perf_counter()
    info = engine._stored_plan(net, width, 1)
    if info is None:
        info = find_path(net.modes, net.out, samples=a.samples, seed=0, target_log2_width=width, min_slices=1,
                         time_budget_s=a.seconds, refine=a.refine)
        engine._store_plan(net, width, 1, info)
    return dtype, k, len(net.tensors), info.flops, info.log2_width, info.num_slices, info.method, time.perf_counter() - t0


if __name__ == "__main__":
    for dtype in a.dtypes.split(","):
        width = engine.default_width(dtype)
        total, worst_w, t_all = 0.0, 0, time.perf_counter()
        jobs = [(dtype, k, net, width) for k, net in enumerate(ev._term_networks(circ, dtype, obs))]
        per_term = []
        if a.procs > 1:
            from multiprocessing import Pool

            with Pool(a.procs) as pool:
                results = list(pool.imap_unordered(plan_term, jobs))
        else:
            results = [plan_term(j) for j in jobs]
        for _d, k, nt, flops, w, ns, meth, secs in sorted(results, key=lambda r: r[1]):
            total += flops
            worst_w = max(worst_w, w)
            per_term.append({"term": k, "tensors": nt, "log10_flops": round(math.log10(max(flops, 1.0)), 3),
                             "log2_width": w, "slices": ns, "method": meth})
            print(dtype, k, nt, "log10 flops %.2f" % math.log10(max(flops, 1.0)), "width", w, "slices", ns, meth,
                  "%.1fs" % secs, flush=True)
        summary[dtype] = {"log10_flops_total": math.log10(total), "max_log2_width": worst_w,
                          "seconds": time.perf_counter() - t_all, "terms": per_term}
    print(json.dumps({"workload": f"{a.n}-qubit QAOA MaxCut p={a.p}, {len(edges)} ZZ terms", "store": engine.PLAN_STORE,
                      "plans": summary}), flush=True)
