"""Offline, process-parallel plan search for a named workload; keeps the cheapest plan in plans/.
Usage: python scripts/search_plan.py <workload> [--procs 8] [--rounds 4] [--width 31] [--min-slices 32]"""
import argparse
import math
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qibotn_b200 import plans, workloads  # noqa: E402
from qibotn_b200.planner import find_path  # noqa: E402


def modeled_seconds(info, net, dtype, tc_tflops=320.0, ffma_tflops=25.0, hbm_tbs=4.1, launch_us=3.0,
                    tile_us=0.5, sms=148):
    """Time model of a plan on one B200: per tree node max(compute, bytes / HBM rate, launch), summed over
    slices (hoisted nodes once).  For the tcgen05 kernel compute = flops / 280 TFLOP/s + 3.2 us per
    128 x 2^tnb tile per SM (accumulator hand-off; with split accumulators a 128-column tile has no second
    TMEM buffer): fitted to profiles/r01_tc_issue_and_pair_ab.jsonl (m24n7k7 10.7 ms, m20n7k10 4.1 ms,
    m19n8k8 1.2 ms); HBM-bound nodes stream at ~3.5 TB/s (m24n5k5, m24n6k6).  Round 2 (staggered fp16 kernel,
    profiles/r02_gather_ab.jsonl): 320 TFLOP/s on tensor-bound nodes (m21n10k7 6.4 ms), 4.1 TB/s on HBM-bound ones
    (m24n7k7 7.9 ms, m25n6k6 9.0 ms), the accumulator hand-off hidden by the half tiles (0.5 us kept)."""
    from qibotn_b200 import executor, lib

    low = executor.lower(net.modes, net.out, info, dtype)
    total = 0.0
    for nd in low.nodes:
        p = nd.plan
        if p.kernel == lib.KERNEL_TC:
            tiles = float(p.grid)
            comp = p.flops / (tc_tflops * 1e12) + tile_us * 1e-6 * max(1.0, tiles / sms)
        else:
            comp = p.flops / (ffma_tflops * 1e12)
        t = max(comp, p.bytes / (hbm_tbs * 1e12), launch_us * 1e-6)
        total += t * (info.num_slices if nd.depends else 1)
    return total


def _one(args):
    name, seed, width, min_slices, samples, k = args
    wl = workloads.build(name)
    net = wl.network()
    info = find_path(net.modes, net.out, samples=samples, seed=seed, target_log2_width=width,
                     min_slices=min_slices, refine=6, subtree_size=k)
    info.extra = dict(info.extra, modeled_seconds=modeled_seconds(info, net, wl.dtype))
    return info


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--width", type=int, default=31)
    ap.add_argument("--min-slices", type=int, default=1)
    ap.add_argument("--samples", type=int, default=16)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--name", default=None)
    a = ap.parse_args()
    wl = workloads.build(a.workload)
    net = wl.network()
    name = a.name or a.workload
    best = plans.load_plan(name, net.modes, net.out)
    if best is not None and "modeled_seconds" not in best.extra:
        best.extra = dict(best.extra, modeled_seconds=modeled_seconds(best, net, wl.dtype))
    print("start:", None if best is None else "log10 flops %.2f, modeled %.1f s" % (
        math.log10(best.flops), best.extra["modeled_seconds"]), flush=True)
    t0 = time.time()
    with mp.Pool(a.procs) as pool:
        for r in range(a.rounds):
            jobs = [(a.workload, 1000 * r + p, a.width, a.min_slices, a.samples, a.k) for p in range(a.procs)]
            for info in pool.imap_unordered(_one, jobs):
                print("%.0fs: seed %d %s log10 flops %.2f W %d slices %d modeled %.1f s" % (
                    time.time() - t0, info.seed, info.method, math.log10(info.flops), info.log2_width,
                    info.num_slices, info.extra["modeled_seconds"]), flush=True)
                if best is None or info.extra["modeled_seconds"] < best.extra["modeled_seconds"]:
                    best = info
                    plans.save_plan(name, best, net.modes, net.out)
                    print("   -> new best (ranked by modeled wall time on one B200)", flush=True)
    print("done: log10 flops %.2f W %d slices %d modeled %.1f s" % (
        math.log10(best.flops), best.log2_width, best.num_slices, best.extra["modeled_seconds"]))
