"""Polish a stored plan by simulated annealing + subtree reconfiguration + re-slicing, ranked by the
modeled wall time on one B200 (scripts/search_plan.py::modeled_seconds).  Process-parallel; the best
plan found is written to plans/<out>.json.

    python scripts/anneal_plan.py c4 --start c4 --out c4_anneal --procs 6 --minutes 20
"""
import argparse
import importlib.util
import math
import multiprocessing as mp
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qibotn_b200 import plans, workloads  # noqa: E402
from qibotn_b200.planner import PathInfo, SymbolicNetwork, _absorb_small, path_cost  # noqa: E402
from qibotn_b200.treeopt import Tree  # noqa: E402

spec = importlib.util.spec_from_file_location("search_plan", os.path.join(ROOT, "scripts", "search_plan.py"))
search_plan = importlib.util.module_from_spec(spec)
spec.loader.exec_module(search_plan)


def to_info(net, pre, tree, nxt, order, sliced, seed, method):
    path = pre + tree.ssa_pairs(nxt)
    cost, width, per_slice, hoisted = path_cost(net, path, sliced)
    labels = [net.bit_label[low.bit_length() - 1] for low in order]
    return PathInfo(path=path, sliced_labels=labels, num_slices=1 << len(labels), opt_cost=cost,
                    flops=8.0 * cost, log2_width=width, per_slice_cost=per_slice, hoisted_cost=hoisted,
                    seed=seed, method=method)


def worker(args):
    name, start, seed, minutes, width = args
    wl = workloads.build(name)
    tn = wl.network()
    net = SymbolicNetwork(tn.modes, tn.out)
    pre, live, nxt = _absorb_small(net)
    if start == "fresh":                    # own starting tree: a different basin per worker
        from qibotn_b200.planner import find_path
        info0 = find_path(tn.modes, tn.out, samples=8, seed=seed, target_log2_width=width, anneal_steps=0)
    else:
        info0 = plans.load_plan(start, tn.modes, tn.out)
    rng = random.Random(seed)
    deadline = time.time() + 60 * minutes
    best = None
    tail = info0.path[len(pre):]
    sliced = net.to_mask(info0.sliced_labels)
    order = [1 << net.label_bit[lab] for lab in info0.sliced_labels]
    tree = Tree(live, tail, nxt, net.out_mask)
    round_no = 0
    while time.time() < deadline:
        t_start = rng.choice([1e-3, 3e-3, 1e-2]) if round_no else 1e-3
        tree.anneal(200000, rng, removed=sliced, t_start=t_start, t_end=t_start / 100, max_width=width,
                    deadline=deadline)
        tree.reconfigure(rounds=2, k=8, removed=sliced, deadline=deadline)
        # re-choose the sliced labels for the new tree (keeps them if still the best choice)
        if round_no % 2 == 1:
            order2, sliced2 = tree.slice_greedy(width, 1, k=8, reconf_every=0)
            if order2:
                order, sliced = order2, sliced2
        if tree.width(sliced) > width:
            more, sliced = tree.slice_greedy(width, 1, removed=sliced, reconf_every=0)
            order = order + more
        info = to_info(net, pre, tree, nxt, order, sliced, seed, "anneal")
        info.extra = {"modeled_seconds": search_plan.modeled_seconds(info, tn, wl.dtype), "start": start}
        if best is None or info.extra["modeled_seconds"] < best.extra["modeled_seconds"]:
            best = info
        round_no += 1
    return best


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--start", default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--procs", type=int, default=4)
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--width", type=int, default=31)
    a = ap.parse_args()
    start = a.start or a.workload
    out = a.out or (a.workload + "_anneal")
    wl = workloads.build(a.workload)
    tn = wl.network()
    if start != "fresh":
        base = plans.load_plan(start, tn.modes, tn.out)
        print("start: log10 flops %.2f modeled %.1f s" % (
            math.log10(base.flops), search_plan.modeled_seconds(base, tn, wl.dtype)), flush=True)
    with mp.Pool(a.procs) as pool:
        results = pool.map(worker, [(a.workload, start, 100 + p, a.minutes, a.width) for p in range(a.procs)])
    best = min(results, key=lambda r: r.extra["modeled_seconds"])
    for r in results:
        print("seed %d: log10 flops %.2f W %d slices %d modeled %.1f s" % (
            r.seed, math.log10(r.flops), r.log2_width, r.num_slices, r.extra["modeled_seconds"]), flush=True)
    plans.save_plan(out, best, tn.modes, tn.out)
    print("saved plans/%s.json: modeled %.1f s" % (out, best.extra["modeled_seconds"]))
