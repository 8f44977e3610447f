"""BASELINE config 5 as named: 40-qubit QAOA MaxCut at p = 4 on a 3-regular graph, sum of the 60 ZZ expectations,
(term, slice) units sharded over the ranks by cost, one NCCL all-reduce of the per-term partials
(the reference's recipe: eval.py:316-382, one plan vote + reduce per term there).

    python scripts/run_c5_p4.py --plan-store plans/c5_p4                       # one GPU
    torchrun --nproc-per-node 8 ... scripts/run_c5_p4.py --plan-store plans/c5_p4 --distributed

Plans come from scripts/plan_c5.py (offline, CPU); a missing plan is an error here, not a search on the GPU box.
``--check-terms K``: the K cheapest terms are also contracted in complex128 with the same trees and the largest
complex64-vs-complex128 difference is reported (the whole sum in complex128 is 10^15.7 FP64 flops: not run).
``--max-seconds S`` stops launching new units after S seconds and reports what was covered (honest partial run)."""
import argparse
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qibotn_b200 import circuits, dist, engine  # noqa: E402
from qibotn_b200 import eval as ev  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--p", type=int, default=4)
ap.add_argument("--n", type=int, default=40)
ap.add_argument("--plan-store", required=True)
ap.add_argument("--distributed", action="store_true")
ap.add_argument("--check-terms", type=int, default=6)
ap.add_argument("--max-seconds", type=float, default=0.0)
ap.add_argument("--terms", default="", help="comma list: only these terms (debugging)")
ap.add_argument("--unit-seconds-from", default="", help="a one-GPU record of this script (rank0_terms): cut the units by "
                "the MEASURED seconds per slice of every term whose plan is still the one that was measured, instead of "
                "the time model's estimate")
ap.add_argument("--no-pool", action="store_true", help="every term allocates and frees its own arenas (A/B of ArenaPool)")
ap.add_argument("--pool-gib", type=float, default=64.0, help="first allocation of the arena pool (it grows if a term needs more)")
a = ap.parse_args()
if not a.no_pool:
    from qibotn_b200 import executor
    # one grow-only arena for all terms instead of a malloc and a free of tens of GiB per term
    executor.ARENA_POOL = executor.ArenaPool(reserve_bytes=int(a.pool_gib * 2 ** 30))
engine.PLAN_STORE = a.plan_store
rank, world = 0, 1
if a.distributed:
    rank, world, _ = dist.initialize()
dtype = "complex64"
circ, edges = circuits.qaoa_maxcut(a.n, a.p, seed=0)
obs = circuits.zz_terms(edges)
nets = list(ev._term_networks(circ, dtype, obs))
only = [int(x) for x in a.terms.split(",")] if a.terms else list(range(len(nets)))
width = engine.default_width(dtype)
infos = []
for k, net in enumerate(nets):
    info = engine._stored_plan(net, width, 1)
    if info is None:
        raise SystemExit(f"no stored plan for term {k} in {a.plan_store}: run scripts/plan_c5.py first")
    infos.append(info)
counts = [infos[k].num_slices if k in only else 0 for k in range(len(nets))]
weights = [infos[k].opt_cost / max(1, infos[k].num_slices) for k in range(len(nets))]          # MACs per unit
unit = ev.unit_seconds(nets, infos, dtype)                                                     # time model
if a.unit_seconds_from:
    with open(a.unit_seconds_from) as f:
        measured = json.loads([ln for ln in f.read().splitlines() if ln.startswith("{")][-1])["rank0_terms"]
    for k, row in measured.items():
        k = int(k)
        same_plan = abs(row["log10_flops"] - math.log10(max(1.0, 8.0 * infos[k].opt_cost))) < 2e-3
        if row["slices"] == row["of"] == infos[k].num_slices and same_plan:
            unit[k] = row["seconds"] / row["of"]
mine = dist.partition_units(counts, rank, world, unit)                                         # cut by time
my_flops = sum(8.0 * weights[t] * len(rg) for t, rg in mine.items())
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
partials = torch.zeros(len(nets), dtype=torch.complex64, device="cuda")
done_units = 0
per_term = {}
skipped = 0
for t, slices in mine.items():
    if a.max_seconds and time.perf_counter() - t0 > a.max_seconds:
        skipped += len(slices)
        continue
    t1 = time.perf_counter()
    partials[t] = engine.contract(nets[t], dtype, info=infos[t], slice_ids=slices).reshape(())
    torch.cuda.synchronize()
    per_term[t] = {"slices": len(slices), "of": infos[t].num_slices, "seconds": round(time.perf_counter() - t1, 4),
                   "log10_flops": round(math.log10(max(1.0, 8.0 * weights[t] * len(slices))), 3),
                   "launches": engine.LAST.launches}
    done_units += len(slices)
    engine.clear_caches()                      # the next term has another shape: drop its runner
    if a.no_pool:
        torch.cuda.empty_cache()               # ... and hand the arenas back
torch.cuda.synchronize()
my_seconds = time.perf_counter() - t0
partials = dist.reduce_result(partials, method="NCCL") if a.distributed else partials
torch.cuda.synchronize()
dist.barrier()
wall = time.perf_counter() - t0
value = float(partials.sum().real.item())
stats = torch.tensor([my_seconds, my_flops, float(skipped), float(done_units)], dtype=torch.float64, device="cuda")
if a.distributed and world > 1:
    import torch.distributed as td

    gathered = [torch.zeros_like(stats) for _ in range(world)]
    td.all_gather(gathered, stats)
else:
    gathered = [stats]
out = None
if rank == 0:
    g = [x.cpu().tolist() for x in gathered]
    total_flops = sum(8.0 * infos[k].opt_cost for k in only)
    out = {"workload": f"{a.n}-qubit QAOA MaxCut p={a.p} on a 3-regular graph, {len(edges)} ZZ terms, complex64",
           "ranks": world, "value": value, "complete": all(x[2] == 0 for x in g),
           "wall_seconds": wall, "log10_flops_total": math.log10(total_flops),
           "tflops_whole_job": total_flops / wall / 1e12 if all(x[2] == 0 for x in g) else None,
           "per_rank": [{"seconds": round(x[0], 3), "log10_flops": round(math.log10(max(1.0, x[1])), 3),
                         "units": int(x[3]), "units_skipped": int(x[2])} for x in g],
           "plan_widths": sorted({i.log2_width for i in infos}), "max_slices": max(i.num_slices for i in infos),
           "rank0_terms": per_term, "arena_pool": None if a.no_pool else {"grown": executor.ARENA_POOL.grown,
                                                                         "bytes": int(executor.ARENA_POOL.buf.numel())
                                                                         if executor.ARENA_POOL.buf is not None else 0},
           "note": "first and only call: plan load, lowering, arena allocation and graph capture are inside wall_seconds"}
# complex64 against complex128 on the cheapest terms, same trees (rank 0 only, after the timed region)
if rank == 0 and a.check_terms > 0:
    order = sorted(only, key=lambda k: infos[k].opt_cost)
    pick = [k for k in order if infos[k].log2_width <= 29 and infos[k].num_slices == 1][-a.check_terms:]
    worst, rows = 0.0, []
    for k in pick:
        engine.clear_caches(); torch.cuda.empty_cache()
        v64 = complex(engine.contract(nets[k], "complex64", info=infos[k]).reshape(()).item())
        engine.clear_caches(); torch.cuda.empty_cache()
        net128 = list(ev._term_networks(circ, "complex128", obs))[k]
        v128 = complex(engine.contract(net128, "complex128", info=infos[k]).reshape(()).item())
        rel = abs(v64 - v128) / max(abs(v128), 1e-30)
        rows.append({"term": k, "log10_flops": round(math.log10(8.0 * infos[k].opt_cost), 2), "c64": v64.real,
                     "c128": v128.real, "abs_diff": abs(v64 - v128), "rel_diff": rel})
        worst = max(worst, abs(v64 - v128))
    out["c64_vs_c128"] = {"terms": rows, "max_abs_diff": worst,
                          "note": "|<ZZ>| <= 1: the absolute difference is relative to the operator norm of a term"}
if rank == 0:
    print(json.dumps(out), flush=True)
if a.distributed:
    dist.barrier()
