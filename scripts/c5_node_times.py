"""Per-node device times of one slice of one config-5 term (GPU box): where that term's time goes.

    python scripts/c5_node_times.py --plan-store plans/c5_p4 --term 49 [--p 4]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from qibotn_b200 import circuits, engine, executor  # noqa: E402
from qibotn_b200 import eval as ev  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--plan-store", required=True)
ap.add_argument("--term", type=int, required=True)
ap.add_argument("--p", type=int, default=4)
ap.add_argument("--n", type=int, default=40)
ap.add_argument("--top", type=int, default=24)
a = ap.parse_args()
engine.PLAN_STORE = a.plan_store
circ, edges = circuits.qaoa_maxcut(a.n, a.p, seed=0)
net = list(ev._term_networks(circ, "complex64", circuits.zz_terms(edges)))[a.term]
info = engine._stored_plan(net, engine.default_width("complex64"), 1)
low = executor.lower(net.modes, net.out, info, "complex64")
runner = executor.ContractionRunner(low, use_graph=False)
runner.upload(net.tensors)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = runner.time_nodes(0, reps=3, flush=flush)
tot = sum(ms for _, ms in rows)
KN = {0: "tile", 1: "reduce", 2: "tc", 3: "kred"}
out = []
for node, ms in rows:
    P = node.plan
    out.append({"m": P.m_real + P.n_mhi - (P.n_batch if P.kernel == 2 else 0), "n": P.n_real + P.n_nhi,
                "k": P.k_real + P.n_khi, "batch": P.n_batch,
                "kernel": KN[P.kernel] + (f"_pair{P.tc_pair}" if P.kernel == 2 else ""), "kind": P.tc_kind,
                "tile": [P.tmb, P.tnb, P.tkb], "split": P.split_bits, "ms": round(ms, 4), "share": round(ms / tot, 4),
                "GBs": round(P.bytes / ms / 1e6, 1), "TFs": round(P.flops / ms / 1e9, 1)})
out.sort(key=lambda r: -r["ms"])
for r in out[:a.top]:
    print(json.dumps(r))
by_kernel = {}
for r in out:
    by_kernel[r["kernel"]] = round(by_kernel.get(r["kernel"], 0.0) + r["ms"], 3)
print(json.dumps({"summary": True, "term": a.term, "nodes": len(rows), "slices": info.num_slices,
                  "slice_ms_sum_of_nodes": tot, "ms_by_kernel": by_kernel,
                  "slice_TFs": low.flops_per_slice / tot / 1e9, "slice_GBs": low.bytes_per_slice / tot / 1e6}))
