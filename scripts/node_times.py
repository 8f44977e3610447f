"""Per-node device times of one slice of a committed plan (GPU box): where the slice's time goes.

    python scripts/node_times.py [workload] [plan name]

Prints one JSON line per node class (m, n, k, kernel) sorted by time, with the algorithmic GB/s and TFLOP/s
of each, and a summary line."""
import json
import sys

sys.path.insert(0, ".")
import torch  # noqa: E402

from qibotn_b200 import executor, lib, plans, workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
plan = sys.argv[2] if len(sys.argv) > 2 else name
wl = workloads.build(name)
net = wl.network()
info = plans.load_plan(plan, net.modes, net.out)
low = executor.lower(net.modes, net.out, info, wl.dtype)
runner = executor.ContractionRunner(low, use_graph=False)
runner.upload(net.tensors)
torch.cuda.synchronize()
rows = runner.time_nodes(0)
tot = sum(ms for _, ms in rows)
KN = {0: "tile", 1: "reduce", 2: "tc", 3: "kred"}
out = []
for node, ms in rows:
    P = node.plan
    ra, rb, rc = (len(low.tensors[t].labels) for t in (node.a, node.b, node.c))
    k = (ra + rb - rc) // 2
    m, n = max(ra, rb) - k, min(ra, rb) - k
    out.append({"m": m, "n": n, "k": k, "kernel": KN[P.kernel] + (f"_pair{P.tc_pair}" if P.kernel == 2 else ""),
                "kind": P.tc_kind, "tile_n": P.tnb, "ms": round(ms, 4), "share": round(ms / tot, 4),
                "GBs": round(P.bytes / ms / 1e6, 1), "TFs": round(P.flops / ms / 1e9, 1)})
out.sort(key=lambda r: -r["ms"])
for r in out[:24]:
    print(json.dumps(r))
print(json.dumps({"summary": True, "nodes": len(rows), "slice_ms_sum_of_nodes": tot,
                  "slice_TFs": low.flops_per_slice / tot / 1e9, "slice_GBs": low.bytes_per_slice / tot / 1e6}))
