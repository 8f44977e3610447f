"""Worker of tests/test_dist_gloo.py: one rank of a world_size-N gloo job on the CPU.
Exercises the host side of the sliced multi-GPU path: per-rank path
search + cost vote, the reference's slice partition, per-rank partial sums (computed here by
the CPU oracle -- this is a test), and the single sum of the partials."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import contract as oc, network as onet, statevector as osv  # noqa: E402
from qibotn_b200 import circuits, dist  # noqa: E402
from qibotn_b200.planner import find_path  # noqa: E402

rank, size, _ = dist.initialize(backend="gloo")
circ = circuits.brickwork(8, 5, seed=2)
inter = onet.state_vector_operands(circ, bitstring="0" * 8)
ops, modes, out = oc.split_interleaved(inter)
info = find_path(modes, out, samples=2, seed=rank, min_slices=max(8, size), target_log2_width=6)
best, winner = dist.vote_best(info.opt_cost, info)
costs = [None] * size
torch.distributed.all_gather_object(costs, float(info.opt_cost))
mine = dist.compute_slices(best.num_slices, rank, size)
partial = oc.contract_sliced(inter, best.path, best.sliced_labels, mine)
res = {}
for method in ("NCCL", "MPI"):
    t = torch.tensor([complex(partial)], dtype=torch.complex128)
    red = dist.reduce_result(t, method=method)
    res[method] = [float(red[0].real), float(red[0].imag)]
f32 = dist.reduce_result(torch.tensor([1 + 2j, 3 - 1j], dtype=torch.complex64) * (rank + 1), method="NCCL")
bad = None
try:
    dist.reduce_result(torch.zeros(2), method="NCCL")
except TypeError as exc:
    bad = str(exc)
want = complex(osv.simulate(circ)[0])
print("RESULT " + json.dumps({
    "rank": rank, "size": size, "winner": winner, "costs": costs, "best_cost": float(best.opt_cost),
    "slices": [mine.start, mine.stop], "num_slices": best.num_slices, "sum": res,
    "want": [want.real, want.imag], "f32": [float(f32[0].real), float(f32[1].imag)], "bad": bad}), flush=True)
dist.barrier()
torch.distributed.destroy_process_group()
