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
# many-term observable: (term, slice) units sharded over the ranks, one reduction of the per-term partials
from qibotn_b200.network import QiboCircuitToEinsum  # noqa: E402
from qibotn_b200.observables import check_observable, extract_gates_and_qubits, get_ham_gates  # noqa: E402

qcirc, edges = circuits.qaoa_maxcut(8, 2, seed=3)
conv = QiboCircuitToEinsum(qcirc, "complex128")
terms = list(extract_gates_and_qubits(check_observable(circuits.zz_terms(edges), 8)))
nets = [conv.expectation_network(get_ham_gates(t, dtype="complex128"), hyper=True) for t in terms]
infos = []
for k, net in enumerate(nets):
    info_k = find_path(net.modes, net.out, samples=1, seed=rank, min_slices=1 if k % 3 else 4)
    infos.append(dist.vote_best(info_k.opt_cost, info_k)[0])
units = dist.partition_units([i.num_slices for i in infos], rank, size)
part = torch.zeros(len(nets), dtype=torch.complex128)
for tk, sl in units.items():
    part[tk] = complex(np.asarray(oc.contract_sliced(nets[tk].interleaved(), infos[tk].path,
                                                     infos[tk].sliced_labels, sl)).reshape(-1)[0])
ham_sum = dist.reduce_result(part, method="NCCL").sum()
psi = osv.simulate(qcirc)
ham_want = osv.pauli_expectation(psi, [(1.0, [("Z", a), ("Z", b)]) for a, b in edges], 8)
print("RESULT " + json.dumps({
    "rank": rank, "size": size, "winner": winner, "costs": costs, "best_cost": float(best.opt_cost),
    "slices": [mine.start, mine.stop], "num_slices": best.num_slices, "sum": res,
    "want": [want.real, want.imag], "ham": [float(ham_sum.real), float(ham_want.real)],
    "units": {str(k): [v.start, v.stop] for k, v in units.items()}, "unit_counts": [i.num_slices for i in infos],
    "f32": [float(f32[0].real), float(f32[1].imag)], "bad": bad}), flush=True)
dist.barrier()
torch.distributed.destroy_process_group()
