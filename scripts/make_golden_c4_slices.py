"""Golden fixture for the full-size config-4 circuit from the CPU oracle: partial sums of individual slices of the
CPU-sized plan (plans/c4_cpu.json: width 2^24, 2^21 slices) of the 53-qubit depth-14 amplitude <0..0|U|0..0>,
numpy.tensordot per tree node in complex128 (oracle/contract.py).  The GPU test contracts the SAME slices of the same
plan with the CUDA engine: an engine-independent check on the benchmark circuit itself (a whole amplitude is
10^17.5 flops on this plan -- out of the oracle's reach, so single slices are what can be pinned).

    python scripts/make_golden_c4_slices.py          # ~1 minute of CPU, writes tests/golden/c4_slice_partials.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import contract as oc  # noqa: E402
from qibotn_b200 import plans, workloads  # noqa: E402

from qibotn_b200.network import QiboCircuitToEinsum  # noqa: E402

wl = workloads.build("c4")
# gates built in complex128 (the workload's own network rounds them to complex64 first)
net = QiboCircuitToEinsum(wl.circuit, dtype="complex128").amplitude_network(wl.bitstring)
info = plans.load_plan("c4_cpu", net.modes, net.out)
assert info is not None
ops = [np.asarray(t, dtype=np.complex128) for t in net.tensors]
assert all(np.asarray(t).dtype == np.complex128 for t in net.tensors)
ids = [0, 1, 2, 3, 777, 123456, 1048576, info.num_slices - 1]
out = {"plan": "plans/c4_cpu.json", "num_slices": info.num_slices, "log2_width": info.log2_width,
       "sliced_labels": len(info.sliced_labels), "dtype": "complex128", "slices": {}}
t0 = time.time()
for s in ids:
    v = complex(np.asarray(oc.contract_ssa(ops, net.modes, net.out, info.path, sliced=info.sliced_labels,
                                           slice_id=s, dtype=np.complex128)).reshape(-1)[0])
    out["slices"][str(s)] = [v.real, v.imag]
    print(s, v, "%.1fs" % (time.time() - t0), flush=True)
first = sum(complex(*out["slices"][str(s)]) for s in (0, 1, 2, 3))
out["sum_0_to_3"] = [first.real, first.imag]
with open(os.path.join(ROOT, "tests", "golden", "c4_slice_partials.json"), "w") as f:
    json.dump(out, f, indent=1)
print("written", time.time() - t0)
