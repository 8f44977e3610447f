"""BASELINE config 1: 10-qubit QFT (with swaps) of a random normalised dense state, complex128, through the reference's
legacy call ``eval_qu.dense_vector_tn_qu(qasm, initial_state, mps_opts)`` (reference eval_qu.py:21-46,
tests/test_quimb_backend.py:23-66) -- contraction route (mps_opts None) and MPS route -- against the closed form
(QFT = normalised inverse FFT), with the CPU oracle timed beside it.  Also larger sizes to show where the GPU starts
to pay: at 10 qubits the work is a few thousand flops and the call is launch- and planning-bound.
    python scripts/c1_bench.py [--sizes 10,16,20] [--repeats 5]"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import statevector as osv  # noqa: E402  (the checker / CPU baseline, not the product path)
from qibotn_b200 import circuits, engine, eval_qu  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="10,16,20")
ap.add_argument("--repeats", type=int, default=5)
ap.add_argument("--mps-max", type=int, default=16, help="largest size run through the MPS route (a random state has "
                "full bond dimension: 2^(n/2); the gate loop then is hundreds of large SVDs)")
a = ap.parse_args()


def qft_qasm(n):
    lines = ['OPENQASM 2.0;', 'include "qelib1.inc";', f"qreg q[{n}];"]
    for i in range(n):
        lines.append(f"h q[{i}];")
        for j in range(i + 1, n):
            lines.append(f"cu1({math.pi / 2 ** (j - i)!r}) q[{j}],q[{i}];")
    for i in range(n // 2):
        lines.append(f"swap q[{i}],q[{n - 1 - i}];")
    return "\n".join(lines) + "\n"


def timed(fn, repeats):
    fn()                                                     # first call: plan search, lowering, graph capture
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return out, best


rows = []
for n in [int(x) for x in a.sizes.split(",")]:
    rng = np.random.default_rng(n)
    init = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    init /= np.linalg.norm(init)
    want = np.fft.ifft(init) * math.sqrt(1 << n)             # closed form, qubit 0 most significant
    text = qft_qasm(n)
    row = {"qubits": n, "gates": n + n * (n - 1) // 2 + n // 2}
    t0 = time.perf_counter()
    first = eval_qu.dense_vector_tn_qu(text, init, None)
    row["tn_first_call_s"] = time.perf_counter() - t0
    got, row["tn_s"] = timed(lambda: eval_qu.dense_vector_tn_qu(text, init, None), a.repeats)
    row["tn_max_err"] = float(np.abs(got - want).max())
    row["tn_launches"] = engine.LAST.launches
    opts = {"method": "svd", "cutoff": 1e-6, "cutoff_mode": "abs"}            # the reference test's gate_opts
    if n <= a.mps_max:
        got, row["mps_s"] = timed(lambda: eval_qu.dense_vector_tn_qu(text, init, opts), max(1, a.repeats // 2))
        row["mps_max_err"] = float(np.abs(got - want).max())
    if n <= 22:
        t0 = time.perf_counter()
        ref = osv.simulate(circuits.qft(n, True), initial_state=init)
        row["cpu_oracle_s"] = time.perf_counter() - t0
        row["cpu_oracle_max_err"] = float(np.abs(ref - want).max())
    rows.append(row)
    engine.clear_caches()
print(json.dumps({"workload": "QFT with swaps on a random normalised state, complex128, eval_qu.dense_vector_tn_qu "
                              "(QASM text in, host array out; timed region includes parsing, upload and read-back)",
                  "device": torch.cuda.get_device_name(0), "rows": rows}))
