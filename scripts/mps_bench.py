"""MPS evolution benchmark (BASELINE config 3 and reduced variants).

    python scripts/mps_bench.py [--n 64] [--steps 20] [--chi 256] [--dtype complex128] [--cpu-seconds 0]

Prints one JSON line: wall time of the gate loop (CUDA-synchronised), gates/s (circuit gates, swaps
excluded), bond profile, <psi|psi>, SVD sweep statistics; with --cpu-seconds > 0 also times the CPU
oracle (numpy.linalg.svd gate loop) for that many seconds and reports its gate rate."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qibotn_b200 import circuits, mps as gmps  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=64)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--chi", type=int, default=256)
ap.add_argument("--dtype", default="complex128")
ap.add_argument("--cpu-seconds", type=float, default=0.0)
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--tol", type=float, default=0.0, help="Jacobi convergence threshold (0 = library default)")
ap.add_argument("--lq-min-rows", type=int, default=-1, help="override mps.LQ_MIN_ROWS (0 = no LQ preconditioner)")
ap.add_argument("--shots", type=int, default=0, help="also draw this many samples from the final MPS (timed)")
ap.add_argument("--phases", action="store_true", help="time the LQ preconditioner separately (extra syncs)")
ap.add_argument("--opt", action="append", default=[], help="library option name=value (qtn_set_option)")
a = ap.parse_args()
from qibotn_b200 import lib as _lib  # noqa: E402
for kv in a.opt:
    _k, _v = kv.split("=")
    _lib.set_option(_k, int(_v))
gmps.PROFILE_PHASES = a.phases
if a.lq_min_rows >= 0:
    gmps.LQ_MIN_ROWS = a.lq_min_rows

circ = circuits.tfim_c3(a.n, a.steps)          # config 3 angles (circuits.C3_ANGLES)
algo = {"qr_method": False, "svd_method": {"partition": "UV", "abs_cutoff": 1e-12, "max_extent": a.chi}}
ngates = sum(1 for g in circ.queue)
best = None
for _ in range(a.repeat + 1):          # first pass warms up (module load, allocator)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    conv = gmps.QiboCircuitToMPS(circ, algo, dtype=a.dtype, tol=a.tol)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
helper = gmps.MPSContractionHelper(a.n)
norm = float(helper.contract_norm(conv.mps_tensors).cpu())
zz = complex(helper.pauli_expectation(conv.mps_tensors, {a.n // 2: "Z", a.n // 2 + 1: "Z"}).cpu())
bonds = [int(t.shape[2]) for t in conv.mps_tensors[:-1]]
line = {"workload": f"{a.n}-site TFIM Trotter, {a.steps} steps, chi<={a.chi}, {a.dtype}", "gates": ngates,
        "seconds": best, "gates_per_s": ngates / best, "max_bond": max(bonds), "bonds": bonds, "norm": norm,
        "zz_mid": [zz.real, zz.imag], "stats": conv.stats, "tol": a.tol, "lq_min_rows": gmps.LQ_MIN_ROWS, "options": a.opt}
if a.shots > 0:
    helper.sample(conv.mps_tensors, 16, seed=0)                       # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    bits, probs = helper.sample(conv.mps_tensors, a.shots, seed=1)
    dt = time.perf_counter() - t0
    line["sampling"] = {"shots": a.shots, "seconds": dt, "shots_per_s": a.shots / dt,
                        "mean_log2_prob_over_norm": float(np.mean(np.log2(np.maximum(probs / norm, 1e-300)))),
                        "first": "".join(map(str, bits[0]))}
if a.cpu_seconds > 0:
    from oracle import mps as omps
    from oracle.network import gate_tensors
    gts, _ = gate_tensors(circ, np.complex128 if a.dtype == "complex128" else np.complex64)
    sites = omps.initial(a.n, np.complex128 if a.dtype == "complex128" else np.complex64)
    t0 = time.perf_counter()
    done = 0
    for g, qs in gts:
        omps.apply_gate(sites, g, tuple(qs), algo)
        done += 1
        if time.perf_counter() - t0 > a.cpu_seconds:
            break
    wall = time.perf_counter() - t0
    line["cpu"] = {"gates_done": done, "seconds": wall, "gates_per_s": done / wall,
                   "max_bond_reached": max(t.shape[2] for t in sites), "cores": len(os.sched_getaffinity(0)),
                   "note": "oracle gate loop (numpy.linalg.svd), first gates only: bonds are still small early on"}
print(json.dumps(line), flush=True)
