"""Batched Jacobi SVD microbenchmark through the C ABI (qtn_svd_rows on a batch of square complex matrices).

    python scripts/svd_bench.py [--batch 31] [--n 512] [--dtype complex128] [--decades 6] [--opt name=value ...]

Prints one JSON line: milliseconds per call (CUDA events), sweeps, ms per sweep, credited TFLOP/s with the
4*(14 p q^2 + 8 q^3) convention of SURVEY 8(d), and the rotation work actually done per sweep
(n(n-1)/2 row pairs x m columns x 20 real FMAs) as a rate."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qibotn_b200 import lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=31)
ap.add_argument("--n", type=int, default=512)
ap.add_argument("--dtype", default="complex128")
ap.add_argument("--decades", type=float, default=6.0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--max-sweeps", type=int, default=60)
ap.add_argument("--opt", action="append", default=[])
a = ap.parse_args()
for kv in a.opt:
    k, v = kv.split("=")
    lib.set_option(k, int(v))
h = lib.load()
n, nb = a.n, a.batch
tdt = torch.complex128 if a.dtype == "complex128" else torch.complex64
gen = torch.Generator(device="cuda").manual_seed(1)
s = 10.0 ** (-a.decades * torch.arange(n, device="cuda", dtype=torch.float64) / n)
mats = []
for _ in range(nb):
    g = torch.randn(n, n, dtype=torch.complex128, device="cuda", generator=gen)
    u, _ = torch.linalg.qr(g)
    g = torch.randn(n, n, dtype=torch.complex128, device="cuda", generator=gen)
    v, _ = torch.linalg.qr(g)
    mats.append(((u * s) @ v).to(tdt).contiguous())
code = lib.dtype_code(a.dtype)
ld = n
sigma = torch.empty(nb * ld, dtype=torch.float64, device="cuda")
perm = torch.empty(nb * ld, dtype=torch.int32, device="cuda")
nbytes = int(h.qtn_svd_workspace_bytes(nb, ld))
ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
trunc = lib.SvdTrunc(0.0, 0.0, 0.0, 0, 0)
st = torch.cuda.current_stream().cuda_stream
times, sweeps_seen = [], []
for rep in range(a.reps + 1):
    work = [m.clone() for m in mats]
    items = (lib.SvdItem * nb)()
    for k, w in enumerate(work):
        items[k].w, items[k].n, items[k].m = w.data_ptr(), n, n
    rank, scale, sweeps = (C.c_int32 * nb)(), (C.c_double * nb)(), C.c_int32(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    lib.check(h.qtn_svd_rows(code, nb, items, 0.0, a.max_sweeps, C.byref(trunc), sigma.data_ptr(), perm.data_ptr(),
                             ld, rank, scale, C.byref(sweeps), ws.data_ptr(), nbytes, st))
    e1.record()
    torch.cuda.synchronize()
    if rep:
        times.append(e0.elapsed_time(e1))
        sweeps_seen.append(int(sweeps.value))
err = float((sigma.view(nb, ld)[0] - s).abs().max())
ms = min(times)
sw = sweeps_seen[0]
credited = nb * 4.0 * (14 * n ** 3 + 8 * n ** 3)
rot = nb * (n * (n - 1) / 2) * n * 40.0 * sw
print(json.dumps({"workload": f"{nb} x ({n}x{n}) {a.dtype}, sigma over {a.decades} decades", "options": a.opt,
                  "ms": ms, "sweeps": sw, "ms_per_sweep": ms / sw, "credited_tflops": credited / ms / 1e9,
                  "rotation_tflops_upper": rot / ms / 1e9, "sigma_abs_err": err}), flush=True)
