"""One launch of an HBM-bound kernel on a tensor larger than L2, for ncu and for an event-timed GB/s figure:

    python scripts/hbm_one.py permute|transpose|pauli|pauli_x|absmax

permute: random bit-permutation of a rank-28 complex64 tensor (2 GB);  transpose: 16384 x 16384 complex64;
pauli: 16 Z-type terms fused in one pass over a 2^28 complex64 state;  pauli_x: 16 terms sharing an X/Y pattern;
absmax: the magnitude pass (max |re|, |im|) over 2^28 complex64 values."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from qibotn_b200 import lib  # noqa: E402

h = lib.load()
st = torch.cuda.current_stream().cuda_stream
kind = sys.argv[1]
rng = np.random.default_rng(0)
if kind == "permute":
    rank = 28
    x = torch.randn(1 << rank, dtype=torch.complex64, device="cuda")
    y = torch.empty_like(x)
    pin = [int(v) for v in rng.permutation(rank)]
    pout = [int(v) for v in rng.permutation(rank)]
    fn = lambda: lib.check(h.qtn_permute_bits(lib.QTN_C64, rank, lib.u8(pin), lib.u8(pout), x.data_ptr(), y.data_ptr(), 0, st))
    nbytes = 2 * x.numel() * 8
elif kind == "transpose":
    rows = cols = 16384
    x = torch.randn(rows * cols, dtype=torch.complex64, device="cuda")
    y = torch.empty_like(x)
    ext = (C.c_int64 * 2)(rows, cols)
    perm = (C.c_int32 * 2)(1, 0)
    fn = lambda: lib.check(h.qtn_permute(lib.QTN_C64, 2, ext, perm, x.data_ptr(), y.data_ptr(), st))
    nbytes = 2 * x.numel() * 8
elif kind in ("pauli", "pauli_x"):
    n = 28
    psi = torch.randn(1 << n, dtype=torch.complex64, device="cuda")
    nt = 16
    xmask = 0 if kind == "pauli" else (1 << 27) | (1 << 3)
    xm = (C.c_uint64 * nt)(*([xmask] * nt))
    zm = (C.c_uint64 * nt)(*[int(rng.integers(1, 1 << n)) for _ in range(nt)])
    ny = (C.c_int32 * nt)(*([0] * nt))
    out = torch.empty(nt, dtype=torch.complex128, device="cuda")
    fn = lambda: lib.check(h.qtn_pauli_expectation_host(lib.QTN_C64, n, psi.data_ptr(), nt, xm, zm, ny, out.data_ptr(), st))
    nbytes = psi.numel() * 8 * (1 if kind == "pauli" else 2)
elif kind == "absmax":
    x = torch.randn(1 << 28, dtype=torch.complex64, device="cuda")
    word = torch.zeros(1, dtype=torch.int32, device="cuda")
    fn = lambda: lib.check(h.qtn_absmax(lib.QTN_C64, x.numel(), x.data_ptr(), word.data_ptr(), st))
    nbytes = x.numel() * 8
else:
    raise SystemExit(__doc__)
fn(); fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"op": kind, "ms": ms, "algorithmic_bytes": nbytes, "GBs": nbytes / ms / 1e6}))
