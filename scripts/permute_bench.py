"""Bit-tensor permute and rank-2 transpose: achieved HBM GB/s (2 * N * sizeof(element) bytes per launch,
SURVEY 8d) on tensors larger than L2, CUDA events, median of 5 launches after 2 warm-ups.

    python scripts/permute_bench.py            # table
    python scripts/permute_bench.py one        # warm-up + one launch of the rank-28 complex64 case (for ncu)"""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from qibotn_b200 import lib  # noqa: E402

h = lib.load()
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def permute_case(rank, dtype, seed):
    rng = np.random.default_rng(seed)
    tdt = torch.complex64 if dtype == "complex64" else torch.complex128
    x = torch.randn(1 << rank, dtype=tdt, device="cuda")
    y = torch.empty_like(x)
    pin = [int(v) for v in rng.permutation(rank)]
    pout = [int(v) for v in rng.permutation(rank)]
    fn = lambda: lib.check(h.qtn_permute_bits(lib.dtype_code(dtype), rank, lib.u8(pin), lib.u8(pout), x.data_ptr(),
                                              y.data_ptr(), 0, st))
    ms = timed(fn)
    nbytes = 2 * (1 << rank) * (8 if dtype == "complex64" else 16)
    return {"op": "permute_bits", "rank": rank, "dtype": dtype, "seed": seed, "ms": ms, "GBs": nbytes / ms / 1e6}


def transpose_case(rows, cols, dtype):
    tdt = torch.complex64 if dtype == "complex64" else torch.complex128
    x = torch.randn(rows * cols, dtype=tdt, device="cuda")
    y = torch.empty_like(x)
    ext = (C.c_int64 * 2)(rows, cols)
    perm = (C.c_int32 * 2)(1, 0)
    fn = lambda: lib.check(h.qtn_permute(lib.dtype_code(dtype), 2, ext, perm, x.data_ptr(), y.data_ptr(), st))
    ms = timed(fn)
    nbytes = 2 * rows * cols * (8 if dtype == "complex64" else 16)
    ok = bool(torch.equal(y.view(cols, rows), x.view(rows, cols).t()))
    return {"op": "transpose", "shape": [rows, cols], "dtype": dtype, "ms": ms, "GBs": nbytes / ms / 1e6, "exact": ok}


if len(sys.argv) > 1 and sys.argv[1] == "one":
    print(json.dumps(permute_case(28, "complex64", 0)))
else:
    copy_src = torch.randn(1 << 28, dtype=torch.complex64, device="cuda")
    copy_dst = torch.empty_like(copy_src)
    ms = timed(lambda: copy_dst.copy_(copy_src))
    print(json.dumps({"op": "torch copy_ (reference)", "ms": ms, "GBs": 2 * (8 << 28) / ms / 1e6}), flush=True)
    del copy_src, copy_dst
    for dtype, rank in (("complex64", 28), ("complex128", 27)):
        for seed in (0, 1, 2):
            print(json.dumps(permute_case(rank, dtype, seed)), flush=True)
    for shape in ((16384, 16384), (512, 512), (4096, 70000)):
        for dtype in ("complex64", "complex128"):
            print(json.dumps(transpose_case(*shape, dtype)), flush=True)
