"""Kernel micro-benchmarks (CUDA events, L2 flushed between timed launches).
Usage: python scripts/microbench.py [pair|permute|all]  -> JSON lines on stdout."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from qibotn_b200 import lib  # noqa: E402

PEAKS = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}
h = lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


def pair_case(dtype, m, n, k, seed=0, layout="random"):
    rng = np.random.default_rng(seed)
    labels = list(range(m + n + k))
    M, N, K = labels[:m], labels[m:m + n], labels[m + n:]
    la, lb = M + K, K + N
    if layout == "random":
        pa = [int(x) for x in rng.permutation(len(la))]
        pb = [int(x) for x in rng.permutation(len(lb))]
    else:   # "gemm": row-major (m,k) x (k,n)
        pa = list(range(len(la)))[::-1]; pb = list(range(len(lb)))[::-1]
    tdt = torch.complex64 if dtype == "complex64" else torch.complex128
    a = torch.randn(1 << len(la), dtype=tdt, device="cuda")
    b = torch.randn(1 << len(lb), dtype=tdt, device="cuda")
    c = torch.zeros(1 << (m + n), dtype=tdt, device="cuda")
    desc = lib.make_pair_desc(dtype, list(zip(la, pa)), list(zip(lb, pb)), M + N)
    plan = lib.pair_plan(desc)
    ws = torch.empty(max(int(plan.workspace_bytes), 16), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    fn = lambda: lib.check(h.qtn_pair_run(C.byref(plan), a.data_ptr(), b.data_ptr(), c.data_ptr(),
                                          ws.data_ptr(), None, st))
    med, best = timeit(fn)
    gbs = plan.bytes / med / 1e6
    tfs = plan.flops / med / 1e9
    print(json.dumps({"op": "pair", "dtype": dtype, "m": m, "n": n, "k": k, "layout": layout,
                      "kernel": plan.kernel, "tile": [plan.tmb, plan.tnb, plan.tkb], "split": plan.split_bits,
                      "ms": round(med, 4), "ms_best": round(best, 4), "GBs": round(gbs, 1),
                      "TFs": round(tfs, 2), "hbm_frac": round(gbs / PEAKS["hbm_gbs"], 3)}), flush=True)


def permute_case(dtype, rank, kind, seed=0):
    rng = np.random.default_rng(seed)
    tdt = torch.complex64 if dtype == "complex64" else torch.complex128
    x = torch.randn(1 << rank, dtype=tdt, device="cuda")
    y = torch.empty_like(x)
    pin = list(range(rank))
    if kind == "random":
        pout = [int(v) for v in rng.permutation(rank)]
    elif kind == "reverse":
        pout = pin[::-1]
    else:       # swap low and high halves (matrix transpose)
        half = rank // 2
        pout = [(p + half) % rank for p in pin]
    st = torch.cuda.current_stream().cuda_stream
    fn = lambda: lib.check(h.qtn_permute_bits(lib.dtype_code(dtype), rank, lib.u8(pin), lib.u8(pout),
                                              x.data_ptr(), y.data_ptr(), 0, st))
    med, best = timeit(fn)
    nbytes = 2 * x.numel() * x.element_size()
    print(json.dumps({"op": "permute", "dtype": dtype, "rank": rank, "kind": kind, "ms": round(med, 4),
                      "GBs": round(nbytes / med / 1e6, 1),
                      "hbm_frac": round(nbytes / med / 1e6 / PEAKS["hbm_gbs"], 3)}), flush=True)
    # torch reference point for the same data movement
    t = x.view((2,) * rank)
    perm = sorted(range(rank), key=lambda l: -pout[l])
    perm_axes = [rank - 1 - pin[l] for l in perm]
    fn2 = lambda: torch.permute(t, perm_axes).contiguous()
    if rank <= 26:
        med2, _ = timeit(fn2)
        print(json.dumps({"op": "permute_torch", "dtype": dtype, "rank": rank, "kind": kind,
                          "ms": round(med2, 4), "GBs": round(nbytes / med2 / 1e6, 1)}), flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("pair", "all"):
    for dt in ("complex64", "complex128"):
        big = 26 if dt == "complex64" else 25
        for (m, n, k) in [(big, 2, 2), (big - 2, 4, 4), (big - 1, 1, 1), (big - 4, 6, 3), (13, 13, 6), (12, 12, 10),
                          (11, 11, 12), (1, 1, 24), (0, 0, 26), (3, 3, 20), (8, 8, 12)]:
            for layout in ("random", "gemm"):
                pair_case(dt, m, n, k, layout=layout)
if which in ("permute", "all"):
    for dt in ("complex64", "complex128"):
        for kind in ("random", "reverse", "transpose"):
            permute_case(dt, 27 if dt == "complex64" else 26, kind)
