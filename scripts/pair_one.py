"""One pairwise contraction through the C ABI, for ncu:  python scripts/pair_one.py dtype m n k [batch] [gemm|random]
(two warm-up launches, then one; prints the plan's kernel and the event time)."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from qibotn_b200 import lib  # noqa: E402

dtype = sys.argv[1]
m, n, k = (int(x) for x in sys.argv[2:5])
batch = int(sys.argv[5]) if len(sys.argv) > 5 else 0
layout = sys.argv[6] if len(sys.argv) > 6 else "random"
h = lib.load()
rng = np.random.default_rng(0)
labels = list(range(m + n + k + batch))
M, N, K, Bt = labels[:m], labels[m:m + n], labels[m + n:m + n + k], labels[m + n + k:]
la, lb = M + K + Bt, K + N + Bt
if layout == "random":
    pa = [int(x) for x in rng.permutation(len(la))]
    pb = [int(x) for x in rng.permutation(len(lb))]
else:
    pa = list(range(len(la)))[::-1]; pb = list(range(len(lb)))[::-1]
tdt = torch.complex64 if dtype == "complex64" else torch.complex128
a = torch.randn(1 << len(la), dtype=tdt, device="cuda")
b = torch.randn(1 << len(lb), dtype=tdt, device="cuda")
c = torch.zeros(1 << (m + n + batch), dtype=tdt, device="cuda")
desc = lib.make_pair_desc(dtype, list(zip(la, pa)), list(zip(lb, pb)), M + N + Bt)
plan = lib.pair_plan(desc)
ws = torch.empty(max(int(plan.workspace_bytes), 16), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
fn = lambda: lib.check(h.qtn_pair_run(C.byref(plan), a.data_ptr(), b.data_ptr(), c.data_ptr(), ws.data_ptr(), None, st))
fn(); fn()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); flush.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"dtype": dtype, "m": m, "n": n, "k": k, "batch": batch, "kernel": plan.kernel,
                  "tile": [plan.tmb, plan.tnb, plan.tkb], "ms": ms, "GBs": plan.bytes / ms / 1e6, "TFs": plan.flops / ms / 1e9}))
