"""CTA-pair (cta_group::2) variant of the tcgen05 pair kernel against the single-CTA one (GPU box).

check mode: random layouts, both variants vs a complex128 einsum and vs each other (bitwise);
time mode : device-resident random data, median of 5 launches with an L2 flush in between."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from qibotn_b200 import lib  # noqa: E402

import os  # noqa: E402
if os.environ.get("QTN_LIB"):                     # A/B runs against another build of the library
    lib.LIB_PATH = os.environ["QTN_LIB"]
h = lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def plan_for(la, pa, lb, pb, lc, pair):
    lib.set_option("tc_pair", pair)
    lib.set_option("tc_min_log2", 0)
    desc = lib.make_pair_desc("complex64", list(zip(la, pa)), list(zip(lb, pb)), list(lc))
    plan = lib.pair_plan(desc)
    lib.set_option("tc_pair", 0)
    lib.set_option("tc_min_log2", 18)
    return desc, plan


def launch(plan, da, db, nc, reps=0):
    dc = torch.zeros(1 << nc, dtype=torch.complex64, device="cuda")
    ws = torch.empty(max(int(plan.workspace_bytes), 16), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    fn = lambda: lib.check(h.qtn_pair_run(C.byref(plan), da.data_ptr(), db.data_ptr(), dc.data_ptr(),
                                          ws.data_ptr(), None, st))
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return dc, (float(np.median(ts)) if ts else None)


def layouts(m, n, k, seed):
    rng = np.random.default_rng(seed)
    labels = list(range(m + n + k))
    M, N, K = labels[:m], labels[m:m + n], labels[m + n:]
    la, lb, lc = M + K, N + K, M + N
    return la, [int(x) for x in rng.permutation(len(la))], lb, [int(x) for x in rng.permutation(len(lb))], lc


def check(m, n, k, seed):
    la, pa, lb, pb, lc = layouts(m, n, k, seed)
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(1 << len(la), dtype=torch.complex64, device="cuda", generator=g)
    b = torch.randn(1 << len(lb), dtype=torch.complex64, device="cuda", generator=g)
    out = {}
    res = {"mode": "check", "m": m, "n": n, "k": k, "seed": seed}
    for tag, pair in (("single", 0), ("pair", 5)):
        desc, plan = plan_for(la, pa, lb, pb, lc, pair)
        dc, _ = launch(plan, a, b, len(lc))
        lcu = [desc.label_c[i] for i in range(len(lc))]
        pcu = [desc.pos_c[i] for i in range(len(lc))]
        # reference: complex128 einsum on the device, on dense views of the flat buffers
        def dense(flat, labels, pos):
            order = sorted(range(len(labels)), key=lambda i: -pos[i])       # axis order of the flat buffer
            t = flat.reshape((2,) * len(labels))
            return t, [labels[i] for i in order]
        ta, oa = dense(a.to(torch.complex128), la, pa)
        tb, ob = dense(b.to(torch.complex128), lb, pb)
        tc_, oc = dense(dc.to(torch.complex128), lcu, pcu)
        sym = lambda ls: "".join(chr(ord("a") + x) if x < 26 else chr(ord("A") + x - 26) for x in ls)
        want = torch.einsum(f"{sym(oa)},{sym(ob)}->{sym(oc)}", ta, tb)
        err = float((tc_ - want).abs().max() / want.abs().max())
        out[tag] = dc
        res[tag] = {"kernel": plan.kernel, "pair": plan.tc_pair, "tile": [plan.tmb, plan.tnb, plan.tkb],
                    "split": plan.split_bits, "stages": plan.tc_stages, "prepack": plan.tc_prepack, "rel_err": err}
    res["bitwise_equal"] = bool(torch.equal(out["single"], out["pair"]))
    print(json.dumps(res), flush=True)


def time_case(m, n, k):
    la, pa, lb, pb, lc = layouts(m, n, k, 0)
    a = torch.randn(1 << len(la), dtype=torch.complex64, device="cuda")
    b = torch.randn(1 << len(lb), dtype=torch.complex64, device="cuda")
    res = {"mode": "time", "m": m, "n": n, "k": k}
    for tag, pair in (("single", 0), ("pair", 5)):
        _, plan = plan_for(la, pa, lb, pb, lc, pair)
        _, ms = launch(plan, a, b, len(lc), reps=5)
        res[tag] = {"pair": plan.tc_pair, "tile": [plan.tmb, plan.tnb, plan.tkb], "stages": plan.tc_stages,
                    "ms": ms, "TFs": plan.flops / ms / 1e9, "GBs": plan.bytes / ms / 1e6}
    print(json.dumps(res), flush=True)


if sys.argv[1] == "one":                          # one launch pair (warm-up + 1) for ncu: one m n k [pair]
    m, n, k = (int(x) for x in sys.argv[2:5])
    la, pa, lb, pb, lc = layouts(m, n, k, 0)
    a = torch.randn(1 << len(la), dtype=torch.complex64, device="cuda")
    b = torch.randn(1 << len(lb), dtype=torch.complex64, device="cuda")
    _, plan = plan_for(la, pa, lb, pb, lc, int(sys.argv[5]) if len(sys.argv) > 5 else 0)
    _, ms = launch(plan, a, b, len(lc), reps=1)
    print(json.dumps({"mode": "one", "m": m, "n": n, "k": k, "pair": plan.tc_pair, "ms": ms}), flush=True)
elif sys.argv[1] == "check":
    for (m, n, k), seed in (((10, 7, 7), 1), ((12, 6, 5), 2), ((11, 5, 8), 3), ((13, 7, 4), 4), ((10, 9, 6), 5), ((14, 7, 9), 6)):
        check(m, n, k, seed)
else:
    for m, n, k in ((24, 7, 7), (22, 7, 7), (21, 10, 7), (19, 8, 8), (24, 6, 6), (24, 5, 5), (20, 7, 10)):
        time_case(m, n, k)
