"""Staggered CTA-pair kernel (plan value tc_pair = 2, 128-column tiles as two 64-column halves) against
the round-1 pair kernel (tc_stagger = 0) and a complex128 einsum (GPU box).

check : random layouts; both kernels vs the einsum and vs each other
time  : device-resident random data, median of 5 launches with an L2 flush in between
one   : warm-up + one launch of one shape (for ncu):  one m n k [stagger]"""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from qibotn_b200 import lib  # noqa: E402

h = lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
DEFAULTS = dict(tc_pair=5, tc_stagger=1, tc_fp16=1, tc_min_log2=18, tc_gather=lib.get_option("tc_gather"))


def plan_for(la, pa, lb, pb, lc, **opts):
    try:
        for key, val in dict(opts, tc_min_log2=0).items():
            lib.set_option(key, val)
        desc = lib.make_pair_desc("complex64", list(zip(la, pa)), list(zip(lb, pb)), list(lc))
        plan = lib.pair_plan(desc)
    finally:
        for key, val in DEFAULTS.items():
            lib.set_option(key, val)
    return desc, plan


def launch(plan, da, db, nc, reps=0):
    dc = torch.zeros(1 << nc, dtype=torch.complex64, device="cuda")
    ws = torch.empty(max(int(plan.workspace_bytes), 16), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    # magnitude words of A, B (measured once, as the executor does for leaves) and C (written by the kernel)
    amax = torch.zeros(4, dtype=torch.int32, device="cuda")
    lib.check(h.qtn_absmax(lib.QTN_C64, da.numel(), da.data_ptr(), amax.data_ptr(), st))
    lib.check(h.qtn_absmax(lib.QTN_C64, db.numel(), db.data_ptr(), amax.data_ptr() + 4, st))
    fn = lambda: lib.check(h.qtn_pair_run_scaled(C.byref(plan), da.data_ptr(), db.data_ptr(), dc.data_ptr(),
                                                 ws.data_ptr(), None, amax.data_ptr(), amax.data_ptr() + 4,
                                                 amax.data_ptr() + 8, st))
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    launch.amax = amax.cpu().numpy().view(np.float32)[:3].tolist()
    launch.cmax = float(torch.view_as_real(dc).abs().max())
    return dc, (float(np.median(ts)) if ts else None)


def layouts(m, n, k, seed):
    rng = np.random.default_rng(seed)
    labels = list(range(m + n + k))
    M, N, K = labels[:m], labels[m:m + n], labels[m + n:]
    la, lb, lc = M + K, N + K, M + N
    return la, [int(x) for x in rng.permutation(len(la))], lb, [int(x) for x in rng.permutation(len(lb))], lc


def dense(flat, labels, pos):
    order = sorted(range(len(labels)), key=lambda i: -pos[i])       # axis order of the flat buffer
    return flat.reshape((2,) * len(labels)), [labels[i] for i in order]


def sym(ls):
    return "".join(chr(ord("a") + x) if x < 26 else chr(ord("A") + x - 26) for x in ls)


def check(m, n, k, seed):
    la, pa, lb, pb, lc = layouts(m, n, k, seed)
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(1 << len(la), dtype=torch.complex64, device="cuda", generator=g)
    b = torch.randn(1 << len(lb), dtype=torch.complex64, device="cuda", generator=g)
    out = {}
    res = {"mode": "check", "m": m, "n": n, "k": k, "seed": seed}
    for tag, opts in (("pair", dict(tc_stagger=0)), ("stagger", dict(tc_fp16=0)), ("fp16", dict()),
                      ("fp16_own", dict(tc_gather=0))):
        desc, plan = plan_for(la, pa, lb, pb, lc, **opts)
        dc, _ = launch(plan, a, b, len(lc))
        lcu = [desc.label_c[i] for i in range(len(lc))]
        pcu = [desc.pos_c[i] for i in range(len(lc))]
        ta, oa = dense(a.to(torch.complex128), la, pa)
        tb, ob = dense(b.to(torch.complex128), lb, pb)
        tc_, oc = dense(dc.to(torch.complex128), lcu, pcu)
        want = torch.einsum(f"{sym(oa)},{sym(ob)}->{sym(oc)}", ta, tb)
        err = float((tc_ - want).abs().max() / want.abs().max())
        out[tag] = dc
        res[tag] = {"kernel": plan.kernel, "pair": plan.tc_pair, "kind": plan.tc_kind, "gather": plan.tc_gather, "tile": [plan.tmb, plan.tnb, plan.tkb],
                    "split": plan.split_bits, "stages": plan.tc_stages, "grid": int(plan.grid), "rel_err": err,
                    "amax_abc": launch.amax, "true_cmax": launch.cmax}
        del ta, tb, tc_, want
    res["max_abs_diff"] = float((out["pair"] - out["stagger"]).abs().max())
    res["bitwise_equal"] = bool(torch.equal(out["pair"], out["stagger"]))
    print(json.dumps(res), flush=True)


def time_case(m, n, k):
    la, pa, lb, pb, lc = layouts(m, n, k, 0)
    a = torch.randn(1 << len(la), dtype=torch.complex64, device="cuda")
    b = torch.randn(1 << len(lb), dtype=torch.complex64, device="cuda")
    res = {"mode": "time", "m": m, "n": n, "k": k}
    for tag, opts in (("fp16_own", dict(tc_gather=0)), ("fp16_gather", dict(tc_gather=1)),
                      ("fp16_own_again", dict(tc_gather=0)), ("fp16_gather_again", dict(tc_gather=1)),
                      ("tf32_gather", dict(tc_fp16=0, tc_gather=1))):
        _, plan = plan_for(la, pa, lb, pb, lc, **opts)
        _, ms = launch(plan, a, b, len(lc), reps=5)
        res[tag] = {"pair": plan.tc_pair, "kind": plan.tc_kind, "gather": plan.tc_gather, "tile": [plan.tmb, plan.tnb, plan.tkb], "stages": plan.tc_stages,
                    "ms": ms, "TFs": plan.flops / ms / 1e9, "GBs": plan.bytes / ms / 1e6}
    print(json.dumps(res), flush=True)


if sys.argv[1] == "one":
    m, n, k = (int(x) for x in sys.argv[2:5])
    la, pa, lb, pb, lc = layouts(m, n, k, 0)
    a = torch.randn(1 << len(la), dtype=torch.complex64, device="cuda")
    b = torch.randn(1 << len(lb), dtype=torch.complex64, device="cuda")
    mode = int(sys.argv[5]) if len(sys.argv) > 5 else 2          # 0 pair, 1 staggered tf32, 2 staggered fp16
    _, plan = plan_for(la, pa, lb, pb, lc, tc_stagger=int(mode > 0), tc_fp16=int(mode > 1))
    _, ms = launch(plan, a, b, len(lc), reps=1)
    print(json.dumps({"mode": "one", "m": m, "n": n, "k": k, "pair": plan.tc_pair, "ms": ms,
                      "TFs": plan.flops / ms / 1e9}), flush=True)
elif sys.argv[1] == "check":
    # one tile per pair; several n tiles; deep k (32 chunks); split-K; many tiles per CTA pair (the stagger
    # crosses tile boundaries); k of one chunk only
    for (m, n, k), seed in (((10, 7, 7), 1), ((11, 9, 5), 2), ((10, 7, 9), 3), ((10, 7, 4), 4), ((17, 7, 6), 5),
                            ((15, 8, 5), 6), ((13, 7, 3), 7), ((18, 8, 7), 8), ((12, 6, 6), 9), ((18, 6, 5), 10),
                            ((12, 5, 7), 11), ((19, 5, 4), 12), ((13, 6, 3), 13), ((13, 5, 3), 14)):
        check(m, n, k, seed)
else:
    for m, n, k in ((21, 10, 7), (24, 7, 7), (23, 7, 8), (25, 6, 6), (24, 6, 6), (23, 5, 7), (20, 7, 10)):
        time_case(m, n, k)
