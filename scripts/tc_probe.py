"""Diagnostics for the tcgen05 pair kernel (run on the GPU box).

1. structured probe: B = delta(n, k), A[r, k] = small integers -> C reveals every index mapping;
2. random cases vs a complex128 einsum: max error relative to max |C|;
3. timing of the same contraction on the tensor-core kernel and on the FFMA kernel.
Results go to stdout as JSON lines; mismatching arrays are dumped under gpurun_out/."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from qibotn_b200 import lib  # noqa: E402

h = lib.load()
os.makedirs("gpurun_out", exist_ok=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def dense(flat, labels, pos):
    idx = np.zeros((2,) * len(labels), dtype=np.int64)
    for ax, p in enumerate(pos):
        shape = [1] * len(labels)
        shape[ax] = 2
        idx = idx + (np.arange(2).reshape(shape) << p)
    return flat[idx]


def run(la, pa, lb, pb, lc, a, b, tc=True, time_it=False, persistent=True):
    lib.set_option("tc_persistent", 1 if persistent else 0)
    lib.set_option("tc_enable", 1 if tc else 0)
    lib.set_option("tc_min_log2", 0)
    desc = lib.make_pair_desc("complex64", list(zip(la, pa)), list(zip(lb, pb)), list(lc))
    plan = lib.pair_plan(desc)
    lib.set_option("tc_enable", 1)
    lib.set_option("tc_min_log2", 18)
    lib.set_option("tc_persistent", 1)
    da = a if torch.is_tensor(a) else torch.from_numpy(a).cuda()
    db = b if torch.is_tensor(b) else torch.from_numpy(b).cuda()
    dc = torch.zeros(1 << len(lc), dtype=torch.complex64, device="cuda")
    ws = torch.empty(max(int(plan.workspace_bytes), 16), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    fn = lambda: lib.check(h.qtn_pair_run(C.byref(plan), da.data_ptr(), db.data_ptr(), dc.data_ptr(),
                                          ws.data_ptr(), None, st))
    fn()
    torch.cuda.synchronize()
    ms = None
    if time_it:
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
    lcu = [desc.label_c[i] for i in range(len(lc))]
    pcu = [desc.pos_c[i] for i in range(len(lc))]
    if torch.is_tensor(a):                      # timing-only run: nothing to compare on the host
        return None, lcu, plan, ms
    return dense(dc.cpu().numpy(), lcu, pcu), lcu, plan, ms


def case(m, n, k, seed=0, layout="random", structured=False, time_it=False, check=True):
    rng = np.random.default_rng(seed)
    labels = list(range(m + n + k))
    M, N, K = labels[:m], labels[m:m + n], labels[m + n:]
    la, lb, lc = M + K, N + K, M + N
    if layout == "random":
        pa = [int(x) for x in rng.permutation(len(la))]
        pb = [int(x) for x in rng.permutation(len(lb))]
    else:
        pa = list(range(len(la)))[::-1]
        pb = list(range(len(lb)))[::-1]
    if not check and not structured:
        a = torch.randn(1 << len(la), dtype=torch.complex64, device="cuda")
        b = torch.randn(1 << len(lb), dtype=torch.complex64, device="cuda")
        res = {"m": m, "n": n, "k": k, "layout": layout, "structured": structured}
        for tag, tc, pers in (("tc", True, True), ("tc_np", True, False), ("ffma", False, False)):
            try:
                _, _, plan, ms = run(la, pa, lb, pb, lc, a, b, tc=tc, time_it=time_it, persistent=pers)
            except Exception as e:  # noqa: BLE001
                res[tag] = {"error": str(e)}
                continue
            res[tag] = {"kernel": plan.kernel, "tile": [plan.tmb, plan.tnb, plan.tkb], "split": plan.split_bits,
                        "stages": plan.tc_stages, "rel_err": -1.0, "ms": ms}
            if ms:
                res[tag]["TFs"] = plan.flops / ms / 1e9
                res[tag]["GBs"] = plan.bytes / ms / 1e6
        print(json.dumps(res), flush=True)
        return
    if structured:
        Ad = np.zeros((1 << m, 1 << k)); Bd = np.zeros((1 << n, 1 << k))
        Ad[:, :] = (np.arange(1 << m)[:, None] % 64) + 64 * np.arange(1 << k)[None, :]
        for i in range(min(1 << n, 1 << k)):
            Bd[i, i] = 1.0
        Ad = Ad + 1j * (Ad + 0.5)
        Bd = Bd.astype(np.complex128)
    else:
        Ad = rng.normal(size=(1 << m, 1 << k)) + 1j * rng.normal(size=(1 << m, 1 << k))
        Bd = rng.normal(size=(1 << n, 1 << k)) + 1j * rng.normal(size=(1 << n, 1 << k))
    # flat address-indexed buffers for the chosen layouts
    def flat(dense_arr, labels_, pos):
        r = len(labels_)
        t = dense_arr.reshape((2,) * r)
        out = np.zeros(1 << r, dtype=np.complex64)
        idx = np.zeros((2,) * r, dtype=np.int64)
        for ax, p in enumerate(pos):
            shape = [1] * r
            shape[ax] = 2
            idx = idx + (np.arange(2).reshape(shape) << p)
        out[idx] = t
        return out
    a, b = flat(Ad, la, pa), flat(Bd, lb, pb)
    if check:
        A64 = dense(a.astype(np.complex128), la, pa)
        B64 = dense(b.astype(np.complex128), lb, pb)
    res = {"m": m, "n": n, "k": k, "layout": layout, "structured": structured}
    for tag, tc, pers in (("tc", True, True), ("tc_np", True, False), ("ffma", False, False)):
        try:
            got, lcu, plan, ms = run(la, pa, lb, pb, lc, a, b, tc=tc, time_it=time_it, persistent=pers)
        except Exception as e:  # noqa: BLE001
            res[tag] = {"error": str(e)}
            continue
        if check:
            want = np.tensordot(A64, B64, axes=(list(range(m, m + k)), list(range(n, n + k))))
            want = np.transpose(want, [lc.index(x) for x in lcu])
            err = float(np.abs(got - want).max() / (np.abs(want).max() + 1e-30))
        else:
            want, err = None, -1.0
        res[tag] = {"kernel": plan.kernel, "tile": [plan.tmb, plan.tnb, plan.tkb], "split": plan.split_bits,
                    "stages": plan.tc_stages, "rel_err": err, "ms": ms}
        if ms:
            res[tag]["TFs"] = plan.flops / ms / 1e9
            res[tag]["GBs"] = plan.bytes / ms / 1e6
        if tag == "tc" and (err > 1e-4 or structured):
            name = f"gpurun_out/tc_probe_m{m}n{n}k{k}_{layout}{'_s' if structured else ''}"
            np.save(name + "_got.npy", got.reshape(1 << m, 1 << n) if lcu == lc else got)
            np.save(name + "_want.npy", want.reshape(1 << m, 1 << n) if lcu == lc else want)
            res["dump"] = name
            res["lc_order_kept"] = lcu == lc
    print(json.dumps(res), flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which == "one":        # python scripts/tc_probe.py one m n k [layout]  (short run for ncu)
    mm, nn, kk = (int(x) for x in sys.argv[2:5])
    case(mm, nn, kk, layout=sys.argv[5] if len(sys.argv) > 5 else "random", seed=8, time_it=True, check=False)
if which in ("probe", "all"):
    case(7, 4, 3, layout="gemm", structured=True)
    case(7, 4, 3, layout="gemm")
    case(7, 4, 3, layout="random", seed=1)
    case(7, 4, 4, layout="gemm", structured=True)
    case(7, 5, 4, layout="random", seed=2)
    case(8, 6, 5, layout="random", seed=3)
    case(7, 7, 6, layout="random", seed=4)
    case(7, 8, 7, layout="random", seed=5)
    case(9, 8, 4, layout="random", seed=6)
    case(7, 6, 12, layout="random", seed=7)      # split-K
if which in ("time", "all"):
    case(14, 9, 7, layout="random", seed=8, time_it=True)
    case(13, 8, 8, layout="random", seed=9, time_it=True)
    case(12, 12, 6, layout="random", seed=13, time_it=True)
    case(20, 9, 7, layout="random", seed=8, time_it=True, check=False)
    case(20, 9, 7, layout="gemm", seed=8, time_it=True, check=False)
    case(19, 8, 8, layout="random", seed=9, time_it=True, check=False)
    case(24, 5, 5, layout="random", seed=10, time_it=True, check=False)
    case(23, 4, 6, layout="random", seed=15, time_it=True, check=False)
    case(24, 5, 5, layout="gemm", seed=10, time_it=True, check=False)
    case(20, 6, 8, layout="random", seed=11, time_it=True, check=False)
    case(7, 6, 20, layout="random", seed=12, time_it=True, check=False)
    case(4, 4, 24, layout="random", seed=16, time_it=True, check=False)
    case(4, 4, 24, layout="gemm", seed=16, time_it=True, check=False)
    case(24, 7, 7, layout="random", seed=17, time_it=True, check=False)
    case(21, 10, 7, layout="random", seed=18, time_it=True, check=False)
    case(26, 5, 4, layout="random", seed=19, time_it=True, check=False)
    case(13, 13, 6, layout="random", seed=13, time_it=True, check=False)
    case(12, 12, 10, layout="random", seed=14, time_it=True, check=False)
