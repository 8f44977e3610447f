"""GPU parity of the pairwise contraction and permute kernels through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from qibotn_b200 import lib

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _dense(flat, labels, pos, off=0):
    idx = np.zeros((2,) * len(labels), dtype=np.int64) + off
    for ax, p in enumerate(pos):
        shape = [1] * len(labels)
        shape[ax] = 2
        idx = idx + (np.arange(2).reshape(shape) << p)
    return flat[idx]


def run_pair(dtype, la, pa, lb, pb, lc, pc, a, b, c0=None, slice_a=(), slice_b=(), slice_id=0,
             accumulate=False):
    h = lib.load()
    tdt = torch.complex64 if dtype == "complex64" else torch.complex128
    free = pc is None
    desc = lib.make_pair_desc(dtype, list(zip(la, pa)), list(zip(lb, pb)),
                              list(lc) if free else list(zip(lc, pc)),
                              accumulate=accumulate, slice_a=slice_a, slice_b=slice_b)
    plan = lib.pair_plan(desc)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    dc = torch.zeros(1 << len(lc), dtype=tdt, device="cuda") if c0 is None else torch.from_numpy(c0).cuda()
    ws = torch.empty(max(int(plan.workspace_bytes), 16), dtype=torch.uint8, device="cuda")
    sid = torch.tensor([slice_id], dtype=torch.int32, device="cuda")
    lib.check(h.qtn_pair_run(C.byref(plan), da.data_ptr(), db.data_ptr(), dc.data_ptr(),
                             ws.data_ptr(), sid.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    lc_used = [desc.label_c[i] for i in range(len(lc))]
    pc_used = [desc.pos_c[i] for i in range(len(lc))]
    return _dense(dc.cpu().numpy(), lc_used, pc_used), lc_used, plan


CASES = [
    (6, 6, 3, 0), (7, 7, 5, 0), (12, 2, 2, 0), (3, 13, 2, 0), (1, 1, 1, 0), (0, 0, 3, 0),
    (2, 0, 1, 0), (9, 8, 6, 0), (4, 4, 2, 2), (7, 3, 2, 1), (1, 1, 14, 0), (2, 2, 16, 0),
    (0, 0, 20, 0), (16, 0, 1, 0), (6, 7, 0, 0), (3, 3, 12, 0), (10, 10, 4, 0), (14, 3, 3, 0),
    (5, 5, 10, 0), (8, 8, 8, 0), (4, 4, 14, 0), (3, 4, 17, 0), (4, 2, 12, 0), (1, 4, 10, 0),
]


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("case", CASES)
def test_pair_contraction_matches_einsum(case, dtype):
    m, n, k, batch = case
    rng = np.random.default_rng(1000 * m + 100 * n + 10 * k + batch)
    labels = list(range(m + n + k + batch))
    M, N, K, Bt = labels[:m], labels[m:m + n], labels[m + n:m + n + k], labels[m + n + k:]
    la, lb, lc = M + K + Bt, K + N + Bt, M + N + Bt
    rng.shuffle(la); rng.shuffle(lb); rng.shuffle(lc)
    pa = [int(x) for x in rng.permutation(len(la))]
    pb = [int(x) for x in rng.permutation(len(lb))]
    pc = [int(x) for x in rng.permutation(len(lc))]
    a = (rng.normal(size=1 << len(la)) + 1j * rng.normal(size=1 << len(la))).astype(dtype)
    b = (rng.normal(size=1 << len(lb)) + 1j * rng.normal(size=1 << len(lb))).astype(dtype)
    A = _dense(a.astype(np.complex128), la, pa)
    B = _dense(b.astype(np.complex128), lb, pb)
    tol = 2e-6 if dtype == "complex64" else 1e-13
    for prescribed in (False, True):
        got, lc_used, plan = run_pair(dtype, la, pa, lb, pb, lc, pc if prescribed else None, a, b)
        want = np.einsum(A, la, B, lb, lc_used)
        scale = np.abs(want).max() + 1e-30
        assert np.abs(got - want).max() / scale < tol * max(1, k), (case, prescribed, plan.kernel)


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
def test_slices_and_accumulate(dtype):
    rng = np.random.default_rng(5)
    # A stored with 2 extra (sliced) bits, B with 1; C accumulates over 8 slice ids
    la, pa, sa = [0, 1, 2, 3, 4, 5, 6], [8, 0, 3, 4, 6, 1, 7], [(0, 2), (2, 5)]
    lb, pb, sb = [4, 5, 6, 7, 8], [0, 5, 2, 3, 1], [(1, 4)]
    lc, pc = [0, 1, 2, 3, 7, 8], [5, 0, 1, 4, 3, 2]
    a = (rng.normal(size=1 << 9) + 1j * rng.normal(size=1 << 9)).astype(dtype)
    b = (rng.normal(size=1 << 6) + 1j * rng.normal(size=1 << 6)).astype(dtype)
    c = np.zeros(1 << 6, dtype=dtype)
    want = np.zeros((2,) * 6, dtype=np.complex128)
    for sid in range(8):
        offa = sum(((sid >> bt) & 1) << p for bt, p in sa)
        offb = sum(((sid >> bt) & 1) << p for bt, p in sb)
        want += np.einsum(_dense(a.astype(np.complex128), la, pa, offa), la,
                          _dense(b.astype(np.complex128), lb, pb, offb), lb, lc)
        got, _, _ = run_pair(dtype, la, pa, lb, pb, lc, pc, a, b, c0=c, slice_a=sa, slice_b=sb,
                             slice_id=sid, accumulate=True)
        c = np.zeros(1 << 6, dtype=dtype)
        idx = np.zeros((2,) * 6, dtype=np.int64)
        for ax, p in enumerate(pc):
            shape = [1] * 6; shape[ax] = 2
            idx = idx + (np.arange(2).reshape(shape) << p)
        c[idx] = got
    tol = 1e-5 if dtype == "complex64" else 1e-12
    assert np.abs(got - want).max() < tol * np.abs(want).max()


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("rank", [0, 1, 5, 12, 13, 20, 24])
def test_permute_bits(rank, dtype):
    h = lib.load()
    rng = np.random.default_rng(rank)
    x = (rng.normal(size=1 << rank) + 1j * rng.normal(size=1 << rank)).astype(dtype)
    pin = [int(v) for v in rng.permutation(rank)]
    pout = [int(v) for v in rng.permutation(rank)]
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros_like(dx)
    lib.check(h.qtn_permute_bits(lib.dtype_code(dtype), rank, lib.u8(pin), lib.u8(pout),
                                 dx.data_ptr(), dy.data_ptr(), 0,
                                 torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    labels = list(range(rank))
    assert np.array_equal(_dense(dy.cpu().numpy(), labels, pout), _dense(x, labels, pin))


def test_general_permute():
    h = lib.load()
    rng = np.random.default_rng(0)
    x = (rng.normal(size=(3, 2, 5, 4)) + 1j * rng.normal(size=(3, 2, 5, 4))).astype(np.complex128)
    perm = [2, 0, 3, 1]
    dx = torch.from_numpy(x).cuda()
    dy = torch.zeros(5, 3, 4, 2, dtype=torch.complex128, device="cuda")
    ext = (C.c_int64 * 4)(*x.shape)
    pm = (C.c_int32 * 4)(*perm)
    lib.check(h.qtn_permute(lib.QTN_C128, 4, ext, pm, dx.data_ptr(), dy.data_ptr(),
                            torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(dy.cpu().numpy(), x.transpose(perm))


@pytest.mark.parametrize("shape", [(11, 7, 5), (10, 8, 4), (12, 5, 7), (10, 4, 3), (10, 7, 3), (15, 7, 7), (12, 9, 5),
                                   (14, 10, 7)])
def test_tensor_core_variants_agree_with_einsum(shape):
    """Every switch of the tcgen05 path (persistent or one CTA per tile, pre-packed B or gathered B,
    split or single accumulators, 3 or 4 split terms) against a complex128 einsum, with a sliced
    label on B so that the pre-pack kernel sees the slice offset."""
    m, n, k = shape
    rng = np.random.default_rng(7 * m + 3 * n + k)
    labels = list(range(m + n + k))
    M, N, K = labels[:m], labels[m:m + n], labels[m + n:]
    la, lb, lc = M + K, K + N, M + N
    rng.shuffle(la); rng.shuffle(lb)
    pa = [int(x) for x in rng.permutation(len(la))]
    pb_all = [int(x) for x in rng.permutation(len(lb) + 1)]
    pb, sb = pb_all[:-1], [(0, pb_all[-1])]
    a = (rng.normal(size=1 << len(la)) + 1j * rng.normal(size=1 << len(la))).astype(np.complex64)
    b = (rng.normal(size=1 << (len(lb) + 1)) + 1j * rng.normal(size=1 << (len(lb) + 1))).astype(np.complex64)
    A = _dense(a.astype(np.complex128), la, pa)
    B = _dense(b.astype(np.complex128), lb, pb, off=1 << pb_all[-1])          # slice id 1
    variants = [dict(), dict(tc_persistent=0), dict(tc_prepack=0), dict(tc_accum_split=0),
                dict(tc_terms=4), dict(tc_prepack=0, tc_accum_split=0), dict(tc_pair=0), dict(tc_pair=0, tc_terms=4), dict(tc_pair=7),
                dict(tc_stagger=0), dict(tc_fp16=0), dict(tc_gather=1), dict(tc_gather=1, tc_fp16=0), dict(tc_prepack_a=0),
                dict(tc_gather=0), dict(tc_gather=0, tc_prepack_a=0), dict(tc_gather=2, tc_fp16=0), dict(tc_prepack_a=2)]
    defaults = dict(tc_persistent=1, tc_prepack=1, tc_accum_split=1, tc_terms=3, tc_min_log2=18, tc_pair=5, tc_stagger=1,
                    tc_fp16=1, tc_gather=2, tc_prepack_a=1)
    wants = {}
    for opts in variants:
        try:
            for key, val in dict(opts, tc_min_log2=0).items():
                lib.set_option(key, val)
            got, lc_used, plan = run_pair("complex64", la, pa, lb, pb, lc, None, a, b, slice_b=sb, slice_id=1)
        finally:
            for key, val in defaults.items():
                lib.set_option(key, val)
        assert plan.kernel == lib.KERNEL_TC
        want_pair = opts.get("tc_pair", 5)
        paired = (want_pair and n >= want_pair and opts.get("tc_persistent", 1) and opts.get("tc_prepack", 1)
                  and opts.get("tc_accum_split", 1))
        # paired tiles take the staggered kernel (two half-width halves per tile, plan value 2)
        staggered = paired and opts.get("tc_stagger", 1) and opts.get("tc_terms", 3) == 3
        assert plan.tc_pair == (2 if staggered else 1 if paired else 0), (shape, opts)
        # ... and, with 16 contracted values per stage, the scaled 3xFP16 split (plain qtn_pair_run then measures
        # the operands' magnitudes itself)
        assert plan.tc_kind == (1 if staggered and k >= 4 and opts.get("tc_fp16", 1) else 0), (shape, opts)
        want_pa = opts.get("tc_prepack_a", 1)                                         # >= 4 n tiles per A tile, or forced
        prepack_a = plan.tc_kind == 1 and want_pa and (n >= 9 or want_pa == 2)
        assert plan.tc_prepack_a == (1 if prepack_a else 0), (shape, opts)
        want_g = opts.get("tc_gather", 2)
        if not staggered or prepack_a or want_g == 0:
            assert plan.tc_gather == 0, (shape, opts)
        elif want_g == 1:
            assert plan.tc_gather == 1, (shape, opts)
        else:
            assert plan.tc_gather in (0, 2), (shape, opts)      # bulk copies when the tile owns >= 6 low address bits
        if tuple(lc_used) not in wants:
            wants[tuple(lc_used)] = np.einsum(A, la, B, lb, lc_used, optimize=True)
        want = wants[tuple(lc_used)]
        assert np.abs(got - want).max() / np.abs(want).max() < 4e-6 * max(1, k), (shape, opts)


@pytest.mark.parametrize("case", [
    # kind, m, n, k, batch: batch (hyper) labels -- the diagonal gates of QAOA networks (BASELINE config 5) -- on the
    # staggered tcgen05 pair kernel (one set of B images per batch value) and on the streaming K reduction (grid.y)
    ("tc", 10, 5, 4, 2), ("tc", 11, 7, 5, 3), ("tc", 10, 6, 3, 1), ("tc", 12, 7, 6, 2), ("tc", 10, 5, 8, 4),
    ("tc", 10, 4, 4, 1), ("tc", 13, 4, 5, 3), ("tc", 11, 4, 3, 2), ("tc", 14, 4, 6, 5),       # 16 columns: phantom column bit
    ("kred", 4, 4, 12, 3), ("kred", 2, 3, 14, 6), ("kred", 0, 1, 16, 1), ("kred", 4, 3, 10, 9),
    ("kred", 5, 3, 14, 4), ("kred", 5, 5, 12, 0), ("kred", 2, 5, 13, 2),
])
def test_batch_labels_on_tensor_core_and_kred_kernels(case):
    kind, m, n, k, batch = case
    rng = np.random.default_rng(31 * m + 17 * n + 5 * k + batch)
    labels = list(range(m + n + k + batch))
    M, N, K, Bt = labels[:m], labels[m:m + n], labels[m + n:m + n + k], labels[m + n + k:]
    la, lb, lc = M + K + Bt, K + N + Bt, M + N + Bt
    rng.shuffle(la); rng.shuffle(lb); rng.shuffle(lc)
    pa_all = [int(x) for x in rng.permutation(len(la) + 1)]
    pa, sa = pa_all[:-1], [(0, pa_all[-1])]                   # one sliced label on A, slice id 1
    pb = [int(x) for x in rng.permutation(len(lb))]
    a = (rng.normal(size=1 << (len(la) + 1)) + 1j * rng.normal(size=1 << (len(la) + 1))).astype(np.complex64)
    b = (rng.normal(size=1 << len(lb)) + 1j * rng.normal(size=1 << len(lb))).astype(np.complex64)
    A = _dense(a.astype(np.complex128), la, pa, off=1 << pa_all[-1])
    B = _dense(b.astype(np.complex128), lb, pb)
    dtypes = ["complex64"] if kind == "tc" else ["complex64", "complex128"]
    for dtype in dtypes:
        try:
            lib.set_option("tc_min_log2", 0)
            got, lc_used, plan = run_pair(dtype, la, pa, lb, pb, lc, None, a.astype(dtype), b.astype(dtype),
                                          slice_a=sa, slice_id=1)
        finally:
            lib.set_option("tc_min_log2", 18)
        if kind == "kred":
            assert plan.n_batch == batch - plan.kred_inner + max(0, m - 4) + max(0, n - 4)
        else:
            assert plan.n_batch == batch
        if kind == "tc":
            assert plan.kernel == lib.KERNEL_TC and plan.tc_pair == 2 and plan.tc_kind == (1 if k >= 4 else 0)
        else:
            assert plan.kernel == lib.KERNEL_KRED
        want = np.einsum(A, la, B, lb, lc_used, optimize=True)
        tol = 4e-6 if dtype == "complex64" else 1e-13
        assert np.abs(got - want).max() / np.abs(want).max() < tol * max(1, k), (case, dtype)


@pytest.mark.parametrize("shape", [(15, 7, 7, 0), (16, 6, 5, 1), (13, 5, 4, 2), (17, 7, 3, 0), (14, 8, 6, 0)])
def test_bulk_copied_raw_chunks_of_the_staggered_kernel(shape):
    """tc_gather = 2: when the labels of a tile own the lowest address bits of A -- the usual case, A being
    an earlier node's output tile -- a chunk reaches the raw ring by one cp.async.bulk per contiguous run and is
    converted across threads.  Layouts built so that the run has 6, 7 or more bits (some k labels among the low bits,
    a sliced label in the middle), against a complex128 einsum and against the per-thread gather (tc_gather = 0)."""
    m, n, k, variant = shape
    rng = np.random.default_rng(11 * m + 5 * n + k + variant)
    labels = list(range(m + n + k))
    M, N, K = labels[:m], labels[m:m + n], labels[m + n:]
    la, lb, lc = M + K, K + N, M + N
    # A: `low` labels at address bits 0..: variant 0 = m labels only, 1 = two k labels among them, 2 = run of 6 bits
    low = {0: M[:8], 1: M[:4] + K[:2] + M[4:7], 2: M[:5] + K[:1]}[variant]
    rest = [x for x in la if x not in low]
    rng.shuffle(rest)
    order = low + rest
    if variant == 2:
        order = low + [None] + rest                   # a sliced label at bit 6 ends the run
    pa = {lab: p for p, lab in enumerate(order) if lab is not None}
    sa = [(0, order.index(None))] if variant == 2 else []
    ra = len(order)
    pb = [int(x) for x in rng.permutation(len(lb))]
    a = (rng.normal(size=1 << ra) + 1j * rng.normal(size=1 << ra)).astype(np.complex64)
    b = (rng.normal(size=1 << len(lb)) + 1j * rng.normal(size=1 << len(lb))).astype(np.complex64)
    pa_list = [pa[x] for x in la]
    A = _dense(a.astype(np.complex128), la, pa_list, off=(1 << sa[0][1]) if sa else 0)
    B = _dense(b.astype(np.complex128), lb, pb)
    got = {}
    for g in (2, 0):
        try:
            lib.set_option("tc_min_log2", 0)
            lib.set_option("tc_gather", g)
            lib.set_option("tc_gather_min_run", 6)    # default 9 (4 KB runs): here every run length is exercised
            got[g], lc_used, plan = run_pair("complex64", la, pa_list, lb, pb, lc, None, a, b, slice_a=sa, slice_id=1)
        finally:
            lib.set_option("tc_min_log2", 18)
            lib.set_option("tc_gather", 2)
            lib.set_option("tc_gather_min_run", 10)
        assert plan.kernel == lib.KERNEL_TC and plan.tc_pair == 2
        if g == 2:
            assert plan.tc_gather == 2 and plan.tc_run_bits == {0: 7, 1: 9, 2: 6}[variant] + (0 if variant else 0), \
                (shape, plan.tc_gather, plan.tc_run_bits)
        else:
            assert plan.tc_gather == 0
    want = np.einsum(A, la, B, lb, lc_used, optimize=True)
    for g in (2, 0):
        assert np.abs(got[g] - want).max() / np.abs(want).max() < 4e-6 * max(1, k), (shape, g)
    assert np.array_equal(got[2], got[0])             # the same values reach the same planes: bit-identical results
