"""CPU checks of the host-side pair planner (C library) through a numpy emulation
of the kernel's address algebra: the tables it emits must describe the einsum."""
import numpy as np
import pytest

import emu
from qibotn_b200 import lib


def random_case(rng, m, n, k, batch=0, nslice_a=0, nslice_b=0):
    labels = list(range(m + n + k + batch))
    M, N, K, Bt = labels[:m], labels[m:m + n], labels[m + n:m + n + k], labels[m + n + k:]
    la, lb, lc = M + K + Bt, K + N + Bt, M + N + Bt
    rng.shuffle(la); rng.shuffle(lb); rng.shuffle(lc)
    # sliced labels occupy extra address bits of the stored operand
    ra, rb = len(la) + nslice_a, len(lb) + nslice_b
    pa_all = list(rng.permutation(ra)); pb_all = list(rng.permutation(rb))
    pa, sa = pa_all[:len(la)], pa_all[len(la):]
    pb, sb = pb_all[:len(lb)], pb_all[len(lb):]
    pc = list(rng.permutation(len(lc)))
    return la, [int(x) for x in pa], [int(x) for x in sa], lb, [int(x) for x in pb], \
        [int(x) for x in sb], lc, [int(x) for x in pc]


def reference(a_flat, la, pa, sa_pos, b_flat, lb, pb, sb_pos, lc, slice_id, bits_a, bits_b):
    off_a = sum(((slice_id >> bt) & 1) << p for bt, p in zip(bits_a, sa_pos))
    off_b = sum(((slice_id >> bt) & 1) << p for bt, p in zip(bits_b, sb_pos))
    def dense(flat, labels, pos, off):
        out = np.zeros((2,) * len(labels), dtype=flat.dtype)
        for idx in (np.ndindex(*out.shape) if labels else [()]):
            out[idx] = flat[off + sum(v << p for v, p in zip(idx, pos))]
        return out
    A = dense(a_flat, la, pa, off_a)
    B = dense(b_flat, lb, pb, off_b)
    return np.einsum(A, la, B, lb, lc)


CASES = [
    # m, n, k, batch, slices a/b, C free?
    (6, 6, 3, 0, 0, 0, True), (7, 7, 5, 0, 0, 0, False), (8, 2, 2, 0, 0, 0, True),
    (3, 9, 2, 0, 0, 0, True), (1, 1, 1, 0, 0, 0, True), (0, 0, 3, 0, 0, 0, True),
    (2, 0, 1, 0, 0, 0, False), (5, 4, 6, 0, 0, 0, True), (4, 4, 2, 2, 0, 0, False),
    (7, 3, 2, 1, 2, 1, True), (1, 1, 11, 0, 0, 0, True), (2, 2, 10, 0, 1, 1, False),
    (0, 0, 12, 0, 2, 0, True), (13, 0, 1, 0, 0, 0, True), (6, 7, 0, 0, 0, 0, True),
    (3, 3, 8, 0, 0, 0, True), (4, 4, 11, 0, 0, 0, True), (3, 4, 10, 0, 1, 2, False), (4, 1, 12, 0, 0, 0, True),
    (0, 4, 10, 0, 0, 0, False),
]


@pytest.mark.parametrize("dtype", ["complex64", "complex128"])
@pytest.mark.parametrize("case", CASES)
def test_plan_tables_describe_the_einsum(case, dtype):
    m, n, k, batch, nsa, nsb, free = case
    rng = np.random.default_rng(hash(case) % 2**32)
    la, pa, sa_pos, lb, pb, sb_pos, lc, pc = random_case(rng, m, n, k, batch, nsa, nsb)
    ra, rb = len(la) + nsa, len(lb) + nsb
    a = (rng.normal(size=1 << ra) + 1j * rng.normal(size=1 << ra)).astype(dtype)
    b = (rng.normal(size=1 << rb) + 1j * rng.normal(size=1 << rb)).astype(dtype)
    bits_a = list(range(nsa)); bits_b = list(range(1, 1 + nsb))
    desc = lib.make_pair_desc(dtype, list(zip(la, pa)), list(zip(lb, pb)),
                              lc if free else list(zip(lc, pc)),
                              slice_a=list(zip(bits_a, sa_pos)), slice_b=list(zip(bits_b, sb_pos)))
    plan = lib.pair_plan(desc)
    pc_used = [desc.pos_c[i] for i in range(len(lc))]
    lc_used = [desc.label_c[i] for i in range(len(lc))]
    assert sorted(pc_used) == list(range(len(lc)))
    for slice_id in ([0] if nsa + nsb == 0 else [0, 3, 5]):
        c = np.zeros(1 << len(lc), dtype=dtype)
        emu.run_plan(plan, a, b, c, slice_id)
        got = emu.dense_from_flat(c, lc_used, pc_used)
        want = reference(a, la, pa, sa_pos, b, lb, pb, sb_pos, lc_used, slice_id, bits_a, bits_b)
        tol = 1e-4 if dtype == "complex64" else 1e-11
        assert np.allclose(got, want, rtol=tol, atol=tol * (1 << (k // 2)))
    assert plan.flops == 8.0 * 2 ** (m + n + k + batch)


def test_accumulate_and_conj_flags():
    rng = np.random.default_rng(1)
    la, pa, _, lb, pb, _, lc, pc = random_case(rng, 3, 3, 2)
    a = rng.normal(size=32) + 1j * rng.normal(size=32)
    b = rng.normal(size=32) + 1j * rng.normal(size=32)
    desc = lib.make_pair_desc("complex128", list(zip(la, pa)), list(zip(lb, pb)), list(zip(lc, pc)),
                              accumulate=True, conj_a=True)
    plan = lib.pair_plan(desc)
    c0 = rng.normal(size=64) + 0j
    c = c0.copy()
    emu.run_plan(plan, a, b, c)
    want = reference(a.conj(), la, pa, [], b, lb, pb, [], lc, 0, [], [])
    assert np.allclose(emu.dense_from_flat(c - c0, lc, pc), want)


def test_bad_descriptors_are_rejected():
    d = lib.make_pair_desc("complex64", [(0, 0), (1, 1)], [(1, 0), (2, 1)], [(0, 0), (3, 1)])
    with pytest.raises(lib.QtnError):
        lib.pair_plan(d)
    d = lib.make_pair_desc("complex64", [(0, 0), (0, 1)], [(1, 0)], [1])
    with pytest.raises(lib.QtnError):
        lib.pair_plan(d)


TC_CASES = [
    # m, n, k, slices a/b  (complex64, C free: the tensor-core kernel's domain)
    (7, 4, 3, 0, 0), (8, 5, 4, 0, 0), (7, 8, 5, 0, 0), (9, 9, 3, 1, 1), (7, 6, 9, 0, 0),
    (10, 4, 6, 2, 0), (8, 7, 4, 0, 1),
]


@pytest.mark.parametrize("case", TC_CASES)
def test_tensor_core_plan_tables(case):
    m, n, k, nsa, nsb = case
    lib.set_option("tc_min_log2", 0)
    try:
        rng = np.random.default_rng(hash(case) % 2**32)
        la, pa, sa_pos, lb, pb, sb_pos, lc, _ = random_case(rng, m, n, k, 0, nsa, nsb)
        ra, rb = len(la) + nsa, len(lb) + nsb
        a = (rng.normal(size=1 << ra) + 1j * rng.normal(size=1 << ra)).astype(np.complex64)
        b = (rng.normal(size=1 << rb) + 1j * rng.normal(size=1 << rb)).astype(np.complex64)
        bits_a = list(range(nsa)); bits_b = list(range(1, 1 + nsb))
        desc = lib.make_pair_desc("complex64", list(zip(la, pa)), list(zip(lb, pb)), lc,
                                  slice_a=list(zip(bits_a, sa_pos)), slice_b=list(zip(bits_b, sb_pos)))
        plan = lib.pair_plan(desc, 4)
    finally:
        lib.set_option("tc_min_log2", 18)
    assert plan.kernel == lib.KERNEL_TC
    pc_used = [desc.pos_c[i] for i in range(len(lc))]
    lc_used = [desc.label_c[i] for i in range(len(lc))]
    assert sorted(pc_used) == list(range(len(lc)))
    for slice_id in ([0] if nsa + nsb == 0 else [0, 3]):
        c = np.zeros(1 << len(lc), dtype=np.complex64)
        emu.run_plan(plan, a, b, c, slice_id)
        got = emu.dense_from_flat(c, lc_used, pc_used)
        want = reference(a, la, pa, sa_pos, b, lb, pb, sb_pos, lc_used, slice_id, bits_a, bits_b)
        assert np.allclose(got, want, rtol=1e-4, atol=1e-4 * (1 << (k // 2)))


def test_tensor_core_kernel_is_optional():
    la = [(i, i) for i in range(12)]          # m = 8 labels, k = 4 labels
    lb = [(8 + i, i) for i in range(10)]      # k = 4, n = 6
    lib.set_option("tc_min_log2", 0)
    try:
        assert lib.pair_plan(lib.make_pair_desc("complex64", la, lb)).kernel == lib.KERNEL_TC
        assert lib.pair_plan(lib.make_pair_desc("complex128", la, lb)).kernel == lib.KERNEL_TILE
        lib.set_option("tc_enable", 0)
        assert lib.pair_plan(lib.make_pair_desc("complex64", la, lb)).kernel == lib.KERNEL_TILE
    finally:
        lib.set_option("tc_enable", 1)
        lib.set_option("tc_min_log2", 18)


@pytest.mark.parametrize("shape", [(10, 5, 4), (11, 6, 5), (12, 7, 7), (10, 9, 4), (10, 4, 5), (8, 7, 4)])
def test_cta_pair_plans_are_well_formed(shape):
    """The CTA-pair variant (tcgen05 cta_group::2) is chosen for pre-packed tiles of 32..128 columns;
    its launch geometry must pair up: even tile count, even CTA count, half of B per stage."""
    m, n, k = shape
    la = [(i, i) for i in range(m + k)]                   # m labels then k labels
    lb = [(m + i, i) for i in range(k)] + [(m + k + j, k + j) for j in range(n)]
    lib.set_option("tc_min_log2", 0)
    try:
        plan = lib.pair_plan(lib.make_pair_desc("complex64", la, lb))
        lib.set_option("tc_pair", 0)
        single = lib.pair_plan(lib.make_pair_desc("complex64", la, lb))
    finally:
        lib.set_option("tc_min_log2", 18)
        lib.set_option("tc_pair", 5)
    assert plan.kernel == lib.KERNEL_TC and single.kernel == lib.KERNEL_TC and single.tc_pair == 0
    eligible = plan.tc_prepack == 1 and plan.tc_pad == 1 and 5 <= plan.tnb <= 7
    # paired tiles take the staggered variant (plan value 2): two half-width halves per tile
    assert plan.tc_pair == (2 if eligible else 0)
    assert (m >= 10 and n >= 5) == bool(plan.tc_pair)
    if plan.tc_pair:
        assert plan.grid % 2 == 0 and plan.tc_reserved % 2 == 0 and plan.tc_reserved > 0
        # same tiling and C layout as the single-CTA plan: the kernels are interchangeable
        assert (plan.tmb, plan.tnb, plan.tkb, plan.split_bits) == (single.tmb, single.tnb, single.tkb, single.split_bits)
        assert plan.c_elems == single.c_elems
    if plan.tc_pair == 1:
        stage = 4 * (plan.tc_plane_a + plan.tc_plane_b // 2)
        assert plan.smem_bytes == plan.tc_stages * stage + 256 <= 227 * 1024
        assert 2 <= plan.tc_stages <= 6 and plan.tc_stages >= single.tc_stages
        assert plan.workspace_bytes == single.workspace_bytes
    if plan.tc_pair == 2:
        kt = 1 << plan.tkb
        assert plan.tc_kind == (1 if plan.tkb == 4 else 0)      # scaled 3xFP16 split: 2-byte operand planes
        eb = 2 if plan.tc_kind else 4
        assert plan.tc_plane_a == 128 * kt * eb
        rank_b = 3 * (1 << plan.tnb) * kt * eb   # per CTA: 2 halves x {hi, lo} x [Br | Bi | -Br] x TN/4 rows x kt values
        stage = 4 * plan.tc_plane_a + rank_b
        raw_ring = 6 * 128 * kt * 8    # cp.async gathers of A in flight: 6 chunks of 8-byte elements
        if plan.tc_prepack_a:
            raw_ring = 0               # a pre-packed A is not gathered: the bytes go to a deeper operand ring
        assert plan.smem_bytes == plan.tc_stages * stage + raw_ring + 512 <= 227 * 1024
        assert 2 <= plan.tc_stages <= 8
        images = 1 << (plan.n_nhi + plan.n_khi)
        b_end = plan.tc_prepack_off + images * 2 * rank_b + (256 if plan.tc_kind else 0)
        # A is pre-packed too when at least four n tiles read every A tile (fp16 split only): one 16 KB image per
        # (m tile, k chunk) behind the B images, the magnitude words stay the last 256 bytes of the workspace
        assert plan.tc_prepack_a == (1 if plan.tc_kind and plan.n_nhi >= 2 else 0)
        if plan.tc_prepack_a:
            a_images = 1 << (plan.n_mhi + plan.n_khi)
            assert plan.tc_prepack_a_off == (b_end + 255) // 256 * 256
            assert plan.workspace_bytes == plan.tc_prepack_a_off + a_images * 4 * plan.tc_plane_a + 256
        else:
            assert plan.workspace_bytes == b_end
        # address-ordered gather ("tc_gather", default): bit r of the gathered element index walks the tile label
        # of rank r in address order, and the element a thread converts -- enumeration (t, i) of a_lo_pos -- sits
        # in the raw ring at the index built from the ranks of its bits: both must name the same address
        assert plan.tc_gather == 0                          # default: bulk copies only for runs of >= 8 KB (7 bits here)
        # tc_gather = 2: bulk-copied raw chunks when the tile owns the lowest address bits of A and A is not
        # pre-packed; one cp.async.bulk per contiguous run
        lib.set_option("tc_min_log2", 0)
        lib.set_option("tc_gather", 2)
        lib.set_option("tc_gather_min_run", 6)
        try:
            bplan = lib.pair_plan(lib.make_pair_desc("complex64", la, lb))
        finally:
            lib.set_option("tc_min_log2", 18)
            lib.set_option("tc_gather", 2)
            lib.set_option("tc_gather_min_run", 10)
        assert bplan.tc_gather in (0, 2) and (bplan.tc_gather == 0 or not bplan.tc_prepack_a)
        if bplan.tc_gather == 2:
            rb = bplan.tc_run_bits
            assert rb >= 6 and bplan.na_lo - rb <= 5 and list(bplan.tc_a_gbit[:rb]) == list(range(rb))
            assert rb == bplan.na_lo or bplan.tc_a_gbit[rb] != rb
        lib.set_option("tc_min_log2", 0)
        lib.set_option("tc_gather", 1)
        lib.set_option("tc_prepack_a", 0)                   # a pre-packed A is not gathered by the main kernel at all
        try:
            gplan = lib.pair_plan(lib.make_pair_desc("complex64", la, lb))
        finally:
            lib.set_option("tc_min_log2", 18)
            lib.set_option("tc_gather", 2)
            lib.set_option("tc_prepack_a", 1)
        assert gplan.tc_gather == 1 and list(gplan.a_lo_pos) == list(plan.a_lo_pos)
        plan_own, plan = plan, gplan
        n_lo = plan.na_lo
        gbit, rank = list(plan.tc_a_gbit[:n_lo]), list(plan.tc_a_rank[:n_lo])
        assert gbit == sorted(plan.a_lo_pos[:n_lo]) and sorted(rank) == list(range(n_lo))
        for e in range(0, 1 << n_lo, 37):
            t_, i_ = e & 255, e >> 8
            addr_s = sum(1 << plan.a_lo_pos[j] for j in range(8) if (t_ >> j) & 1) | plan.tc_a_it_g[i_]
            raw = sum(1 << rank[j] for j in range(8) if (t_ >> j) & 1) + plan.tc_a_it_raw[i_]
            tg, ig = raw & 255, raw >> 8
            addr_g = sum(1 << gbit[j] for j in range(8) if (tg >> j) & 1) | plan.tc_a_it_gg[ig]
            assert addr_s == addr_g
        plan = plan_own
        if plan.tc_kind:
            # the lowest k label is the first iteration bit (a thread packs two k-adjacent fp16 values per
            # store) and the five lane bits own the bank-selecting word offsets
            assert plan.a_lo_sstr[8] == 1 and sorted(plan.a_lo_sstr[:5]) == [2, 4, 8, 16, 32]
            assert all(plan.tc_a_it_s[2 * q + 1] == plan.tc_a_it_s[2 * q] + 1 for q in range(4))
        lib.set_option("tc_min_log2", 0)
        lib.set_option("tc_stagger", 0)
        try:
            old = lib.pair_plan(lib.make_pair_desc("complex64", la, lb))
        finally:
            lib.set_option("tc_min_log2", 18)
            lib.set_option("tc_stagger", 1)
        assert old.tc_pair == 1 and old.tc_prepack_off == plan.tc_prepack_off


def test_options_per_call_do_not_touch_process_state():
    """``qtn_pair_plan_opts``: the switches travel with the call (a ``qtn_options`` struct); the process-wide
    copy that ``qtn_set_option`` edits is neither read nor written."""
    import ctypes as C

    la = [(i, i) for i in range(12 + 7)]
    lb = [(12 + i, i) for i in range(7)] + [(19 + j, 7 + j) for j in range(7)]
    before = {k: lib.get_option(k) for k in ("tc_enable", "tc_min_log2", "tc_pair", "tc_stagger", "tc_fp16")}
    assert lib.load().qtn_sizeof(b"qtn_options") == C.sizeof(lib.Options)
    plans = {}
    for tag, opt in (("default", lib.options(tc_min_log2=0)), ("tf32", lib.options(tc_min_log2=0, tc_fp16=0)),
                     ("pair1", lib.options(tc_min_log2=0, tc_stagger=0)), ("ffma", lib.options(tc_enable=0))):
        plans[tag] = lib.pair_plan(lib.make_pair_desc("complex64", la, lb), options=opt)
    assert (plans["default"].kernel, plans["default"].tc_pair, plans["default"].tc_kind) == (lib.KERNEL_TC, 2, 1)
    assert (plans["tf32"].tc_pair, plans["tf32"].tc_kind) == (2, 0)
    assert (plans["pair1"].tc_pair, plans["pair1"].tc_kind) == (1, 0)
    assert plans["ffma"].kernel == lib.KERNEL_TILE
    assert before == {k: lib.get_option(k) for k in before}
    with pytest.raises(lib.QtnError):
        lib.options(no_such_switch=1)


BATCH_CASES = [
    # m, n, k, batch, slices a/b: batch (hyper) labels on the tensor-core kernel (staggered CTA pairs only) and on
    # the streaming K reduction
    ("tc", 10, 5, 3, 2, 0, 0), ("tc", 10, 5, 4, 1, 1, 0), ("tc", 11, 6, 5, 2, 0, 1), ("tc", 10, 7, 4, 3, 0, 0),
    ("kred", 2, 3, 10, 2, 0, 0), ("kred", 4, 4, 10, 1, 1, 1), ("kred", 0, 1, 11, 3, 0, 0),
    ("kred", 5, 3, 10, 2, 0, 0), ("kred", 5, 5, 10, 0, 1, 0), ("kred", 3, 5, 11, 1, 0, 1),
    ("kred", 4, 4, 10, 4, 0, 0), ("kred", 3, 2, 11, 5, 0, 0), ("kred", 1, 1, 12, 3, 1, 0),
    ("tile", 9, 5, 4, 2, 0, 0), ("tc", 10, 4, 4, 1, 0, 0), ("tc", 11, 4, 3, 2, 1, 0), ("tile", 10, 3, 4, 1, 0, 0),
]


@pytest.mark.parametrize("case", BATCH_CASES)
def test_batch_labels_in_tensor_core_and_kred_plans(case):
    kind, m, n, k, batch, nsa, nsb = case
    lib.set_option("tc_min_log2", 0)
    try:
        rng = np.random.default_rng(abs(hash(case)) % 2**32)
        la, pa, sa_pos, lb, pb, sb_pos, lc, _ = random_case(rng, m, n, k, batch, nsa, nsb)
        ra, rb = len(la) + nsa, len(lb) + nsb
        a = (rng.normal(size=1 << ra) + 1j * rng.normal(size=1 << ra)).astype(np.complex64)
        b = (rng.normal(size=1 << rb) + 1j * rng.normal(size=1 << rb)).astype(np.complex64)
        bits_a = list(range(nsa)); bits_b = list(range(1, 1 + nsb))
        desc = lib.make_pair_desc("complex64", list(zip(la, pa)), list(zip(lb, pb)), lc,
                                  slice_a=list(zip(bits_a, sa_pos)), slice_b=list(zip(bits_b, sb_pos)))
        plan = lib.pair_plan(desc, 4)
    finally:
        lib.set_option("tc_min_log2", 18)
    if kind == "kred":
        # batch labels at low address bits (below bit 6, at most three) are INNER labels of the K reduction; the others
        # and a fifth m / n label ride on grid.y
        assert 0 <= plan.kred_inner <= min(3, batch)
        assert plan.n_batch == batch - plan.kred_inner + max(0, m - 4) + max(0, n - 4)
    else:
        assert plan.n_batch == batch
    if kind == "tc":
        # m >= 10 (8 m tiles per B image), n >= 5: the staggered pair kernel; fp16 split when 16 k values fill a stage
        assert plan.kernel == lib.KERNEL_TC and plan.tc_pair == 2 and plan.tc_prepack == 1
        assert plan.tc_kind == (1 if k >= 4 else 0)
        assert plan.n_mhi == m - 7 + batch and plan.grid == (1 << (plan.n_mhi + plan.n_nhi + plan.split_bits))
        assert plan.n_real == min(n, 7) and plan.tnb == max(min(n, 7), 5)      # 16 columns: a phantom fifth column bit
    elif kind == "kred":
        assert plan.kernel == lib.KERNEL_KRED and plan.grid >= 1
    else:
        assert plan.kernel == lib.KERNEL_TILE          # too few m tiles (or columns) for the pair kernel
    pc_used = [desc.pos_c[i] for i in range(len(lc))]
    lc_used = [desc.label_c[i] for i in range(len(lc))]
    assert sorted(pc_used) == list(range(len(lc)))
    for slice_id in ([0] if nsa + nsb == 0 else [0, 3]):
        c = np.zeros(1 << len(lc), dtype=np.complex64)
        emu.run_plan(plan, a, b, c, slice_id)
        got = emu.dense_from_flat(c, lc_used, pc_used)
        want = reference(a, la, pa, sa_pos, b, lb, pb, sb_pos, lc_used, slice_id, bits_a, bits_b)
        assert np.allclose(got, want, rtol=1e-4, atol=1e-4 * (1 << (k // 2)))
