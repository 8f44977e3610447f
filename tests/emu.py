"""Numpy emulation of the *address algebra* of the GETT kernels, driven by the
launch-ready ``qtn_pair_plan_t`` the C library produces.

Test infrastructure only: it lets CPU-only runs check that the host planner's
tables (tile membership, enumeration orders, staging indices, split-K, slices)
describe the right contraction, without a GPU.  It mirrors the kernel loop
structure on purpose (tile -> k_hi loop -> staged epilogue) but computes the
tile product with numpy.
"""

import numpy as np


def dep(idx, pos, n):
    off = 0
    for i in range(n):
        if (idx >> i) & 1:
            off |= 1 << pos[i]
    return off


def dep_present(idx, pos, n):
    """``dep`` for label lists where position 255 marks a label the operand does not carry."""
    return dep(idx, [p if p != 255 else 0 for p in pos], n) if all(p != 255 for p in pos[:n]) else \
        sum(((idx >> i) & 1) << pos[i] for i in range(n) if pos[i] != 255)


def slice_off(slice_id, n, bits, pos):
    off = 0
    for i in range(n):
        if (slice_id >> bits[i]) & 1:
            off |= 1 << pos[i]
    return off


def flat_from_dense(arr, labels, pos):
    """Dense numpy array with axes ``labels`` -> flat address-indexed vector for layout ``pos``."""
    r = len(labels)
    size = 1 << (max(pos) + 1) if r else 1
    flat = np.zeros(size, dtype=arr.dtype)
    it = np.ndindex(*arr.shape) if r else [()]
    for idx in it:
        off = sum(v << p for v, p in zip(idx, pos))
        flat[off] = arr[idx]
    return flat


def dense_from_flat(flat, labels, pos):
    r = len(labels)
    out = np.zeros((2,) * r, dtype=flat.dtype)
    it = np.ndindex(*out.shape) if r else [()]
    for idx in it:
        off = sum(v << p for v, p in zip(idx, pos))
        out[idx] = flat[off]
    return out


def run_plan(P, a, b, c, slice_id=0):
    """Emulate qtn_pair_run on flat address-indexed numpy vectors (c is updated in place)."""
    if P.swapped:
        a, b = b, a
    if P.conj_a:
        a = a.conj()
    if P.conj_b:
        b = b.conj()
    sa = slice_off(slice_id, P.n_slice_a, P.slice_bit_a, P.slice_pos_a)
    sb = slice_off(slice_id, P.n_slice_b, P.slice_bit_b, P.slice_pos_b)
    if P.kernel == 1:
        return _run_reduce(P, a, b, c, sa, sb)
    if P.kernel == 2:
        return _run_tc(P, a, b, c, sa, sb)
    if P.kernel == 3:
        return _run_kred(P, a, b, c, sa, sb)
    TM, TN, TK = 1 << P.tmb, 1 << P.tnb, 1 << P.tkb
    nsplit = 1 << P.split_bits
    ws = np.zeros((nsplit, P.c_elems), dtype=c.dtype)
    kbits = P.n_khi - P.split_bits
    for cta in range(P.grid):
        split = cta & (nsplit - 1)
        rest = cta >> P.split_bits
        mh = rest & ((1 << P.n_mhi) - 1)
        rest >>= P.n_mhi
        nh = rest & ((1 << P.n_nhi) - 1)
        bt = rest >> P.n_nhi
        a_base = dep(mh, P.a_mhi, P.n_mhi) | dep(bt, P.a_bat, P.n_batch) | sa
        b_base = dep(nh, P.b_nhi, P.n_nhi) | dep(bt, P.b_bat, P.n_batch) | sb
        c_base = dep(mh, P.c_mhi, P.n_mhi) | dep(nh, P.c_nhi, P.n_nhi) | dep(bt, P.c_bat, P.n_batch)
        acc = np.zeros((TM, TN), dtype=c.dtype)
        for kk in range(1 << kbits):
            kh = (split << kbits) | kk
            a_k, b_k = dep(kh, P.a_khi, P.n_khi), dep(kh, P.b_khi, P.n_khi)
            s_a = np.zeros(TK * P.sa_stride, dtype=c.dtype)
            s_b = np.zeros(TK * P.sb_stride, dtype=c.dtype)
            for (n_lo, lo_pos, lo_sstr, smem, src, base) in (
                    (P.na_lo, P.a_lo_pos, P.a_lo_sstr, s_a, a, a_base + a_k),
                    (P.nb_lo, P.b_lo_pos, P.b_lo_sstr, s_b, b, b_base + b_k)):
                for e in range(1 << n_lo):
                    g, s, ph = 0, 0, False
                    for j in range(n_lo):
                        if (e >> j) & 1:
                            if lo_pos[j] == 255:
                                ph = True
                            else:
                                g |= 1 << lo_pos[j]
                            s += lo_sstr[j]
                    smem[s] = 0 if ph else src[base + g]
            am = s_a.reshape(TK, P.sa_stride)[:, :TM]
            bm = s_b.reshape(TK, P.sb_stride)[:, :TN]
            acc += am.T @ bm
        count = 1 << P.nc_lo
        stage = np.zeros(4096, dtype=c.dtype)
        for m in range(TM):
            sm = sum(P.c_m_sidx[bb] for bb in range(P.tmb) if (m >> bb) & 1)
            for n in range(TN):
                sn = sum(P.c_n_sidx[bb] for bb in range(P.tnb) if (n >> bb) & 1)
                if sm + sn < count:
                    stage[sm + sn] = acc[m, n]
        for e in range(count):
            g = c_base + dep(e, P.c_lo_pos, P.nc_lo)
            if P.split_bits:
                ws[split, g] = stage[e]
            elif P.accumulate:
                c[g] += stage[e]
            else:
                c[g] = stage[e]
    if P.split_bits:
        tot = ws.sum(0)
        if P.accumulate:
            c[: P.c_elems] += tot
        else:
            c[: P.c_elems] = tot
    return c


def _run_reduce(P, a, b, c, sa, sb):
    M, N = 1 << P.n_mhi, 1 << P.n_nhi
    acc = np.zeros((M, N), dtype=c.dtype)
    for k in range(1 << P.n_khi):
        a_k = sa | dep(k, P.a_khi, P.n_khi)
        b_k = sb | dep(k, P.b_khi, P.n_khi)
        av = np.array([a[a_k | dep(m, P.a_mhi, P.n_mhi)] for m in range(M)])
        bv = np.array([b[b_k | dep(n, P.b_nhi, P.n_nhi)] for n in range(N)])
        acc += np.outer(av, bv)
    for m in range(M):
        for n in range(N):
            g = dep(m, P.c_mhi, P.n_mhi) | dep(n, P.c_nhi, P.n_nhi)
            if P.accumulate:
                c[g] += acc[m, n]
            else:
                c[g] = acc[m, n]
    return c


def _canonical_kmajor(plane, rows, kt):
    """Read a ``rows x kt`` operand out of the tensor core's K-major no-swizzle layout
    (8-row x 4-value core matrices of 32 floats; K-neighbours 32 floats apart, 8-row groups
    ``kt * 8`` floats apart) -- written independently of the planner's stride tables."""
    out = np.zeros((rows, kt), dtype=plane.dtype)
    for r in range(rows):
        for k in range(kt):
            out[r, k] = plane[(r & 7) * 4 + (k & 3) + (r >> 3) * (kt * 8) + (k >> 2) * 32]
    return out


def _canonical_kmajor16(plane, rows, kt):
    """The fp16 planes of the scaled 3xFP16 split: 8-row x 8-value core matrices of 64 halves; K-neighbours
    64 halves apart, 8-row groups ``kt * 8`` halves apart."""
    out = np.zeros((rows, kt), dtype=plane.dtype)
    for r in range(rows):
        for k in range(kt):
            out[r, k] = plane[(r & 7) * 8 + (k & 7) + (r >> 3) * (kt * 8) + (k >> 3) * 64]
    return out


def _run_tc(P, a, b, c, sa, sb):
    """Emulate gett_tc.cu: gather into canonical-layout planes, tile product, fixed C tile layout."""
    TN, KT = 1 << P.tnb, 1 << P.tkb
    assert P.tmb == 7 and not P.accumulate
    f16 = P.tc_kind == 1                       # scaled 3xFP16 split: 2-byte planes, 8 x 8 core matrices
    eb = 2 if f16 else 4
    assert P.tc_plane_a == 128 * KT * eb and P.tc_plane_b == TN * KT * eb
    if P.tc_pair != 2:
        assert P.n_batch == 0                  # batch labels ride on the staggered CTA-pair kernel only
        assert P.smem_bytes >= P.tc_stages * 4 * (P.tc_plane_a + P.tc_plane_b) + 8 * (2 * P.tc_stages + 2)
    assert P.smem_bytes <= 227 * 1024 and 1 <= P.tc_stages <= 8
    # the epilogue hard-codes C's tile layout: m bits 0-4 | n bits | m bits 5-6; a phantom column bit (16-column batch
    # nodes on the 32-column staggered tile: n_real = tnb - 1) has no position in C
    nr = P.n_real
    assert nr in (P.tnb, P.tnb - 1) and (nr == P.tnb or (P.n_batch > 0 and P.tc_pair == 2))
    want = list(range(5)) + [5 + nr, 6 + nr] + [5 + j for j in range(nr)] + [255] * (P.tnb - nr)
    assert [P.c_lo_pos[i] for i in range(7 + P.tnb)] == want
    nsplit = 1 << P.split_bits
    ws = np.zeros((nsplit, P.c_elems), dtype=c.dtype)
    kbits = P.n_khi - P.split_bits
    for cta in range(P.grid):
        split = cta & (nsplit - 1)
        rest = cta >> P.split_bits
        nh = rest & ((1 << P.n_nhi) - 1)
        mh = rest >> P.n_nhi
        # batch labels = the top n_batch bits of the m-tile index (A, C) and their own deposits in B
        bt = mh >> (P.n_mhi - P.n_batch) if P.n_batch else 0
        a_base = dep(mh, P.a_mhi, P.n_mhi) | sa
        b_base = dep(nh, P.b_nhi, P.n_nhi) | dep(bt, P.b_bat, P.n_batch) | sb
        c_base = dep(mh, P.c_mhi, P.n_mhi) | dep(nh, P.c_nhi, P.n_nhi)
        acc = np.zeros((128, TN), dtype=np.complex128)
        for kk in range(1 << kbits):
            kh = (split << kbits) | kk
            a_k, b_k = dep(kh, P.a_khi, P.n_khi), dep(kh, P.b_khi, P.n_khi)
            planes = []
            for (n_lo, lo_pos, lo_sstr, size, src, base) in (
                    (P.na_lo, P.a_lo_pos, P.a_lo_sstr, 128 * KT, a, a_base | a_k),
                    (P.nb_lo, P.b_lo_pos, P.b_lo_sstr, TN * KT, b, b_base | b_k)):
                assert (1 << n_lo) == size
                # the five lane bits own the five bank-selecting offsets (1, 2, 4, 8, 16 floats)
                if not f16:
                    assert sorted(lo_sstr[j] for j in range(5)) == [1, 2, 4, 8, 16]
                plane = np.full(size, np.nan, dtype=np.complex128)
                for e in range(size):
                    g = s = 0
                    ph = False
                    for j in range(n_lo):
                        if (e >> j) & 1:
                            if lo_pos[j] == 255:
                                ph = True                # phantom column bit: rows of zeros
                            else:
                                g |= 1 << lo_pos[j]
                            s += lo_sstr[j]
                    assert np.isnan(plane[s])           # every slot written exactly once
                    plane[s] = 0.0 if ph else src[base | g]
                planes.append(plane)
            cm = _canonical_kmajor16 if f16 else _canonical_kmajor
            acc += cm(planes[0], 128, KT) @ cm(planes[1], TN, KT).T
        out = ws[split] if P.split_bits else c
        for r in range(128):
            for n in range(1 << nr):                   # phantom columns are never stored
                out[c_base + (r & 31) + (n << 5) + ((r >> 5) << (5 + nr))] = acc[r, n]
    if P.split_bits:
        c[: P.c_elems] = ws.sum(0)
    return c


def _run_kred(P, a, b, c, sa, sb):
    """Emulate gett_kred_kernel: per CTA a range of k chunks, 16 x 16 (padded) outputs per inner batch value, one
    partial per CTA summed at the end.  A chunk row is (k value, inner batch value): row = (k << q) + bl."""
    KC = 1 << P.tkb
    QI = P.kred_inner
    BL, ROWS = 1 << QI, KC << QI
    assert P.tmb == 4 and P.tnb == 4 and 0 <= P.n_batch <= 15 and P.n_mhi == 0 and P.n_nhi == 0 and 0 <= QI <= 3
    ib = min(3, P.n_khi)
    ngroups = 1 << (P.n_khi - ib)
    assert P.smem_bytes >= 2 * 2 * ROWS * 16 * c.itemsize and 1 <= P.grid <= ngroups and P.smem_bytes <= 80 * 1024
    assert P.na_lo <= 11 and P.nb_lo <= 11
    ws = np.zeros((P.grid, P.c_elems), dtype=np.complex128)
    for cta, bt in ((x, y) for x in range(P.grid) for y in range(1 << P.n_batch)):      # grid.y = batch value
        acc = np.zeros((BL, 16, 16), dtype=np.complex128)
        # groups of 2^ib consecutive chunks; neighbouring CTAs take neighbouring groups
        for kh in (g << ib | i for g in range(cta, ngroups, P.grid) for i in range(1 << ib)):
            a_k = sa | dep(kh, P.a_khi, P.n_khi) | dep_present(bt, P.a_bat, P.n_batch)
            b_k = sb | dep(kh, P.b_khi, P.n_khi) | dep_present(bt, P.b_bat, P.n_batch)
            chunks = []
            for (n_lo, lo_pos, lo_sstr, src, base) in ((P.na_lo, P.a_lo_pos, P.a_lo_sstr, a, a_k),
                                                       (P.nb_lo, P.b_lo_pos, P.b_lo_sstr, b, b_k)):
                buf = np.zeros(ROWS * 16, dtype=np.complex128)
                for e in range(1 << n_lo):
                    g = s = 0
                    for j in range(n_lo):
                        if (e >> j) & 1:
                            g |= 1 << lo_pos[j]
                            s += lo_sstr[j]
                    buf[s] = src[base | g]
                chunks.append(buf.reshape(ROWS, 16))
            for bl in range(BL):
                acc[bl] += chunks[0][bl::BL].T @ chunks[1][bl::BL]
        for bl in range(BL):
            for m in range(1 << P.m_real):
                for n in range(1 << P.n_real):
                    off = (dep(m, P.c_mhi, P.m_real) | dep(n, P.c_nhi, P.n_real) | dep(bt, P.c_bat, P.n_batch) |
                           dep(bl, P.kred_c_in, QI))
                    ws[cta, off] = acc[bl, m, n]
    tot = ws.sum(0)
    if P.accumulate:
        c[: P.c_elems] += tot
    else:
        c[: P.c_elems] = tot
    return c
