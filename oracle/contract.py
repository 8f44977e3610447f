"""Pairwise ``numpy.tensordot`` contraction of an interleaved operand list
(oracle; test infrastructure only).

This is the arithmetic the reference's CPU path ends in: quimb hands a
contraction tree to cotengra, which executes each node with
``tensordot``/``einsum`` on numpy arrays (quimb 1.13.0, cotengra 0.7.5 -- not
vendored; SURVEY.md section 3.6, ``/root/reference/src/qibotn/eval_qu.py:43-44``).
The reference fixes no contraction order (cotengra's search is randomised), so
the order is an input here: either an SSA path supplied by the caller (the
same tree the GPU executes, for the CPU baseline) or a plain greedy one.
Results are order-independent up to rounding, which ``tests/`` checks.
"""

import numpy as np


def split_interleaved(inter):
    ops = list(inter[0:-1:2])
    modes = [list(m) for m in inter[1:-1:2]]
    out = list(inter[-1])
    return ops, modes, out


def _pair(a, ma, b, mb, keep):
    """Contract two tensors over every label they share that is not in ``keep``;
    labels in ``keep`` that appear in both are batch (hyper) labels."""
    sa, sb = set(ma), set(mb)
    shared = [m for m in ma if m in sb]
    summed = [m for m in shared if m not in keep]
    batch = [m for m in shared if m in keep]
    if not batch:
        c = np.tensordot(a, b, axes=([ma.index(m) for m in summed], [mb.index(m) for m in summed]))
        mc = [m for m in ma if m not in summed] + [m for m in mb if m not in summed]
        return c, mc
    # rare (hyper-index) case: fall back to einsum with local letters
    labels = {m: i for i, m in enumerate(dict.fromkeys(ma + mb))}
    mc = [m for m in ma if m not in summed] + [m for m in mb if m not in sa]
    c = np.einsum(a, [labels[m] for m in ma], b, [labels[m] for m in mb], [labels[m] for m in mc])
    return c, mc


def greedy_ssa_path(modes, out):
    """Smallest-output-first greedy order (deterministic); SSA ids."""
    live = {i: set(m) for i, m in enumerate(modes)}
    count = {}
    for m in modes:
        for x in m:
            count[x] = count.get(x, 0) + 1
    for x in out:
        count[x] = count.get(x, 0) + 1
    nxt, path = len(modes), []
    while len(live) > 1:
        best = None
        ids = sorted(live)
        for ii, i in enumerate(ids):
            for j in ids[ii + 1:]:
                sh = live[i] & live[j]
                if not sh and len(live) > 2 and any(live[i] & live[k] for k in ids if k not in (i, j)):
                    continue
                gone = {x for x in sh if count[x] == 2}
                size = len(live[i] | live[j]) - len(gone)
                key = (size - max(len(live[i]), len(live[j])), size, i, j)
                if best is None or key < best[0]:
                    best = (key, i, j, gone)
        _, i, j, gone = best
        shared = live[i] & live[j]
        new = (live.pop(i) | live.pop(j)) - gone
        for x in shared:
            count[x] -= 2 if x in gone else 1   # a surviving shared label loses one owner
        live[nxt] = new
        path.append((i, j))
        nxt += 1
    return path


def contract_ssa(ops, modes, out, path, sliced=None, slice_id=0, dtype=None):
    """Run an SSA ``path`` (pairs of tensor ids; new tensors get the next id).

    ``sliced`` is an ordered list of labels fixed for this slice: bit ``k`` of
    ``slice_id`` (most significant first) is the value of ``sliced[k]``.
    """
    ts = {}
    fixed = {}
    if sliced:
        for k, lab in enumerate(sliced):
            fixed[lab] = (slice_id >> (len(sliced) - 1 - k)) & 1
    for i, (t, m) in enumerate(zip(ops, modes)):
        t = np.asarray(t if dtype is None else np.asarray(t).astype(dtype))
        m = list(m)
        if fixed:
            idx = tuple(fixed[x] if x in fixed else slice(None) for x in m)
            t = t[idx]
            m = [x for x in m if x not in fixed]
        ts[i] = (t, m)
    # how many live tensors (plus the output) carry each label
    cnt = {}
    for _, m in ts.values():
        for x in m:
            cnt[x] = cnt.get(x, 0) + 1
    outset = set(out) - set(fixed)
    nxt = len(ops)
    for i, j in path:
        a, ma = ts.pop(i)
        b, mb = ts.pop(j)
        keep = {x for x in set(ma) & set(mb) if cnt[x] > 2 or x in outset}
        c, mc = _pair(a, ma, b, mb, keep)
        for x in set(ma) & set(mb):
            cnt[x] -= 1 if x in keep else 2
        ts[nxt] = (c, mc)
        nxt += 1
    (res, mres), = ts.values()
    want = [x for x in out if x not in fixed]
    if mres != want:
        res = np.transpose(res, [mres.index(x) for x in want])
    return res


def contract(inter, path=None, dtype=None):
    """Contract a whole interleaved list; greedy order unless ``path`` is given."""
    ops, modes, out = split_interleaved(inter)
    if path is None:
        path = greedy_ssa_path(modes, out)
    return contract_ssa(ops, modes, out, path, dtype=dtype)


def contract_sliced(inter, path, sliced, slice_ids, dtype=None):
    """Sum of the slices ``slice_ids`` -- what ``Network.contract(slices=range)``
    returns on one rank (``/root/reference/src/qibotn/eval.py:265,297``)."""
    ops, modes, out = split_interleaved(inter)
    total = None
    for s in slice_ids:
        r = contract_ssa(ops, modes, out, path, sliced=sliced, slice_id=s, dtype=dtype)
        total = r if total is None else total + r
    return total
