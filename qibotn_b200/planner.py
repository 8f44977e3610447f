"""Contraction-path and slicing planner.

Replaces what the reference delegates to cuTensorNet's
``Network.contract_path(optimize={"samples": n, "slicing": {...}})``
(``/root/reference/src/qibotn/eval.py:180-190``) and to cotengra on the CPU
path.  Everything here is symbolic (label sets and extents-2 sizes); no tensor
data is touched.

A network is a list of label tuples (one per tensor) plus the output labels.
Label sets are Python ints used as bit masks, so unions/intersections and
``bit_count`` are single operations.  A plan is an *SSA path*: pairs of tensor
ids, leaves being ``0..n-1`` and the k-th contraction creating id ``n+k``.

Search = a deterministic rank-<=2 absorption prologue, then the better of
(a) Boltzmann-randomised greedy and (b) recursive spectral bisection with
FM refinement, over ``samples`` seeds; then greedy index slicing down to a
memory target, with at least ``min_slices`` slices when asked
(``eval.py:186``: ``max(32, size)`` in the distributed drivers).
Cost model: one complex MAC per point of the union index space of each
pairwise contraction (8 real flops), the convention of SURVEY.md section 8(d).
"""

from __future__ import annotations

import heapq
import math
import random
from dataclasses import dataclass, field

import numpy as np

__all__ = ["SymbolicNetwork", "PathInfo", "find_path", "path_cost", "estimate_seconds", "compute_slices"]


class SymbolicNetwork:
    """Label structure of a tensor network whose every extent is 2."""

    def __init__(self, modes, out):
        labels = {}
        for m in modes:
            for x in m:
                labels.setdefault(x, len(labels))
        for x in out:
            labels.setdefault(x, len(labels))
        self.label_bit = labels                      # user label -> bit index
        self.bit_label = {b: l for l, b in labels.items()}
        self.masks = [sum(1 << labels[x] for x in m) for m in modes]
        self.out_mask = sum(1 << labels[x] for x in out)
        self.n = len(modes)
        cnt = {}
        for m in self.masks:
            mm = m
            while mm:
                low = mm & -mm
                cnt[low] = cnt.get(low, 0) + 1
                mm ^= low
        # labels owned by more than two tensors (or two + the output) are hyper labels
        self.hyper_mask = 0
        for low, c in cnt.items():
            if c + (1 if low & self.out_mask else 0) > 2:
                self.hyper_mask |= low
        self.count = cnt

    def to_mask(self, labels):
        return sum(1 << self.label_bit[x] for x in labels)

    def to_labels(self, mask):
        out, b = [], 0
        while mask:
            if mask & 1:
                out.append(self.bit_label[b])
            mask >>= 1
            b += 1
        return out


@dataclass
class PathInfo:
    """What ``contract_path`` returns in the reference (``info.path``,
    ``info.slices``, ``info.num_slices``, ``info.opt_cost``), plus our accounting."""

    path: list                     # SSA pairs
    sliced_labels: list            # user labels, most significant slice bit first
    num_slices: int
    opt_cost: float                # complex MACs, all slices, hoisted work counted once
    flops: float                   # 8 * opt_cost
    log2_width: int                # largest intermediate (elements), per slice
    per_slice_cost: float          # MACs of the slice-dependent part, one slice
    hoisted_cost: float            # MACs computed once
    seed: int = 0
    method: str = ""
    extra: dict = field(default_factory=dict)

    # the reference reads ``info.slices`` and passes it back as ``slicing=``
    @property
    def slices(self):
        return self.sliced_labels

    def to_json(self):
        import json

        return json.dumps({
            "path": [list(p) for p in self.path], "sliced_labels": list(self.sliced_labels),
            "num_slices": self.num_slices, "opt_cost": self.opt_cost, "flops": self.flops,
            "log2_width": self.log2_width, "per_slice_cost": self.per_slice_cost,
            "hoisted_cost": self.hoisted_cost, "seed": self.seed, "method": self.method,
            "extra": self.extra})

    @classmethod
    def from_json(cls, text):
        import json

        d = json.loads(text)
        d["path"] = [tuple(p) for p in d["path"]]
        return cls(**d)


# ------------------------------------------------------------------ cost model
def _walk(net: SymbolicNetwork, path, sliced_mask=0):
    """Yield (id, a_mask, b_mask, c_mask, depends_on_slice) per contraction."""
    masks = {i: m & ~sliced_mask for i, m in enumerate(net.masks)}
    dep = {i: bool(m & sliced_mask) for i, m in enumerate(net.masks)}
    out = net.out_mask & ~sliced_mask
    hyper = net.hyper_mask
    remaining = None
    if hyper:
        remaining = dict(net.count)
    nxt = net.n
    for i, j in path:
        a, b = masks.pop(i), masks.pop(j)
        shared = a & b
        if hyper & shared:
            keep = out & shared
            hs = hyper & shared
            while hs:
                low = hs & -hs
                remaining[low] -= 1
                if remaining[low] + (1 if low & net.out_mask else 0) > 1:
                    keep |= low
                hs ^= low
            c = (a | b) & ~(shared & ~keep)
        else:
            c = (a ^ b) | (shared & out)
        d = dep.pop(i) | dep.pop(j)
        masks[nxt], dep[nxt] = c, d
        yield nxt, a, b, c, d
        nxt += 1


def path_cost(net: SymbolicNetwork, path, sliced_mask=0):
    """(total MACs over all slices with hoisting, log2 width, per-slice MACs, hoisted MACs)."""
    per_slice = hoisted = 0.0
    width = 0
    for _, a, b, c, d in _walk(net, path, sliced_mask):
        u = (a | b).bit_count()
        if d:
            per_slice += 2.0 ** u
        else:
            hoisted += 2.0 ** u
        width = max(width, c.bit_count(), a.bit_count(), b.bit_count())
    ns = 2.0 ** sliced_mask.bit_count()
    return per_slice * ns + hoisted, width, per_slice, hoisted


def estimate_seconds(net: SymbolicNetwork, path, sliced_mask=0, dtype="complex64"):
    """Rough device time of a plan, (seconds of one slice, seconds of the hoisted part): every pairwise contraction
    costs max(flops / rate, bytes / bandwidth) + a launch, with the rates the kernels reach on a B200
    (`profiles/`: tcgen05 kernel ~280 TFLOP/s or ~4.2 TB/s, FFMA tile kernel ~25 TFLOP/s or ~1.8 TB/s, K reduction
    ~2.5 TB/s; complex128 on the DFMA kernels).  Symbolic, no lowering: used to cut many-term observables into
    equal TIME shares (eval._expectation_distributed) -- flops alone misjudge light cones whose nodes are
    memory-bound."""
    esz = 8.0 if str(dtype) == "complex64" else 16.0
    c64 = esz == 8.0
    per_slice = hoisted = 0.0
    for _, a, b, c, d in _walk(net, path, sliced_mask):
        ra, rb, rc = a.bit_count(), b.bit_count(), c.bit_count()
        batch = (a & b & c).bit_count()
        k = (a & b & ~c).bit_count()
        m, n = ra - k - batch, rb - k - batch
        if n > m:
            m, n = n, m
        flops = 8.0 * 2.0 ** (m + n + k + batch)
        nbytes = esz * (2.0 ** ra + 2.0 ** rb + 2.0 ** rc)
        if c64 and m >= (10 if batch else 7) and n >= (5 if batch else 4) and k >= 3 and m + n + k >= 18:
            t = max(flops / 280e12, nbytes / 4.2e12)
        elif m <= 4 and n <= 4 and k >= 10 and batch <= 15:
            t = nbytes / 2.5e12
        else:
            t = max(flops / (25e12 if c64 else 8e12), nbytes / (1.8e12 if c64 else 1.0e12))
        t += 4e-6
        if d:
            per_slice += t
        else:
            hoisted += t
    return per_slice, hoisted


# --------------------------------------------------------------- simplification
def _absorb_small(net: SymbolicNetwork):
    """Deterministic prologue: contract any tensor into a neighbour whenever the
    result is no larger than the bigger of the two (rank-1/2 absorption and
    parallel edges).  Returns (ssa pairs, live {id: mask}, next id)."""
    live = {i: m for i, m in enumerate(net.masks)}
    out = net.out_mask
    owners = {}
    for i, m in live.items():
        mm = m
        while mm:
            low = mm & -mm
            owners.setdefault(low, set()).add(i)
            mm ^= low
    path, nxt = [], net.n
    hyper = net.hyper_mask
    changed = True
    while changed and len(live) > 1:
        changed = False
        for i in sorted(live, key=lambda t: (live[t].bit_count(), t)):
            if i not in live:
                continue
            a = live[i]
            best = None
            mm = a
            seen = set()
            while mm:
                low = mm & -mm
                mm ^= low
                for j in owners.get(low, ()):
                    if j == i or j in seen or j not in live:
                        continue
                    seen.add(j)
                    b = live[j]
                    shared = a & b
                    if shared & hyper:
                        continue
                    c = (a ^ b) | (shared & out)
                    if c.bit_count() <= max(a.bit_count(), b.bit_count()):
                        key = (c.bit_count(), j)
                        if best is None or key < best[0]:
                            best = (key, j, c)
            if best is None:
                continue
            _, j, c = best
            for t in (i, j):
                mm = live[t]
                while mm:
                    low = mm & -mm
                    owners[low].discard(t)
                    mm ^= low
                del live[t]
            live[nxt] = c
            mm = c
            while mm:
                low = mm & -mm
                owners.setdefault(low, set()).add(nxt)
                mm ^= low
            path.append((i, j))
            nxt += 1
            changed = True
    return path, live, nxt


# ------------------------------------------------------- variable elimination
def _elimination_path(live, out, hyper, nxt, rng, temperature=0.0, order=None):
    """Bucket (variable) elimination for networks with hyper labels: repeatedly pick the label whose
    elimination -- contracting every tensor that owns it -- yields the smallest tensor (ties and, with
    ``temperature`` > 0, near-ties broken at random), and contract its owners pairwise, smallest first.
    For circuits whose entangling gates are diagonal (QAOA) the labels are the vertices of a sparse
    "line graph" and this is the classic min-width heuristic on it; pairwise greedy on the tensors,
    which scores one merge at a time, does far worse there.  ``order`` (label bits, as single-bit masks)
    replaces the greedy choice by a given elimination order (``elimorder.search``).  Returns ssa pairs."""
    live = dict(live)
    owners = {}
    for i, m in live.items():
        mm = m
        while mm:
            low = mm & -mm
            owners.setdefault(low, set()).add(i)
            mm ^= low
    path = []

    def merged_size(low):
        u = 0
        for t in owners[low]:
            u |= live[t]
        return (u & ~low).bit_count() if not (low & out) else u.bit_count()

    size = {low: merged_size(low) for low in owners if not (low & out)}
    given = list(order) if order is not None else None
    while size:
        if given is not None:
            while given and given[0] not in size:
                given.pop(0)
            if not given:
                break
            best = given.pop(0)
        elif temperature > 0:
            best = min(size, key=lambda low: (size[low] - temperature * math.log(rng.random() + 1e-300) * -1.0, low))
        else:
            best = min(size, key=lambda low: (size[low], low))
        low = best
        del size[low]
        ids = sorted(owners[low], key=lambda t: (live[t].bit_count(), t))
        touched = 0
        while len(ids) > 1:
            i, j = ids[0], ids[1]
            a, b = live.pop(i), live.pop(j)
            shared = a & b
            keep = shared & out
            hs = shared & hyper
            while hs:
                l2 = hs & -hs
                hs ^= l2
                if len(owners[l2] - {i, j}) > 0:
                    keep |= l2
            c = (a ^ b) | keep
            for t, m in ((i, a), (j, b)):
                mm = m
                while mm:
                    l2 = mm & -mm
                    owners[l2].discard(t)
                    mm ^= l2
            live[nxt] = c
            mm = c
            while mm:
                l2 = mm & -mm
                owners.setdefault(l2, set()).add(nxt)
                mm ^= l2
            touched |= a | b
            path.append((i, j))
            ids = sorted(ids[2:] + [nxt], key=lambda t: (live[t].bit_count(), t))
            nxt += 1
        if len(ids) == 1 and not (low & out):
            # a label owned by a single tensor (trace-free dangling index cannot occur in circuit
            # networks; a label summed inside one tensor would need a reduction we do not have)
            touched |= live[ids[0]]
        mm = touched
        while mm:
            l2 = mm & -mm
            mm ^= l2
            if l2 in size:
                if owners.get(l2):
                    size[l2] = merged_size(l2)
                else:
                    del size[l2]
    # whatever is left shares no summed label: outer products, smallest first
    ids = sorted(live, key=lambda t: (live[t].bit_count(), t))
    while len(ids) > 1:
        i, j = ids[0], ids[1]
        a, b = live.pop(i), live.pop(j)
        live[nxt] = a | b
        path.append((i, j))
        ids = sorted(ids[2:] + [nxt], key=lambda t: (live[t].bit_count(), t))
        nxt += 1
    return path


# ----------------------------------------------------------------------- greedy
def _greedy(live, out, hyper, nxt, rng, alpha=1.0, temperature=0.0):
    """Boltzmann greedy on {id: mask}; returns ssa pairs finishing the contraction."""
    live = dict(live)
    owners = {}
    for i, m in live.items():
        mm = m
        while mm:
            low = mm & -mm
            owners.setdefault(low, set()).add(i)
            mm ^= low
    path = []
    heap = []

    def score(a, b):
        shared = a & b
        c = (a ^ b) | (shared & (out | hyper))
        s = 2.0 ** c.bit_count() - alpha * (2.0 ** a.bit_count() + 2.0 ** b.bit_count())
        if temperature > 0:
            g = -math.log(-math.log(rng.random() + 1e-300) + 1e-300)
            s = s - temperature * g * max(abs(s), 1.0)
        return s, c

    def push_pairs(i):
        a = live[i]
        mm = a
        seen = set()
        while mm:
            low = mm & -mm
            mm ^= low
            for j in owners.get(low, ()):
                if j != i and j not in seen and j in live:
                    seen.add(j)
                    s, c = score(a, live[j])
                    heapq.heappush(heap, (s, min(i, j), max(i, j)))

    for i in list(live):
        push_pairs(i)
    while len(live) > 1:
        pair = None
        while heap:
            s, i, j = heapq.heappop(heap)
            if i in live and j in live:
                pair = (i, j)
                break
        if pair is None:
            # disconnected components: outer-product the two smallest
            ids = sorted(live, key=lambda t: (live[t].bit_count(), t))
            pair = (ids[0], ids[1])
        i, j = pair
        a, b = live.pop(i), live.pop(j)
        shared = a & b
        keep = shared & out
        hs = shared & hyper
        while hs:
            low = hs & -hs
            hs ^= low
            if len(owners[low] - {i, j}) > 0 or (low & out):
                keep |= low
        c = (a ^ b) | keep
        for t, m in ((i, a), (j, b)):
            mm = m
            while mm:
                low = mm & -mm
                owners[low].discard(t)
                mm ^= low
        live[nxt] = c
        mm = c
        while mm:
            low = mm & -mm
            owners.setdefault(low, set()).add(nxt)
            mm ^= low
        path.append((i, j))
        push_pairs(nxt)
        nxt += 1
    return path


# ------------------------------------------------------- recursive bisection
def _bisect(ids, masks, rng, imbalance):
    """Split ``ids`` in two by a Fiedler-vector cut refined with FM passes.
    Edge weight between two tensors = number of shared labels."""
    n = len(ids)
    idx = {t: k for k, t in enumerate(ids)}
    owners = {}
    for t in ids:
        mm = masks[t]
        while mm:
            low = mm & -mm
            owners.setdefault(low, []).append(t)
            mm ^= low
    w = np.zeros((n, n))
    for ts in owners.values():
        for x in range(len(ts)):
            for y in range(x + 1, len(ts)):
                w[idx[ts[x]], idx[ts[y]]] += 1.0
                w[idx[ts[y]], idx[ts[x]]] += 1.0
    if n <= 2 or w.sum() == 0:
        half = max(1, n // 2)
        return list(ids[:half]), list(ids[half:])
    # node weights: log-size, so that big tensors count for more in the balance
    wt = np.array([max(1, masks[t].bit_count()) for t in ids], dtype=float)
    wpert = w * (1.0 + 0.3 * rng.random() * np.triu(np.asarray(
        [[rng.random() for _ in range(n)] for _ in range(n)]), 1))
    wpert = np.triu(wpert, 1)
    wpert = wpert + wpert.T
    lap = np.diag(wpert.sum(1)) - wpert
    try:
        vals, vecs = np.linalg.eigh(lap)
        f = vecs[:, 1] if n > 1 else vecs[:, 0]
    except np.linalg.LinAlgError:      # pragma: no cover
        f = np.array([rng.random() for _ in range(n)])
    order = np.argsort(f)
    total = wt.sum()
    lo_frac = 0.5 - imbalance
    # choose the split point along the Fiedler order with the smallest cut in the allowed window
    cum = np.cumsum(wt[order])
    side = np.zeros(n, dtype=bool)
    best_cut, best_k = None, None
    inside = np.zeros(n, dtype=bool)
    cut = 0.0
    for k in range(n - 1):
        v = order[k]
        cut += w[v, ~inside].sum() - w[v, inside].sum()
        inside[v] = True
        frac = cum[k] / total
        if lo_frac <= frac <= 1 - lo_frac or best_k is None and frac > 1 - lo_frac:
            if best_cut is None or cut < best_cut:
                best_cut, best_k = cut, k
    if best_k is None:
        best_k = n // 2 - 1
    side[order[: best_k + 1]] = True
    # FM-style refinement: move single vertices with positive gain while balance holds
    for _ in range(4):
        moved = False
        gains = np.where(side, w[:, ~side].sum(1) - w[:, side].sum(1),
                         w[:, side].sum(1) - w[:, ~side].sum(1))
        for v in np.argsort(-gains):
            if gains[v] <= 0:
                break
            g = (w[v, ~side].sum() - w[v, side].sum()) if side[v] else (w[v, side].sum() - w[v, ~side].sum())
            if g <= 0:
                continue
            new_in = wt[side].sum() + (-wt[v] if side[v] else wt[v])
            frac = new_in / total
            if not (lo_frac <= frac <= 1 - lo_frac):
                continue
            if side.sum() - (1 if side[v] else -1) in (0, n):
                continue
            side[v] = ~side[v]
            moved = True
        if not moved:
            break
    left = [ids[k] for k in range(n) if side[k]]
    right = [ids[k] for k in range(n) if not side[k]]
    if not left or not right:
        half = max(1, n // 2)
        return list(ids[:half]), list(ids[half:])
    return left, right


def _partition_path(live, out, hyper, nxt, rng, imbalance=0.15, leaf_size=10):
    """Build a contraction tree top-down by recursive bisection, greedy inside the leaves."""
    path = []
    state = {"nxt": nxt}
    masks = dict(live)

    def outer_mask(ids, all_ids_mask_count):
        return None

    def build(ids):
        # returns the id of the tensor that results from contracting ``ids``
        if len(ids) == 1:
            return ids[0]
        if len(ids) <= leaf_size:
            sub = {t: masks[t] for t in ids}
            # labels leaving this group must survive: treat them as output
            inner = 0
            seen_once = 0
            for t in ids:
                inner |= seen_once & masks[t]
                seen_once |= masks[t]
            ext = seen_once & ~inner
            # a label seen twice inside may still be external when hyper; keep those too
            p = _greedy(sub, out | ext | (hyper & seen_once), 0, state["nxt"], rng, temperature=0.0)
            # recompute resulting masks along p
            for (i, j) in p:
                a, b = masks[i], masks[j]
                shared = a & b
                c = (a ^ b) | (shared & (out | ext | hyper))
                masks[state["nxt"]] = c
                path.append((i, j))
                state["nxt"] += 1
            return state["nxt"] - 1
        left, right = _bisect(ids, masks, rng, imbalance)
        li = build(left)
        ri = build(right)
        a, b = masks[li], masks[ri]
        shared = a & b
        # labels shared between the halves and nowhere else are summed
        c = (a ^ b) | (shared & (out | hyper))
        masks[state["nxt"]] = c
        path.append((li, ri))
        state["nxt"] += 1
        return state["nxt"] - 1

    build(sorted(live))
    return path


# --------------------------------------------------------------------- slicing
def _slice_more(net: SymbolicNetwork, path, sliced, target_log2_width, max_sliced=40):
    """Continue slicing ``path`` from the labels already in ``sliced``; returns (mask, the added single-bit masks)."""
    n0 = sliced.bit_count()
    sliced2, order = _slice(net, path, target_log2_width, 1, max_sliced=max(0, max_sliced - n0), sliced0=sliced)
    return sliced2, order


def _slice(net: SymbolicNetwork, path, target_log2_width, min_slices, max_sliced=40, sliced0=0):
    sliced = sliced0
    order = []
    cost, width, _, _ = path_cost(net, path, sliced)
    want_bits = max(0, math.ceil(math.log2(max(1, min_slices))))
    while (width > target_log2_width or len(order) < want_bits) and len(order) < max_sliced:
        # candidates: labels living in the widest intermediates (or anything, when only
        # more slices are wanted)
        cand = 0
        for _, a, b, c, _d in _walk(net, path, sliced):
            for m in (a, b, c):
                if m.bit_count() >= width - (0 if width > target_log2_width else 2):
                    cand |= m
        cand &= ~net.out_mask                # hyper labels included: every owner leaf is indexed at the slice bit
        if not cand:
            break
        best = None
        mm = cand
        while mm:
            low = mm & -mm
            mm ^= low
            c2, w2, _, _ = path_cost(net, path, sliced | low)
            key = (w2 if width > target_log2_width else 0, c2)
            if best is None or key < best[0]:
                best = (key, low, c2, w2)
        _, low, cost, width = best
        sliced |= low
        order.append(low)
    return sliced, order


# ------------------------------------------------------------------ front door
def find_path(modes, out, samples=32, seed=0, target_log2_width=30, min_slices=1,
              methods=("greedy", "partition"), time_budget_s=None, refine=3, reconf_rounds=6,
              subtree_size=8, anneal_steps=None):
    """Search a contraction path and slicing for the network ``(modes, out)``.

    ``samples`` and ``seed`` play the role of cuTensorNet's ``samples`` hyper-
    optimiser knob (the reference passes 32 from the backend,
    ``backends/cutensornet.py:110,119,136,147``); ranks may search with
    different seeds and vote on ``opt_cost`` (``eval.py:191-194``).

    Stages: rank-<=2 absorption; ``samples`` x ``methods`` candidate trees; the
    ``refine`` cheapest are polished by subtree reconfiguration, sliced down to
    ``2^target_log2_width`` elements (and >= ``min_slices`` slices) with the tree
    re-polished as labels are removed, then annealed over tree rotations (``anneal_steps``
    moves; default scales with the network, 0 disables) with the sliced labels removed and
    re-sliced; the cheapest sliced plan wins.
    """
    import time

    from .treeopt import HyperTree, Tree

    net = SymbolicNetwork(modes, out)
    pre, live, nxt = _absorb_small(net)
    rng = random.Random(seed)
    t0 = time.time()
    deadline = None if time_budget_s is None else t0 + time_budget_s
    cands = []
    if len(live) <= 1:
        cands.append((0.0, "trivial", []))
    else:
        if net.hyper_mask:
            # label-graph elimination orders (elimorder.py): randomised min-fill / min-degree restarts, the
            # cheapest few by the order's own cost proxy become candidate trees
            from . import elimorder

            nb = elimorder.label_graph(list(live.values()))
            orders = []
            for s in range(max(4, 4 * samples)):
                rule = "fill" if s % 4 != 3 else "degree"
                noise = 0.0 if s < 4 else rng.choice([0.1, 0.3, 1.0])
                res = elimorder.greedy_order(nb, net.out_mask, rng, noise, rule)
                orders.append((res[2], res[1], res[0]))
                if deadline is not None and time.time() > t0 + 0.2 * time_budget_s:
                    break
            orders.sort(key=lambda o: (o[0], o[1]))
            for _c, _w, order in orders[:max(2, samples)]:
                tail = _elimination_path(live, net.out_mask, net.hyper_mask, nxt, rng, 0.0,
                                         order=[1 << v for v in order])
                cost, width, _, _ = path_cost(net, pre + tail, 0)
                cands.append((cost * (1.0 + 2.0 ** max(0, width - target_log2_width)) ** 0.5, "order", tail))
        for s in range(max(1, samples)):
            for meth in (tuple(methods) + ("eliminate",) if net.hyper_mask else methods):
                if meth == "eliminate":
                    temp = 0.0 if s == 0 else rng.choice([0.3, 0.6, 1.0, 1.5])
                    tail = _elimination_path(live, net.out_mask, net.hyper_mask, nxt, rng, temp)
                elif meth == "greedy":
                    temp = 0.0 if s == 0 else rng.choice([0.01, 0.03, 0.1, 0.3])
                    alpha = 1.0 if s == 0 else rng.choice([0.0, 0.5, 1.0, 1.0, 2.0])
                    tail = _greedy(live, net.out_mask, net.hyper_mask, nxt, rng, alpha, temp)
                else:
                    imb = rng.choice([0.02, 0.05, 0.1, 0.15, 0.2, 0.3, 0.4])
                    leaf = rng.choice([4, 6, 8, 12])
                    tail = _partition_path(live, net.out_mask, net.hyper_mask, nxt, rng, imb, leaf)
                cost, width, _, _ = path_cost(net, pre + tail, 0)
                cands.append((cost * (1.0 + 2.0 ** max(0, width - target_log2_width)) ** 0.5, meth, tail))
            if deadline is not None and time.time() > t0 + 0.4 * time_budget_s:
                break
    if net.hyper_mask and len(cands) > 3 * max(1, refine):
        # polishing a hyper tree is cheap but not free: only the more promising candidates get it
        cands.sort(key=lambda c: c[0])
        cands = cands[:3 * max(1, refine)]
    TreeCls = HyperTree if net.hyper_mask else Tree
    # every candidate gets a cheap polish (small subtrees); ranking happens after it
    polished = []
    can_refine = len(live) >= 3
    for score, meth, tail in cands:
        if can_refine and meth != "trivial":
            tree = TreeCls(live, tail, nxt, net.out_mask)
            tree.reconfigure(rounds=reconf_rounds, k=min(8, subtree_size), deadline=deadline)
            w = tree.width()
            score = tree.total_cost() * (1.0 + 2.0 ** max(0, w - target_log2_width)) ** 0.5
            polished.append((score, meth, tail, tree))
        else:
            polished.append((score, meth, tail, None))
        if deadline is not None and time.time() > t0 + 0.7 * time_budget_s:
            break
    polished.sort(key=lambda c: c[0])
    best = None
    for _score, meth, tail, tree in polished[:max(1, refine)]:
        if tree is None:
            path = pre + tail
            sliced, order = _slice(net, path, target_log2_width, min_slices)
        else:
            if subtree_size > 8:
                tree.reconfigure(rounds=3, k=subtree_size, deadline=deadline)
            order, sliced = tree.slice_greedy(target_log2_width, min_slices, k=min(8, subtree_size),
                                              deadline=deadline)
            if order:
                tree.reconfigure(rounds=3, k=subtree_size, removed=sliced, deadline=deadline)
                # re-polishing may have widened a node again: slice further if so
                more, sliced = tree.slice_greedy(target_log2_width, 1, removed=sliced,
                                                 reconf_every=0)
                order += more
            steps = anneal_steps
            if steps is None:
                steps = 0 if len(live) < 64 else 150 * len(live)
            if steps > 0:
                # annealing moves single subtrees across the tree, which reconfiguration (bounded
                # subtrees) cannot; the result is kept only if the sliced total really went down
                snap = (tree.snapshot(), list(order), sliced)
                before = tree.total_cost(sliced) * (1 << len(order))
                tree.anneal(steps, rng, removed=sliced, max_width=target_log2_width, deadline=deadline)
                tree.reconfigure(rounds=2, k=min(8, subtree_size), removed=sliced, deadline=deadline)
                if tree.width(sliced) > target_log2_width:
                    more, sliced = tree.slice_greedy(target_log2_width, 1, removed=sliced, reconf_every=0)
                    order = order + more
                if tree.total_cost(sliced) * (1 << len(order)) >= before:
                    tree.restore(snap[0])
                    order, sliced = snap[1], snap[2]
            path = pre + tree.ssa_pairs(nxt)
        cost, width, per_slice, hoisted = path_cost(net, path, sliced)
        if best is None or cost < best[0]:
            best = (cost, width, per_slice, hoisted, path, order, meth)
        if deadline is not None and time.time() > deadline:
            break
    cost, width, per_slice, hoisted, path, order, meth = best
    labels = [net.bit_label[low.bit_length() - 1] for low in order]
    return PathInfo(path=path, sliced_labels=labels, num_slices=1 << len(labels),
                    opt_cost=cost, flops=8.0 * cost, log2_width=width,
                    per_slice_cost=per_slice, hoisted_cost=hoisted, seed=seed, method=meth)


def compute_slices(num_slices, rank, size):
    """Contiguous slice-id range of ``rank``; the remainder goes to the low ranks.
    Same partition as ``compute_slices`` in ``/root/reference/src/qibotn/eval.py:197-205``."""
    num_slices = int(getattr(num_slices, "num_slices", num_slices))     # the reference passes its `info` object
    chunk, extra = divmod(num_slices, size)
    begin = rank * chunk + min(rank, extra)
    end = num_slices if rank == size - 1 else (rank + 1) * chunk + min(rank + 1, extra)
    return range(begin, end)
