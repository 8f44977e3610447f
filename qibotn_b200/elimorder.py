"""Elimination-order search with slicing for hyper-indexed networks (BASELINE config 5).

When the entangling gates of a circuit are diagonal (QAOA: RZZ layers, the ZZ observable) the
expectation network ``network.expectation_network(..., hyper=True)`` has few labels, most of them
shared by many tensors.  A contraction tree for such a network is best found on its *label graph*
(labels = vertices, every tensor = a clique): a vertex elimination order is a contraction plan whose
largest intermediate is the largest neighbourhood met on the way (the width of the order), and
fixing a label to 0/1 -- slicing -- deletes its vertex.  This is the formulation of QTensor
(Lykov et al.) and of the bucket-elimination literature; the reference reaches the same thing through
cuTensorNet's hyper-optimizer with ``slicing`` (``/root/reference/src/qibotn/eval.py:334-352``).

Search: randomised min-fill / min-degree orders (noise on the score, many restarts: one restart is a
few hundred set operations on Python ints), then greedy slicing where every candidate label is scored
by re-running the order search on the graph without it, so the order adapts to the slicing instead of
being fixed before it.  ``planner._elimination_path`` turns the winning order into pairwise SSA
contractions; costs quoted in plans come from ``planner.path_cost`` on that path, not from the proxy
used here.
"""

from __future__ import annotations

import math
import random
import time

__all__ = ["label_graph", "greedy_order", "search", "find_path_elimination"]


def _bits(mask):
    while mask:
        low = mask & -mask
        yield low.bit_length() - 1
        mask ^= low


def label_graph(masks, removed=0):
    """Adjacency of the label graph as {label bit index: neighbour mask}; ``removed`` labels are gone."""
    nb = {}
    for m in masks:
        m &= ~removed
        for v in _bits(m):
            nb[v] = nb.get(v, 0) | m
    for v in nb:
        nb[v] &= ~(1 << v)
    return nb


def greedy_order(nb, keep=0, rng=None, noise=0.0, rule="fill", cap=None):
    """One elimination order of the graph ``nb`` (not modified).  Labels in ``keep`` (the output) are never
    eliminated.  Returns (order, width, log2 cost proxy) where width = the largest neighbourhood at elimination
    time and the proxy is log2 sum 2^(neighbourhood + 1).  ``cap``: give up (return None) above this width."""
    nb = dict(nb)
    rng = rng or random
    score = {}

    def fill(v):
        n = nb[v]
        d = n.bit_count()
        if d <= 1:
            return 0
        miss = 0
        for u in _bits(n):
            miss += (n & ~nb[u]).bit_count() - 1
        return miss // 2

    def sc(v):
        d = nb[v].bit_count()
        if rule == "degree":
            s = float(d)
        else:
            # min-fill with the degree as a soft tie-break: a zero-fill vertex of any degree is free
            s = fill(v) + 0.01 * d
        if noise > 0:
            s += noise * rng.random() * (1.0 + 0.25 * d)
        return s

    for v in nb:
        if not (keep >> v) & 1:
            score[v] = sc(v)
    order, width, cost = [], 0, 0.0
    while score:
        v = min(score, key=score.__getitem__)
        del score[v]
        n = nb.pop(v)
        d = n.bit_count()
        if cap is not None and d > cap:
            return None
        width = max(width, d)
        cost += 2.0 ** (d + 1)
        order.append(v)
        vb = 1 << v
        touched = n
        for u in _bits(n):
            new = (nb[u] | n) & ~(1 << u) & ~vb
            nb[u] = new
        if rule != "degree":
            # fill of second neighbours changes too when edges appear between members of n
            for u in _bits(n):
                touched |= nb[u]
        for u in _bits(touched):
            if u in score:
                score[u] = sc(u)
    return order, width, math.log2(max(cost, 1.0))


def _best_of(nb, keep, rng, restarts, deadline=None, cap=None):
    best = None
    for r in range(max(1, restarts)):
        rule = "fill" if r % 3 != 2 else "degree"
        noise = 0.0 if r < 2 else rng.choice([0.3, 0.6, 1.0, 2.0])
        res = greedy_order(nb, keep, rng, noise, rule, cap)
        if res is not None and (best is None or (res[2], res[1]) < (best[2], best[1])):
            best = res
        if deadline is not None and time.time() > deadline:
            break
    return best


def search(masks, out_mask, target_width, min_slices=1, restarts=64, slice_restarts=6, max_candidates=32,
           seed=0, time_budget_s=None, max_sliced=40):
    """Order + sliced labels for the network ``masks`` (ints over label bit indices).
    Returns (order, sliced_bits, width, log2 total-cost proxy)."""
    rng = random.Random(seed)
    t0 = time.time()
    deadline = None if time_budget_s is None else t0 + time_budget_s
    first_deadline = None if time_budget_s is None else t0 + 0.35 * time_budget_s
    nb = label_graph(masks)
    best = _best_of(nb, out_mask, rng, restarts, first_deadline)
    order, width, cost = best
    sliced = []
    removed = 0
    want_bits = max(0, math.ceil(math.log2(max(1, min_slices))))
    while (width > target_width or len(sliced) < want_bits) and len(sliced) < max_sliced:
        # candidates: labels of the widest neighbourhoods of the current order
        g = dict(nb)
        cand = 0
        for v in order:
            n = g.pop(v)
            if n.bit_count() >= width - 1:
                cand |= n | (1 << v)
            vb = 1 << v
            for u in _bits(n):
                g[u] = (g[u] | n) & ~(1 << u) & ~vb
        cand &= ~out_mask
        cands = list(_bits(cand))
        if not cands:
            break
        rng.shuffle(cands)
        # the labels of highest degree first: they cut the most
        cands.sort(key=lambda v: -nb[v].bit_count())
        pick = None
        for v in cands[:max_candidates]:
            g2 = label_graph(masks, removed | (1 << v))
            res = _best_of(g2, out_mask, rng, slice_restarts)
            if res is None:
                continue
            key = (max(res[1], target_width), res[2])
            if pick is None or key < pick[0]:
                pick = (key, v, res)
            if deadline is not None and time.time() > deadline and pick is not None:
                break
        if pick is None:
            break
        _, v, res = pick
        removed |= 1 << v
        sliced.append(v)
        nb = label_graph(masks, removed)
        # polish the order on the reduced graph
        more = _best_of(nb, out_mask, rng, max(4, restarts // 4), None if deadline is None else min(deadline, time.time() + 5.0))
        if more is not None and (more[2], more[1]) < (res[2], res[1]):
            res = more
        order, width, cost = res
    return order, sliced, width, cost + len(sliced)


def find_path_elimination(modes, out, target_log2_width=30, min_slices=1, restarts=64, seed=0,
                          time_budget_s=None):
    """``planner.find_path`` for hyper-indexed networks through the label-graph search; returns a PathInfo."""
    from .planner import PathInfo, SymbolicNetwork, _absorb_small, _elimination_path, path_cost

    net = SymbolicNetwork(modes, out)
    pre, live, nxt = _absorb_small(net)
    masks = list(live.values())
    order, sliced_bits, _w, _c = search(masks, net.out_mask, target_log2_width, min_slices, restarts,
                                        seed=seed, time_budget_s=time_budget_s)
    sliced = 0
    for b in sliced_bits:
        sliced |= 1 << b
    rng = random.Random(seed)
    reduced = {i: m & ~sliced for i, m in live.items()}
    tail = _elimination_path(reduced, net.out_mask, net.hyper_mask, nxt, rng, 0.0,
                             order=[1 << v for v in order])
    path = pre + tail
    cost, width, per_slice, hoisted = path_cost(net, path, sliced)
    # the pairwise realisation of a bucket can be one label wider than the order's neighbourhood (a label that
    # is summed only when its last two owners meet): slice further with the generic greedy rule if so
    extra = []
    if width > target_log2_width:
        from .planner import _slice_more

        sliced, extra = _slice_more(net, path, sliced, target_log2_width)
        cost, width, per_slice, hoisted = path_cost(net, path, sliced)
    labels = [net.bit_label[b] for b in sliced_bits] + [net.bit_label[low.bit_length() - 1] for low in extra]
    return PathInfo(path=path, sliced_labels=labels, num_slices=1 << len(labels), opt_cost=cost,
                    flops=8.0 * cost, log2_width=width, per_slice_cost=per_slice, hoisted_cost=hoisted,
                    seed=seed, method="elimination-order")
