"""Contraction-tree refinement: subtree reconfiguration and index slicing.

Given a binary contraction tree over tensors whose labels are bit masks, this
module (a) repeatedly re-solves small subtrees exactly by dynamic programming
over subsets and (b) picks sliced labels greedily by how much of the total work
they take part in, re-optimising the tree with the sliced labels removed.
It is the part of the reference's path search that cuTensorNet / cotengra do
behind ``contract_path(optimize={"samples": ..., "slicing": ...})``
(``/root/reference/src/qibotn/eval.py:180-190``).

``Tree`` is for networks without hyper labels (every label owned by at most two
tensors, or one tensor and the output): then the open legs of a group of
tensors are simply the XOR of their masks, plus any output labels they carry.
``HyperTree`` (end of the file) counts owners per label and handles the rest.
"""

from __future__ import annotations

import math


class Tree:
    """Binary contraction tree.  ``children[n] = (l, r)``; leaves have no entry."""

    def __init__(self, leaf_masks, pairs, first_id, out_mask):
        self.out = out_mask
        self.mask = dict(leaf_masks)
        self.children = {}
        self.parent = {}
        nxt = first_id
        for i, j in pairs:
            self.children[nxt] = (i, j)
            self.parent[i] = nxt
            self.parent[j] = nxt
            self.mask[nxt] = self._legs(self.mask[i], self.mask[j])
            nxt += 1
        self.root = nxt - 1 if pairs else next(iter(leaf_masks))
        self.next_id = nxt

    def _legs(self, a, b):
        return (a ^ b) | (a & b & self.out)

    def snapshot(self):
        return (dict(self.children), dict(self.mask), dict(self.parent), self.next_id)

    def restore(self, snap):
        self.children, self.mask, self.parent, self.next_id = dict(snap[0]), dict(snap[1]), dict(snap[2]), snap[3]

    # ------------------------------------------------------------------ costs
    def node_cost(self, n, removed=0):
        l, r = self.children[n]
        return 2.0 ** ((self.mask[l] | self.mask[r]) & ~removed).bit_count()

    def total_cost(self, removed=0):
        return sum(self.node_cost(n, removed) for n in self.children)

    def width(self, removed=0):
        return max((m & ~removed).bit_count() for m in self.mask.values())

    def ssa_pairs(self, first_id):
        """Post-order pairs with fresh SSA ids starting at ``first_id``."""
        order, ids = [], {}
        stack = [(self.root, False)]
        nxt = first_id
        while stack:
            n, done = stack.pop()
            if n not in self.children:
                ids[n] = n
                continue
            if done:
                l, r = self.children[n]
                order.append((ids[l], ids[r]))
                ids[n] = nxt
                nxt += 1
            else:
                stack.append((n, True))
                l, r = self.children[n]
                stack.append((r, False))
                stack.append((l, False))
        return order

    # ------------------------------------------------------- subtree reconfiguration
    def _frontier(self, root, k, removed):
        front = [root]
        while len(front) < k:
            best, bi = -1, -1
            for idx, n in enumerate(front):
                if n in self.children:
                    s = (self.mask[n] & ~removed).bit_count()
                    if s > best:
                        best, bi = s, idx
            if bi < 0:
                break
            n = front.pop(bi)
            front.extend(self.children[n])
        return front

    def _inner_nodes(self, root, front):
        fs = set(front)
        inner, stack = [], [root]
        while stack:
            n = stack.pop()
            if n in fs:
                continue
            inner.append(n)
            stack.extend(self.children[n])
        return inner

    def reconfigure_at(self, root, k=8, removed=0, write_weight=64.0):
        """Optimally re-contract the <= k sub-leaves below ``root``.  Returns the gain."""
        if root not in self.children:
            return 0.0
        front = self._frontier(root, k, removed)
        n = len(front)
        if n < 3:
            return 0.0
        inner = self._inner_nodes(root, front)
        keep = ~removed
        out = self.out

        def w_of(mask):          # cost of producing an intermediate with these legs
            return write_weight * 2.0 ** mask.bit_count()

        old = 0.0
        for x in inner:
            l, r = self.children[x]
            old += 2.0 ** ((self.mask[l] | self.mask[r]) & keep).bit_count()
            if x != root:
                old += w_of(self.mask[x] & keep)
        masks = [self.mask[f] & keep for f in front]
        full = (1 << n) - 1
        xorm = [0] * (full + 1)
        orm = [0] * (full + 1)
        for s in range(1, full + 1):
            low = s & -s
            i = low.bit_length() - 1
            rest = s ^ low
            xorm[s] = xorm[rest] ^ masks[i]
            orm[s] = orm[rest] | masks[i]
        legs = [xorm[s] | (orm[s] & out) for s in range(full + 1)]
        inf = float("inf")
        best = [inf] * (full + 1)
        split = [0] * (full + 1)
        for i in range(n):
            best[1 << i] = 0.0
        # subsets in increasing popcount order = increasing integer order suffices for
        # sub < s, since every proper subset is a smaller integer
        for s in range(3, full + 1):
            if s & (s - 1) == 0:
                continue
            low = s & -s
            bs, bsplit = inf, 0
            wr = 0.0 if s == full else w_of(legs[s])
            sub = (s - 1) & s
            while sub:
                if sub & low:
                    other = s ^ sub
                    c = best[sub] + best[other]
                    if c < bs:
                        c += 2.0 ** (legs[sub] | legs[other]).bit_count() + wr
                        if c < bs:
                            bs, bsplit = c, sub
                sub = (sub - 1) & s
            best[s], split[s] = bs, bsplit
        new = best[full]
        if not new < old * (1.0 - 1e-9):
            return 0.0
        # rebuild: reuse the ids of the old inner nodes (root keeps its id)
        spare = [x for x in inner if x != root]
        for x in inner:
            del self.children[x]

        def build(s, node_id=None):
            if s & (s - 1) == 0:
                return front[s.bit_length() - 1]
            a, b = split[s], s ^ split[s]
            la, lb = build(a), build(b)
            nid = node_id if node_id is not None else (spare.pop() if spare else self._fresh())
            self.children[nid] = (la, lb)
            self.parent[la] = nid
            self.parent[lb] = nid
            self.mask[nid] = self._legs(self.mask[la], self.mask[lb])
            return nid

        build(full, root)
        for x in spare:                     # fewer inner nodes can never happen, but be safe
            self.mask.pop(x, None)
            self.parent.pop(x, None)
        return old - new

    def _fresh(self):
        self.next_id += 1
        return self.next_id - 1

    def reconfigure(self, rounds=4, k=8, removed=0, write_weight=64.0, deadline=None):
        """Sweep all internal nodes (largest first) until a round brings < 0.1 %."""
        import time

        for _ in range(rounds):
            before = self.total_cost(removed)
            nodes = sorted(self.children, key=lambda x: -self.node_cost(x, removed))
            for x in nodes:
                if x in self.children:
                    self.reconfigure_at(x, k, removed, write_weight)
                if deadline is not None and time.time() > deadline:
                    return
            after = self.total_cost(removed)
            if after > before * 0.999:
                break

    # ------------------------------------------------------------------- slicing
    def slice_greedy(self, target_log2_width, min_slices=1, removed=0, max_sliced=48,
                     reconf_every=4, k=8, deadline=None):
        """Choose sliced labels one at a time: among the labels carried by every widest
        tensor (or by any, if none is common), take the one that takes part in the most
        work, which minimises the sliced total.  Returns the ordered list of label bits."""
        order = []
        want_bits = max(0, math.ceil(math.log2(max(1, min_slices))))
        forbidden = self.out
        while len(order) < max_sliced:
            width = self.width(removed)
            need_width = width > target_log2_width
            if not need_width and len(order) >= want_bits:
                break
            if need_width:
                widest = [m & ~removed for m in self.mask.values()
                          if (m & ~removed).bit_count() == width]
                cand = ~0
                for m in widest:
                    cand &= m
                cand &= ~forbidden
                if cand == 0 or cand == ~0 & ~forbidden:
                    cand = 0
                    for m in widest:
                        cand |= m
                    cand &= ~forbidden
            else:
                cand = 0
                for m in self.mask.values():
                    if (m & ~removed).bit_count() >= width - 1:
                        cand |= m & ~removed
                cand &= ~forbidden
            if cand == 0:
                break
            score = {}
            for n in self.children:
                l, r = self.children[n]
                u = (self.mask[l] | self.mask[r]) & ~removed
                hit = u & cand
                if not hit:
                    continue
                c = 2.0 ** u.bit_count()
                while hit:
                    low = hit & -hit
                    score[low] = score.get(low, 0.0) + c
                    hit ^= low
            if not score:
                break
            low = max(score, key=lambda b: (score[b], -b))
            removed |= low
            order.append(low)
            if reconf_every and len(order) % reconf_every == 0:
                self.reconfigure(rounds=1, k=k, removed=removed, deadline=deadline)
        return order, removed

    # ------------------------------------------------------------------ annealing
    def anneal(self, steps, rng, removed=0, t_start=3e-3, t_end=3e-5, write_weight=128.0,
               max_width=None, deadline=None):
        """Simulated annealing over tree rotations.  A move takes an internal node p = (x, y) whose
        child x = (a, b) is internal and regroups the three subtrees as (a, (b, y)) or ((a, y), b):
        p keeps its legs, so only x changes and the cost difference is local.  Objective =
        sum over nodes of MACs + write_weight * (elements of the intermediate), the same trade-off
        between tensor-core time and HBM time the executor sees.  Moves that would create an
        intermediate above 2^max_width are rejected.  Returns the number of accepted moves."""
        import math
        import time

        keep = ~removed

        def size(mask):
            return 2.0 ** (mask & keep).bit_count()

        def ncost(l, r):
            return 2.0 ** ((self.mask[l] | self.mask[r]) & keep).bit_count()

        total = sum(ncost(*self.children[n]) + write_weight * size(self.mask[n]) for n in self.children)
        inner = [n for n in self.children]
        accepted = 0
        log_ratio = math.log(t_end / t_start)
        for it in range(steps):
            if deadline is not None and (it & 1023) == 0 and time.time() > deadline:
                break
            temp = t_start * math.exp(log_ratio * it / max(1, steps - 1))
            p = inner[int(rng.random() * len(inner))]
            l, r = self.children[p]
            if rng.random() < 0.5:
                l, r = r, l
            x, y = l, r                                   # x must be internal
            if x not in self.children:
                x, y = y, x
                if x not in self.children:
                    continue
            a, b = self.children[x]
            if rng.random() < 0.5:
                a, b = b, a
            # new x = (b, y), new p = (a, x)
            old = ncost(a, b) + ncost(x, y) + write_weight * size(self.mask[x])
            new_mask = self._legs(self.mask[b], self.mask[y])
            if max_width is not None and (new_mask & keep).bit_count() > max_width:
                continue
            new_x_cost = 2.0 ** ((self.mask[b] | self.mask[y]) & keep).bit_count()
            new_p_cost = 2.0 ** ((self.mask[a] | new_mask) & keep).bit_count()
            new = new_x_cost + new_p_cost + write_weight * size(new_mask)
            delta = new - old
            if delta > 0 and rng.random() >= math.exp(-(delta / total) / temp):
                continue
            self.children[x] = (b, y)
            self.children[p] = (a, x)
            self.parent[y] = x
            self.parent[a] = p
            self.mask[x] = new_mask
            total += delta
            accepted += 1
        return accepted


# ----------------------------------------------------------------------------------------------
# Hyper-aware trees.  With hyper labels (owned by more than two tensors: the diagonal gates of a
# QAOA circuit, BASELINE config 5) the open legs of a group of tensors are no longer the XOR of
# their masks: a label stays open until ALL its owners are inside the group.  Every node therefore
# carries, next to its legs, the number of owners of each label below it, as a bit-sliced counter
# (one Python int per binary digit of the count), so that counts of unions are a few bitwise
# operations -- a ripple-carry add over whole label sets -- and "all owners inside" is a
# comparison of digit planes with the totals.
# ----------------------------------------------------------------------------------------------
def _planes_add(x, y):
    out, carry = [], 0
    for a, b in zip(x, y):
        s = a ^ b
        out.append(s ^ carry)
        carry = (a & b) | (s & carry)
    return tuple(out)


class HyperTree(Tree):
    """Binary contraction tree over tensors with hyper labels; same interface as ``Tree``."""

    def __init__(self, leaf_masks, pairs, first_id, out_mask):
        count = {}
        for m in leaf_masks.values():
            mm = m
            while mm:
                low = mm & -mm
                count[low] = count.get(low, 0) + 1
                mm ^= low
        self.nplanes = max(1, max(count.values(), default=1).bit_length())
        self.total = tuple(sum(low for low, c in count.items() if (c >> k) & 1) for k in range(self.nplanes))
        self.cnt = {i: (m,) + (0,) * (self.nplanes - 1) for i, m in leaf_masks.items()}
        self.out = out_mask
        self.mask = dict(leaf_masks)
        self.children = {}
        self.parent = {}
        nxt = first_id
        for i, j in pairs:
            self.children[nxt] = (i, j)
            self.parent[i] = nxt
            self.parent[j] = nxt
            self.cnt[nxt] = _planes_add(self.cnt[i], self.cnt[j])
            self.mask[nxt] = self._legs_of(self.cnt[nxt])
            nxt += 1
        self.root = nxt - 1 if pairs else next(iter(leaf_masks))
        self.next_id = nxt

    def _legs_of(self, cnt):
        present = diff = 0
        for c, t in zip(cnt, self.total):
            present |= c
            diff |= c ^ t
        return present & (diff | self.out)

    def _legs(self, a, b):                   # the base class's mask-only rule does not apply here
        raise NotImplementedError

    def snapshot(self):
        return (dict(self.children), dict(self.mask), dict(self.parent), self.next_id, dict(self.cnt))

    def restore(self, snap):
        self.children, self.mask, self.parent, self.next_id, self.cnt = (
            dict(snap[0]), dict(snap[1]), dict(snap[2]), snap[3], dict(snap[4]))

    def reconfigure_at(self, root, k=8, removed=0, write_weight=64.0):
        if root not in self.children:
            return 0.0
        front = self._frontier(root, k, removed)
        n = len(front)
        if n < 3:
            return 0.0
        inner = self._inner_nodes(root, front)
        keep = ~removed

        def w_of(mask):
            return write_weight * 2.0 ** mask.bit_count()

        old = 0.0
        for x in inner:
            l, r = self.children[x]
            old += 2.0 ** ((self.mask[l] | self.mask[r]) & keep).bit_count()
            if x != root:
                old += w_of(self.mask[x] & keep)
        full = (1 << n) - 1
        zero = (0,) * self.nplanes
        cnts = [zero] * (full + 1)
        legs = [0] * (full + 1)
        fc = [self.cnt[f] for f in front]
        for s in range(1, full + 1):
            low = s & -s
            i = low.bit_length() - 1
            rest = s ^ low
            c = fc[i] if rest == 0 else _planes_add(cnts[rest], fc[i])
            cnts[s] = c
            legs[s] = self._legs_of(c) & keep
        inf = float("inf")
        best = [inf] * (full + 1)
        split = [0] * (full + 1)
        for i in range(n):
            best[1 << i] = 0.0
        for s in range(3, full + 1):
            if s & (s - 1) == 0:
                continue
            low = s & -s
            bs, bsplit = inf, 0
            wr = 0.0 if s == full else w_of(legs[s])
            sub = (s - 1) & s
            while sub:
                if sub & low:
                    other = s ^ sub
                    c = best[sub] + best[other]
                    if c < bs:
                        c += 2.0 ** (legs[sub] | legs[other]).bit_count() + wr
                        if c < bs:
                            bs, bsplit = c, sub
                sub = (sub - 1) & s
            best[s], split[s] = bs, bsplit
        new = best[full]
        if not new < old * (1.0 - 1e-9):
            return 0.0
        spare = [x for x in inner if x != root]
        for x in inner:
            del self.children[x]

        def build(s, node_id=None):
            if s & (s - 1) == 0:
                return front[s.bit_length() - 1]
            a, b = split[s], s ^ split[s]
            la, lb = build(a), build(b)
            nid = node_id if node_id is not None else (spare.pop() if spare else self._fresh())
            self.children[nid] = (la, lb)
            self.parent[la] = nid
            self.parent[lb] = nid
            self.cnt[nid] = cnts[s]
            self.mask[nid] = self._legs_of(cnts[s])
            return nid

        build(full, root)
        for x in spare:
            self.mask.pop(x, None)
            self.parent.pop(x, None)
            self.cnt.pop(x, None)
        return old - new

    def anneal(self, steps, rng, removed=0, t_start=3e-3, t_end=3e-5, write_weight=128.0,
               max_width=None, deadline=None):
        import time

        keep = ~removed

        def size(mask):
            return 2.0 ** (mask & keep).bit_count()

        def ncost(l, r):
            return 2.0 ** ((self.mask[l] | self.mask[r]) & keep).bit_count()

        total = sum(ncost(*self.children[n]) + write_weight * size(self.mask[n]) for n in self.children)
        inner = [n for n in self.children]
        accepted = 0
        log_ratio = math.log(t_end / t_start)
        for it in range(steps):
            if deadline is not None and (it & 1023) == 0 and time.time() > deadline:
                break
            temp = t_start * math.exp(log_ratio * it / max(1, steps - 1))
            p = inner[int(rng.random() * len(inner))]
            l, r = self.children[p]
            if rng.random() < 0.5:
                l, r = r, l
            x, y = l, r
            if x not in self.children:
                x, y = y, x
                if x not in self.children:
                    continue
            a, b = self.children[x]
            if rng.random() < 0.5:
                a, b = b, a
            old = ncost(a, b) + ncost(x, y) + write_weight * size(self.mask[x])
            new_cnt = _planes_add(self.cnt[b], self.cnt[y])
            new_mask = self._legs_of(new_cnt)
            if max_width is not None and (new_mask & keep).bit_count() > max_width:
                continue
            new_x_cost = 2.0 ** ((self.mask[b] | self.mask[y]) & keep).bit_count()
            new_p_cost = 2.0 ** ((self.mask[a] | new_mask) & keep).bit_count()
            new = new_x_cost + new_p_cost + write_weight * size(new_mask)
            delta = new - old
            if delta > 0 and rng.random() >= math.exp(-(delta / total) / temp):
                continue
            self.children[x] = (b, y)
            self.children[p] = (a, x)
            self.parent[y] = x
            self.parent[a] = p
            self.mask[x] = new_mask
            self.cnt[x] = new_cnt
            total += delta
            accepted += 1
        return accepted
