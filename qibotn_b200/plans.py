"""On-disk cache of contraction plans (path + sliced labels) for named workloads.

Planning is CPU work that can dominate small runs; a plan depends only on the
network *structure*, so benchmark workloads ship theirs under ``plans/`` and
``scripts/search_plan.py`` regenerates them (rank-parallel search, best of many
seeds -- the offline form of the reference's per-rank path vote,
``/root/reference/src/qibotn/eval.py:180-194``).
"""

import hashlib
import os

from .planner import PathInfo

PLAN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "plans")


def structure_hash(modes, out):
    h = hashlib.sha256()
    h.update(repr([list(m) for m in modes]).encode())
    h.update(repr(list(out)).encode())
    return h.hexdigest()[:16]


def plan_file(name):
    return os.path.join(PLAN_DIR, name + ".json")


def save_plan(name, info: PathInfo, modes, out):
    os.makedirs(PLAN_DIR, exist_ok=True)
    info.extra = dict(info.extra, structure=structure_hash(modes, out))
    with open(plan_file(name), "w") as f:
        f.write(info.to_json())


def load_plan(name, modes, out):
    """The stored plan, or None when absent or made for a different network."""
    path = plan_file(name)
    if not os.path.exists(path):
        return None
    info = PathInfo.from_json(open(path).read())
    if info.extra.get("structure") != structure_hash(modes, out):
        return None
    return info


_BY_STRUCTURE = None


def find_by_structure(modes, out, max_log2_width=None, min_slices=1):
    """A committed plan (``plans/*.json``) made for exactly this network structure, or None: lets the public
    entry points (``eval.amplitude_tn`` ...) pick up the plans that ``scripts/search_plan.py`` /
    ``scripts/anneal_plan.py`` produced offline instead of searching for minutes (the cheapest match wins)."""
    global _BY_STRUCTURE
    if _BY_STRUCTURE is None:
        _BY_STRUCTURE = {}
        if os.path.isdir(PLAN_DIR):
            for fn in sorted(os.listdir(PLAN_DIR)):
                if not fn.endswith(".json"):
                    continue
                try:
                    info = PathInfo.from_json(open(os.path.join(PLAN_DIR, fn)).read())
                except (OSError, ValueError, KeyError, TypeError):
                    continue
                key = info.extra.get("structure")
                if key:
                    _BY_STRUCTURE.setdefault(key, []).append(info)
    best = None
    for info in _BY_STRUCTURE.get(structure_hash(modes, out), []):
        if max_log2_width is not None and info.log2_width > max_log2_width:
            continue
        if info.num_slices < min_slices:
            continue
        if best is None or info.flops < best.flops:
            best = info
    return best
