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
