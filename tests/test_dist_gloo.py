"""N > 1 host logic on the CPU: world_size-2 and -3 gloo jobs (one process per rank) run the
reference's distributed recipe -- independent path search, cost vote, contiguous slice ranges,
one reduction of the partials (reference eval.py:153-235)."""
import json
import os
import socket
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_dist_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        try:
            o, e = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        assert p.returncode == 0, e[-2000:]
        line = [l for l in o.splitlines() if l.startswith("RESULT ")][-1]
        outs.append(json.loads(line[len("RESULT "):]))
    return sorted(outs, key=lambda r: r["rank"])


@pytest.mark.parametrize("world", [2, 3])
def test_sliced_contraction_over_ranks(world):
    outs = _run(world)
    costs = outs[0]["costs"]
    winner = min(range(world), key=lambda r: (costs[r], r))               # MINLOC: lowest rank wins ties
    ns = outs[0]["num_slices"]
    covered = []
    for r in outs:
        assert r["size"] == world and r["winner"] == winner and r["num_slices"] == ns
        assert abs(r["best_cost"] - costs[winner]) < 1e-9
        covered += list(range(*r["slices"]))
        assert r["bad"] and "Unsupported dtype" in r["bad"]
        tri = world * (world + 1) // 2
        assert abs(r["f32"][0] - tri) < 1e-6 and abs(r["f32"][1] + tri) < 1e-6
    assert covered == list(range(ns))
    want = complex(*outs[0]["want"])
    for r in outs:                                                        # all-reduce: every rank has the sum
        assert abs(complex(*r["sum"]["NCCL"]) - want) < 1e-12
    assert abs(complex(*outs[0]["sum"]["MPI"]) - want) < 1e-12             # reduce: root has the sum
    # Hamiltonian: every (term, slice) unit is owned by exactly one rank and the reduced sum is right
    counts = outs[0]["unit_counts"]
    owned = sorted((int(t), s) for r in outs for t, (a, b) in r["units"].items() for s in range(a, b))
    assert owned == [(t, s) for t, c in enumerate(counts) for s in range(c)]
    for r in outs:
        assert r["unit_counts"] == counts
        assert abs(r["ham"][0] - r["ham"][1]) < 1e-10

