"""Seeded synthetic circuits for the five BASELINE.json configurations.

They are built from real ``qibo`` gates when qibo is importable and from
``qibotn_b200.qibo_compat`` otherwise; either way the engine only reads
``queue``/``nqubits``/``invert()`` and per-gate ``control_qubits``,
``target_qubits``, ``matrix()``.  Readings of the configs follow SURVEY.md
section 8(a); every choice that changes cost is stated here.
"""

from __future__ import annotations

import math

import numpy as np

try:  # pragma: no cover - qibo is not in this image
    from qibo import Circuit, gates as _g
    from qibo.models import QFT as _QFT

    def _unitary(u, *q, name="unitary"):
        return _g.Unitary(u, *q)

    HAVE_QIBO = True
except Exception:  # ModuleNotFoundError and friends
    from . import qibo_compat as _g
    from .qibo_compat import Circuit

    _QFT = _g.QFT

    def _unitary(u, *q, name="unitary"):
        return _g.Unitary(u, *q, name=name)

    HAVE_QIBO = False

gates = _g

_SQRT_X = 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]])
_SQRT_Y = 0.5 * np.array([[1 + 1j, -1 - 1j], [1 + 1j, 1 + 1j]])
_SQRT_W = (1 / math.sqrt(2)) * np.array([[1, -np.sqrt(1j)], [np.sqrt(-1j), 1]])


def qft(nqubits, swaps=True):
    """C1: ``qibo.models.QFT(n, with_swaps)`` -- n H, n(n-1)/2 CU1, n//2 SWAP."""
    return _QFT(nqubits, swaps)


def ry_rz_cnot_ring(nqubits, nlayers, seed=42):
    """The layered test circuit of ``/root/reference/tests/test_expectation.py:19-31``
    (python ``random.seed(seed)``; RY then RZ per qubit, then a CNOT ring, M at the end)."""
    import random

    random.seed(seed)
    c = Circuit(nqubits)
    for _ in range(nlayers):
        for q in range(nqubits):
            c.add(gates.RY(q, theta=random.uniform(-math.pi, math.pi)))
            c.add(gates.RZ(q, theta=random.uniform(-math.pi, math.pi)))
        if nqubits > 1:
            for q in range(nqubits):
                if q % nqubits != (q + 1) % nqubits:
                    c.add(gates.CNOT(q % nqubits, (q + 1) % nqubits))
    c.add(gates.M(*range(nqubits)))
    return c


def brickwork(nqubits, depth, seed=0, two_qubit="random"):
    """C2: ``depth`` layers, each = one random SU(2)-ish gate (RZ.RY.RZ as a U3) on
    every qubit followed by two-qubit gates on bonds (q, q+1) with q = layer parity,
    parity+2, ... .  ``two_qubit`` = "random" (Haar-ish 4x4 unitary from a seeded QR),
    "cz" or "cnot"."""
    rng = np.random.default_rng(seed)
    c = Circuit(nqubits)
    for layer in range(depth):
        for q in range(nqubits):
            th, ph, la = rng.uniform(-math.pi, math.pi, 3)
            c.add(gates.U3(q, th, ph, la))
        for q in range(layer % 2, nqubits - 1, 2):
            if two_qubit == "random":
                z = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
                qm, r = np.linalg.qr(z)
                qm = qm * (np.diag(r) / np.abs(np.diag(r)))
                c.add(_unitary(qm, q, q + 1, name="su4"))
            elif two_qubit == "cz":
                c.add(gates.CZ(q, q + 1))
            else:
                c.add(gates.CNOT(q, q + 1))
    return c


def tfim_trotter(nqubits, steps, j_dt=0.4, h_dt=0.6):
    """C3: first-order Trotter steps of the 1D transverse-field Ising chain:
    RZZ(2 J dt) on even bonds, RZZ on odd bonds, RX(2 h dt) on every site.
    ``j_dt``/``h_dt`` are large on purpose so that bulk bonds really grow."""
    c = Circuit(nqubits)
    for q in range(nqubits):
        c.add(gates.H(q))
    for _ in range(steps):
        for par in (0, 1):
            for q in range(par, nqubits - 1, 2):
                c.add(gates.RZZ(q, q + 1, theta=2 * j_dt))
        for q in range(nqubits):
            c.add(gates.RX(q, theta=2 * h_dt))
    return c


# BASELINE config 3 (64-site TFIM Trotter brickwork, 20 steps, bond <= 256).  Angles chosen by a scan on
# the engine (scripts/c3_scan.py, profiles/r02_c3_angle_scan.jsonl): every bulk bond reaches 256 AND the
# truncated state keeps <psi|psi> = 0.914 (8.6 % discarded weight) -- the round-1 angles (0.4, 0.6) left a
# state of norm 9e-6, i.e. nothing physical to check at full size.
C3_ANGLES = {"j_dt": 0.2, "h_dt": 0.3}


def tfim_c3(nqubits=64, steps=20):
    return tfim_trotter(nqubits, steps, **C3_ANGLES)


# --------------------------------------------------------------- Sycamore-like
def sycamore_qubits(drop=((3, 2),)):
    """The 54-site diamond of Google's Sycamore as (row, col) grid points, minus
    ``drop`` (one degree-2 corner site by default -> 53 qubits, 86 couplers)."""
    rows = {0: (5, 6), 1: (4, 7), 2: (3, 8), 3: (2, 9), 4: (1, 9), 5: (0, 8), 6: (1, 7),
            7: (2, 6), 8: (3, 5), 9: (4, 4)}
    sites = [(r, c) for r, (lo, hi) in rows.items() for c in range(lo, hi + 1)]
    return [s for s in sites if s not in set(drop)]


def sycamore_couplers(sites):
    """Nearest-neighbour pairs of the grid, split in the four staggered classes A-D.

    A pair is oriented so that the second site is one step down (vertical) or one
    step right (horizontal).  With (r, c) the first site of a vertical pair and
    (c, r) read for a horizontal one, the pair is in the class with column offset
    ``o`` when ``(r % 2, (c - o) % 2)`` is (0,0) or (1,1).  A, B = vertical with
    o = 0, 1; C, D = horizontal with o = 1, 0.  This is the staggered 2x2 unit
    cell of the supremacy experiment; the class *sequence* is what makes an
    instance easy or hard (see ``sycamore_rqc``).
    """
    s = set(sites)
    cls = {"A": [], "B": [], "C": [], "D": []}
    for (r, c) in sites:
        if (r + 1, c) in s:       # vertical pair: transpose coordinates
            a_row, a_col = c, r
            for name, off in (("A", 0), ("B", 1)):
                if (a_row % 2, (a_col - off) % 2) in ((0, 0), (1, 1)):
                    cls[name].append(((r, c), (r + 1, c)))
        if (r, c + 1) in s:       # horizontal pair
            a_row, a_col = r, c
            for name, off in (("C", 1), ("D", 0)):
                if (a_row % 2, (a_col - off) % 2) in ((0, 0), (1, 1)):
                    cls[name].append(((r, c), (r, c + 1)))
    return cls


def sycamore_rqc(depth=14, seed=0, sequence="ABCDCDAB", theta=math.pi / 2, phi=math.pi / 6):
    """C4: 53-qubit Sycamore-like random circuit.

    Each of the ``depth`` cycles = one gate from {sqrt X, sqrt Y, sqrt W} on every
    qubit (never the same gate twice in a row on a qubit) followed by
    fSim(theta, phi) on every coupler of the cycle's class; classes follow
    ``sequence`` cyclically ("ABCDCDAB" = the supremacy ordering, the hard one;
    "ABCD" would be the easier 'simplifiable' ordering).  A final one-qubit layer
    closes the circuit.  fSim(pi/2, pi/6) is a dense rank-4 two-qubit tensor.
    """
    sites = sycamore_qubits()
    index = {s: i for i, s in enumerate(sorted(sites))}
    classes = sycamore_couplers(sites)
    rng = np.random.default_rng(seed)
    n = len(sites)
    c = Circuit(n)
    one = [("sx", _SQRT_X), ("sy", _SQRT_Y), ("sw", _SQRT_W)]
    last = [-1] * n

    def one_qubit_layer():
        for q in range(n):
            k = int(rng.integers(0, 3))
            while k == last[q]:
                k = int(rng.integers(0, 3))
            last[q] = k
            c.add(_unitary(one[k][1], q, name=one[k][0]))

    for cyc in range(depth):
        one_qubit_layer()
        for a, b in classes[sequence[cyc % len(sequence)]]:
            c.add(gates.fSim(index[a], index[b], theta, phi) if hasattr(gates, "fSim")
                  else _unitary(_fsim(theta, phi), index[a], index[b], name="fsim"))
    one_qubit_layer()
    return c


def _fsim(theta, phi):
    cth, sth = math.cos(theta), -1j * math.sin(theta)
    return np.array([[1, 0, 0, 0], [0, cth, sth, 0], [0, sth, cth, 0],
                     [0, 0, 0, np.exp(-1j * phi)]])


def qaoa_maxcut(nqubits=40, p=4, degree=3, seed=0):
    """C5: QAOA MaxCut on a random ``degree``-regular graph: H on all, then p rounds
    of RZZ(gamma) per edge and RX(beta) per node.  Returns (circuit, edges)."""
    import networkx as nx

    graph = nx.random_regular_graph(degree, nqubits, seed=seed)
    edges = sorted(tuple(sorted(e)) for e in graph.edges())
    rng = np.random.default_rng(seed)
    c = Circuit(nqubits)
    for q in range(nqubits):
        c.add(gates.H(q))
    for _ in range(p):
        gamma, beta = rng.uniform(-math.pi, math.pi, 2)
        for a, b in edges:
            c.add(gates.RZZ(a, b, theta=gamma))
        for q in range(nqubits):
            c.add(gates.RX(q, theta=2 * beta))
    return c, edges


def zz_terms(edges, coefficient=1.0):
    """Observable dict (the ``{"terms": ...}`` runcard form) = sum of Z_a Z_b over edges."""
    return {"terms": [{"coefficient": coefficient, "operators": [("Z", a), ("Z", b)]}
                      for a, b in edges]}
