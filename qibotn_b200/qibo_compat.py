"""Minimal qibo-shaped circuit objects, used only when qibo itself is not importable.

The engine never depends on qibo classes: it duck-types on the handful of
attributes the reference reads from a circuit (``circuit.queue``,
``circuit.nqubits``, ``circuit.invert()``) and from a gate
(``control_qubits``, ``target_qubits``, ``matrix()``) -- see
``/root/reference/src/qibotn/circuit_convertor.py:110-122,155-167,225``.
These stand-ins expose exactly that surface, with qibo's conventions:

* ``gate.matrix()`` is the full unitary over ``control_qubits + target_qubits``
  (controls first), the first listed qubit being the most significant bit;
* ``Circuit.invert()`` reverses the queue and daggers every gate;
* measurement gates (``M``) sit in the queue but carry no matrix.

When qibo *is* installed, pass real ``qibo.models.Circuit`` objects instead;
nothing in the engine imports this module on that path.
"""

from __future__ import annotations

import cmath
import math
from typing import Iterable, Sequence

import numpy as np

__all__ = [
    "Gate", "Circuit", "H", "X", "Y", "Z", "S", "SDG", "T", "TDG", "I", "SX", "SY", "SW",
    "RX", "RY", "RZ", "U1", "U3", "CNOT", "CZ", "CU1", "CRZ", "SWAP", "iSWAP", "FSWAP",
    "RZZ", "RXX", "fSim", "Unitary", "M", "QFT", "raise_error",
]


def raise_error(exception, message=None):
    """Same contract as ``qibo.config.raise_error``: raise ``exception(message)``."""
    raise exception(message) if message is not None else exception()


class Gate:
    """A unitary on ``control_qubits + target_qubits`` with an explicit matrix."""

    name = "gate"

    def __init__(self, targets: Sequence[int], controls: Sequence[int] = (), parameters=()):
        self.target_qubits = tuple(int(q) for q in targets)
        self.control_qubits = tuple(int(q) for q in controls)
        self.parameters = tuple(parameters)
        qs = self.control_qubits + self.target_qubits
        if len(set(qs)) != len(qs):
            raise_error(ValueError, f"repeated qubit in gate {self.name}: {qs}")

    @property
    def qubits(self):
        return self.control_qubits + self.target_qubits

    # -- subclasses give the matrix on the *target* qubits only ------------
    def _target_matrix(self) -> np.ndarray:
        raise NotImplementedError

    def matrix(self, backend=None) -> np.ndarray:
        t = np.asarray(self._target_matrix(), dtype=np.complex128)
        nc = len(self.control_qubits)
        if nc == 0:
            return t
        dim_t = t.shape[0]
        full = np.eye(dim_t << nc, dtype=np.complex128)
        full[-dim_t:, -dim_t:] = t          # acts only when every control is |1>
        return full

    def dagger(self) -> "Gate":
        return Unitary(self._target_matrix().conj().T, *self.target_qubits,
                       controls=self.control_qubits, name=self.name + "_dg")

    def __repr__(self):
        return f"{self.name}{self.qubits}"


class _Fixed(Gate):
    _mat: np.ndarray = None
    _self_inverse = False

    def _target_matrix(self):
        return self._mat

    def dagger(self):
        if self._self_inverse:
            return self
        return super().dagger()


def _fixed(name, mat, nt=1, nc=0, self_inverse=False):
    mat = np.asarray(mat, dtype=np.complex128)

    def __init__(self, *qubits):
        if len(qubits) != nt + nc:
            raise_error(ValueError, f"{name} takes {nt + nc} qubits")
        Gate.__init__(self, qubits[nc:], qubits[:nc])

    return type(name, (_Fixed,), {"name": name.lower(), "_mat": mat, "__init__": __init__,
                                  "_self_inverse": self_inverse})


_s2 = 1.0 / math.sqrt(2.0)
I = _fixed("I", np.eye(2), self_inverse=True)
H = _fixed("H", [[_s2, _s2], [_s2, -_s2]], self_inverse=True)
X = _fixed("X", [[0, 1], [1, 0]], self_inverse=True)
Y = _fixed("Y", [[0, -1j], [1j, 0]], self_inverse=True)
Z = _fixed("Z", [[1, 0], [0, -1]], self_inverse=True)
S = _fixed("S", [[1, 0], [0, 1j]])
SDG = _fixed("SDG", [[1, 0], [0, -1j]])
T = _fixed("T", [[1, 0], [0, cmath.exp(1j * math.pi / 4)]])
TDG = _fixed("TDG", [[1, 0], [0, cmath.exp(-1j * math.pi / 4)]])
# sqrt(X), sqrt(Y), sqrt(W) with W = (X + Y)/sqrt(2): the one-qubit set of the
# Sycamore random circuits.
SX = _fixed("SX", 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]))
SY = _fixed("SY", 0.5 * np.array([[1 + 1j, -1 - 1j], [1 + 1j, 1 + 1j]]))
SW = _fixed("SW", _s2 * np.array([[1, -cmath.sqrt(1j)], [cmath.sqrt(-1j), 1]]))
CNOT = _fixed("CNOT", [[0, 1], [1, 0]], nt=1, nc=1, self_inverse=True)
CZ = _fixed("CZ", [[1, 0], [0, -1]], nt=1, nc=1, self_inverse=True)
SWAP = _fixed("SWAP", [[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], nt=2,
              self_inverse=True)
iSWAP = _fixed("iSWAP", [[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], nt=2)
FSWAP = _fixed("FSWAP", [[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, -1]], nt=2,
               self_inverse=True)


class _Param(Gate):
    _ntargets = 1

    def __init__(self, *args, **kw):
        nt, nc = self._ntargets, getattr(self, "_ncontrols", 0)
        names = self._pnames
        qubits = list(args[: nt + nc])
        params = list(args[nt + nc:])
        for k in ("q", "q0", "q1"):
            if k in kw:
                qubits.append(kw.pop(k))
        for n in names[len(params):]:
            params.append(kw.pop(n))
        if kw or len(qubits) != nt + nc or len(params) != len(names):
            raise_error(TypeError, f"bad arguments for {self.name}")
        super().__init__(qubits[nc:], qubits[:nc], [float(p) for p in params])

    def dagger(self):
        return type(self)(*self.qubits, *[-p for p in self.parameters])


class RX(_Param):
    name, _pnames = "rx", ("theta",)

    def _target_matrix(self):
        c, s = math.cos(self.parameters[0] / 2), math.sin(self.parameters[0] / 2)
        return np.array([[c, -1j * s], [-1j * s, c]])


class RY(_Param):
    name, _pnames = "ry", ("theta",)

    def _target_matrix(self):
        c, s = math.cos(self.parameters[0] / 2), math.sin(self.parameters[0] / 2)
        return np.array([[c, -s], [s, c]], dtype=np.complex128)


class RZ(_Param):
    name, _pnames = "rz", ("theta",)

    def _target_matrix(self):
        e = cmath.exp(0.5j * self.parameters[0])
        return np.array([[e.conjugate(), 0], [0, e]])


class U1(_Param):
    name, _pnames = "u1", ("theta",)

    def _target_matrix(self):
        return np.array([[1, 0], [0, cmath.exp(1j * self.parameters[0])]])


class U3(_Param):
    name, _pnames = "u3", ("theta", "phi", "lam")

    def _target_matrix(self):
        th, ph, la = self.parameters
        c, s = math.cos(th / 2), math.sin(th / 2)
        ep, em = cmath.exp(0.5j * (ph + la)), cmath.exp(0.5j * (ph - la))
        return np.array([[ep.conjugate() * c, -em.conjugate() * s], [em * s, ep * c]])

    def dagger(self):
        th, ph, la = self.parameters
        return U3(*self.qubits, -th, -la, -ph)


class CU1(U1):
    name, _ncontrols = "cu1", 1


class CRZ(RZ):
    name, _ncontrols = "crz", 1


class RZZ(_Param):
    name, _pnames, _ntargets = "rzz", ("theta",), 2

    def _target_matrix(self):
        e = cmath.exp(0.5j * self.parameters[0])
        return np.diag([e.conjugate(), e, e, e.conjugate()])


class RXX(_Param):
    name, _pnames, _ntargets = "rxx", ("theta",), 2

    def _target_matrix(self):
        c, s = math.cos(self.parameters[0] / 2), -1j * math.sin(self.parameters[0] / 2)
        return np.array([[c, 0, 0, s], [0, c, s, 0], [0, s, c, 0], [s, 0, 0, c]])


class fSim(_Param):
    name, _pnames, _ntargets = "fsim", ("theta", "phi"), 2

    def _target_matrix(self):
        th, ph = self.parameters
        c, s = math.cos(th), -1j * math.sin(th)
        return np.array([[1, 0, 0, 0], [0, c, s, 0], [0, s, c, 0], [0, 0, 0, cmath.exp(-1j * ph)]])


class Unitary(Gate):
    def __init__(self, unitary, *qubits, controls=(), name="unitary"):
        self._u = np.asarray(unitary, dtype=np.complex128)
        self.name = name
        super().__init__(qubits, controls)
        if self._u.shape != (1 << len(self.target_qubits),) * 2:
            raise_error(ValueError, "unitary shape does not match the target qubits")

    def _target_matrix(self):
        return self._u


class M(Gate):
    """Measurement marker. It has no matrix; the engine skips it."""

    name = "measure"

    def __init__(self, *qubits):
        super().__init__(qubits)

    def matrix(self, backend=None):
        raise_error(NotImplementedError, "measurement gates have no matrix")

    def dagger(self):
        return self


class Circuit:
    def __init__(self, nqubits: int):
        self.nqubits = int(nqubits)
        self.queue: list[Gate] = []

    def add(self, gate):
        if isinstance(gate, Gate):
            if any(q < 0 or q >= self.nqubits for q in gate.qubits):
                raise_error(ValueError, f"gate {gate} outside a {self.nqubits}-qubit circuit")
            self.queue.append(gate)
        elif isinstance(gate, Iterable):
            for g in gate:
                self.add(g)
        else:
            raise_error(TypeError, f"cannot add {type(gate)} to a circuit")
        return self

    def invert(self) -> "Circuit":
        inv = Circuit(self.nqubits)
        for g in reversed(self.queue):
            if isinstance(g, M):
                continue
            inv.queue.append(g.dagger())
        return inv

    def to_qasm(self) -> str:
        """OpenQASM 2.0 text as qibo's ``Circuit.to_qasm()`` writes it (one ``qreg q``, qelib1 names, parameters as
        ``repr(float)``, qubits in ``control, target`` order): the wire format of the reference's CPU paths
        (eval_qu.py:39-41, backends/quimb.py:169-171).  Gates without a qelib1 name raise."""
        names = {"i": "id", "cnot": "cx"}
        plain = {"id", "h", "x", "y", "z", "s", "sdg", "t", "tdg", "sx", "cx", "cz", "swap"}
        rotations = {"rx", "ry", "rz", "u1", "cu1", "crz", "rzz", "rxx"}
        lines = ["// Generated by qibotn_b200.qibo_compat", "OPENQASM 2.0;", 'include "qelib1.inc";',
                 f"qreg q[{self.nqubits}];"]
        measured = [q for g in self.queue if isinstance(g, M) for q in g.target_qubits]
        if measured:
            lines.append(f"creg c[{len(measured)}];")
        slot = 0
        for g in self.queue:
            if isinstance(g, M):
                for q in g.target_qubits:
                    lines.append(f"measure q[{q}] -> c[{slot}];")
                    slot += 1
                continue
            name = names.get(g.name, g.name)
            args = ",".join(f"q[{q}]" for q in g.qubits)
            if name in plain and not g.parameters:
                lines.append(f"{name} {args};")
            elif name in rotations and len(g.parameters) == 1:
                lines.append(f"{name}({float(g.parameters[0])!r}) {args};")
            else:
                raise_error(NotImplementedError, f"gate {g.name} has no OpenQASM 2 (qelib1) form here")
        return "\n".join(lines) + "\n"

    @property
    def ngates(self):
        return len(self.queue)

    @property
    def depth(self):
        front = [0] * self.nqubits
        for g in self.queue:
            d = 1 + max(front[q] for q in g.qubits)
            for q in g.qubits:
                front[q] = d
        return max(front, default=0)


def QFT(nqubits: int, with_swaps: bool = True) -> Circuit:
    """qibo.models.QFT: H(i), then CU1(control=j, target=i, pi/2^(j-i)) for j>i, then swaps."""
    c = Circuit(nqubits)
    for i1 in range(nqubits):
        c.add(H(i1))
        for i2 in range(i1 + 1, nqubits):
            c.add(CU1(i2, i1, math.pi / 2 ** (i2 - i1)))
    if with_swaps:
        for i in range(nqubits // 2):
            c.add(SWAP(i, nqubits - 1 - i))
    return c
