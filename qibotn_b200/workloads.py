"""Named benchmark workloads = BASELINE.json's configurations, as (circuit, network) builders.

Used by ``bench.py``, ``scripts/`` and the tests so that all of them talk about the
same synthetic inputs (seeded)."""

from __future__ import annotations

from dataclasses import dataclass

from . import circuits
from .network import QiboCircuitToEinsum, TensorNetwork
from .observables import get_ham_gates


@dataclass
class Workload:
    name: str
    kind: str                  # "amplitude" | "expectation" | "dense" | "mps"
    dtype: str
    circuit: object
    description: str
    bitstring: str = None
    term: list = None
    observable: object = None

    def converter(self):
        return QiboCircuitToEinsum(self.circuit, self.dtype)

    def network(self) -> TensorNetwork:
        conv = self.converter()
        if self.kind == "amplitude":
            return conv.amplitude_network(self.bitstring)
        if self.kind == "expectation":
            return conv.expectation_network(get_ham_gates(self.term, self.dtype))
        if self.kind == "dense":
            return conv.state_vector_network()
        raise ValueError(f"{self.kind} workloads have no single network")


def build(name: str) -> Workload:
    """``c1`` .. ``c5`` plus reduced-size variants used by tests (``c4_d8`` etc.)."""
    base, _, opt = name.partition("_")
    if base == "c1":
        n = int(opt[1:]) if opt else 10
        return Workload(name, "dense", "complex128", circuits.qft(n, True),
                        f"{n}-qubit QFT dense state vector, complex128")
    if base == "c2":
        n, d = 30, 12
        if opt:
            n, d = (int(x) for x in opt.split("x"))
        circ = circuits.brickwork(n, d, seed=0)
        return Workload(name, "expectation", "complex64", circ,
                        f"{n}-qubit random brickwork depth {d}, <Z..Z>, complex64",
                        term=[(q, "Z", 1.0) for q in range(n)],
                        observable={"pauli_string_pattern": "Z"})
    if base == "c3":
        n, steps = 64, 20
        if opt:
            n, steps = (int(x) for x in opt.split("x"))
        return Workload(name, "mps", "complex128", circuits.tfim_c3(n, steps),
                        f"{n}-site TFIM Trotter, {steps} steps, MPS bond<=256, complex128")
    if base == "c4":
        d = int(opt[1:]) if opt else 14
        circ = circuits.sycamore_rqc(d, seed=0)
        return Workload(name, "amplitude", "complex64", circ,
                        f"53-qubit Sycamore-like RQC (ABCDCDAB, fSim(pi/2,pi/6)), depth {d}, "
                        "amplitude <0..0|U|0..0>, complex64", bitstring="0" * circ.nqubits)
    if base == "c5":
        n, p = 40, 4
        if opt:
            n, p = (int(x) for x in opt.split("x"))
        circ, edges = circuits.qaoa_maxcut(n, p, seed=0)
        return Workload(name, "expectation", "complex64", circ,
                        f"{n}-qubit QAOA MaxCut p={p} on a 3-regular graph, sum of ZZ, complex64",
                        term=[(edges[0][0], "Z", 1.0), (edges[0][1], "Z", 1.0)],
                        observable=circuits.zz_terms(edges))
    raise ValueError(f"unknown workload {name!r}")
