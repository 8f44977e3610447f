"""Circuit -> tensor-network conversion (host side).

Drop-in for ``QiboCircuitToEinsum`` of
``/root/reference/src/qibotn/circuit_convertor.py:7-246``: same constructor,
same ``state_vector_operands()`` / ``expectation_operands(ham_gates)`` return
convention (the *interleaved* list ``[T0, modes0, T1, modes1, ..., out]``),
same attributes (``gate_tensors``, ``active_qubits``, ``circuit``).

Differences, all deliberate:

* operands are host ``numpy`` arrays; the executor ships every leaf to the GPU
  in ONE packed copy instead of one ``cp.asarray`` per gate (K10);
* one dtype throughout -- the reference leaves inverse-circuit and Pauli
  tensors in complex128 even for a complex64 run (quirk 2);
* measurement markers are skipped (quirk 5);
* ``amplitude_operands(bitstring)`` closes the network with <b| vectors, the
  single-amplitude mode BASELINE config 4 needs and the reference lacks;
* the inverse circuit is converted once and cached, not once per term.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_NP_DTYPES = {"complex64": np.complex64, "complex128": np.complex128}


def np_dtype(dtype):
    if isinstance(dtype, str):
        try:
            return _NP_DTYPES[dtype]
        except KeyError:
            raise TypeError(f"unsupported dtype {dtype!r}; use complex64 or complex128")
    dt = np.dtype(dtype).type
    if dt not in (np.complex64, np.complex128):
        raise TypeError(f"unsupported dtype {dtype!r}; use complex64 or complex128")
    return dt


@dataclass
class TensorNetwork:
    """Leaves (host arrays, every extent 2), their labels, and the output labels."""

    tensors: list
    modes: list
    out: list

    @classmethod
    def from_interleaved(cls, inter):
        if len(inter) % 2 == 1:
            body, out = inter[:-1], list(inter[-1])
        else:
            body, out = inter, []
        return cls(list(body[0::2]), [list(m) for m in body[1::2]], out)

    def interleaved(self):
        flat = [x for pair in zip(self.tensors, self.modes) for x in pair]
        flat.append(list(self.out))
        return flat


class QiboCircuitToEinsum:
    def __init__(self, circuit, dtype="complex128"):
        self.dtype = np_dtype(dtype)
        self.circuit = circuit
        self.basis_map = {"0": np.array([1, 0], dtype=self.dtype),
                          "1": np.array([0, 1], dtype=self.dtype)}
        self.gate_tensors, self.active_qubits = self._convert(circuit)
        self._inverse = None

    # ------------------------------------------------------------ conversion
    def _convert(self, circuit):
        """(tensor, control_qubits + target_qubits) per gate; reference
        ``init_intermediate_circuit`` (circuit_convertor.py:98-125)."""
        pairs, touched = [], []
        for gate in circuit.queue:
            qubits = tuple(gate.control_qubits) + tuple(gate.target_qubits)
            if _is_measurement(gate):
                continue
            mat = np.asarray(gate.matrix(), dtype=self.dtype)
            pairs.append((np.ascontiguousarray(mat).reshape((2, 2) * len(qubits)), qubits))
            touched.extend(qubits)
        return pairs, np.unique(np.asarray(touched, dtype=np.int64))

    @property
    def gate_tensors_inverse(self):
        """Tensors of ``circuit.invert()`` (reference ``init_inverse_circuit``, :143-170)."""
        if self._inverse is None:
            self._inverse, _ = self._convert(self.circuit.invert())
        return self._inverse

    @staticmethod
    def _thread(gates, frontier, next_label, hyper=False):
        """Advance the per-qubit frontier through ``gates``; gate modes are the new
        output labels followed by the labels they consume (:69-85).

        ``hyper``: a gate whose matrix is diagonal in the computational basis (RZ, CZ, CU1, RZZ, ...)
        neither creates nor consumes labels -- it is emitted as its diagonal, a rank-k tensor on the k
        labels the qubits carry at that point, which thereby become hyper-indices (owned by more than
        two tensors).  Same value, far smaller treewidth for circuits such as QAOA whose two-qubit
        gates are all diagonal (the reference always emits the full rank-2k tensor)."""
        tensors, modes = [], []
        for tensor, qubits in gates:
            consumed = [frontier[q] for q in qubits]
            if hyper:
                k = len(qubits)
                mat = np.asarray(tensor).reshape(1 << k, 1 << k)
                if np.count_nonzero(mat - np.diag(np.diagonal(mat))) == 0:
                    tensors.append(np.ascontiguousarray(np.diagonal(mat)).reshape((2,) * k))
                    modes.append(consumed)
                    continue
            produced = list(range(next_label, next_label + len(qubits)))
            next_label += len(qubits)
            frontier.update(zip(qubits, produced))
            tensors.append(tensor)
            modes.append(produced + consumed)
        return tensors, modes, next_label

    # ---------------------------------------------------------------- outputs
    def _open_network(self):
        active = [int(q) for q in self.active_qubits]
        frontier = {q: i for i, q in enumerate(active)}
        tensors = [self.basis_map["0"]] * len(active)
        modes = [[i] for i in range(len(active))]
        g_t, g_m, _ = self._thread(self.gate_tensors, frontier, len(active))
        return active, frontier, tensors + g_t, modes + g_m

    def state_vector_network(self, initial_state=None) -> TensorNetwork:
        """``initial_state``: dense vector over ALL qubits (qubit 0 most significant) that replaces |0...0>
        (the legacy quimb path, eval_qu.py:34-41).  It enters as ONE rank-n leaf -- the reference splits it into an
        MPS first (``from_dense``), which changes the factorisation of the input, not the contracted value -- and
        every qubit is active then."""
        if initial_state is None:
            active, frontier, tensors, modes = self._open_network()
            return TensorNetwork(tensors, modes, [frontier[q] for q in active])
        n = int(self.circuit.nqubits)
        psi = np.ascontiguousarray(np.asarray(initial_state, dtype=self.dtype).reshape(-1))
        if psi.size != 1 << n:
            raise ValueError(f"initial state has {psi.size} amplitudes, expected 2^{n}")
        frontier = {q: q for q in range(n)}
        g_t, g_m, _ = self._thread(self.gate_tensors, frontier, n)
        return TensorNetwork([psi.reshape((2,) * n)] + g_t, [list(range(n))] + g_m, [frontier[q] for q in range(n)])

    def state_vector_operands(self):
        """Dense state vector over the *active* qubits (:28-58)."""
        return self.state_vector_network().interleaved()

    def amplitude_network(self, bitstring) -> TensorNetwork:
        active, frontier, tensors, modes = self._open_network()
        bits = [str(b) for b in bitstring]
        if len(bits) == self.circuit.nqubits and len(active) != len(bits):
            bits = [bits[q] for q in active]
        if len(bits) != len(active) or any(b not in self.basis_map for b in bits):
            raise ValueError("bitstring must have one 0/1 character per active qubit")
        tensors = tensors + [self.basis_map[b] for b in bits]
        modes = modes + [[frontier[q]] for q in active]
        return TensorNetwork(tensors, modes, [])

    def amplitude_operands(self, bitstring):
        """<bitstring| U |0...0> as a closed network (scalar output)."""
        return self.amplitude_network(bitstring).interleaved()

    def _light_cone(self, ham):
        """Indices of the circuit gates inside the reverse light cone of the observable: walking the
        circuit backwards, a gate is kept iff it touches a qubit on which the operator accumulated so
        far acts non-trivially.  Every other gate G meets its own inverse in <0|U^dag P U|0> with
        nothing in between, and G^dag G = 1 exactly -- dropping the pair changes no value."""
        live = set()
        for t, qubits in ham:
            if len(qubits) != 1 or not np.allclose(t, t[0, 0] * np.eye(2)):
                live.update(qubits)            # anything but a multiple of the identity
        keep = []
        for idx in range(len(self.gate_tensors) - 1, -1, -1):
            qubits = self.gate_tensors[idx][1]
            if live.intersection(qubits):
                live.update(qubits)
                keep.append(idx)
        return sorted(keep)

    def expectation_network(self, ham_gates, light_cone=True, hyper=False) -> TensorNetwork:
        n = self.circuit.nqubits
        frontier = {q: q for q in range(n)}
        kets = [self.basis_map["0"]] * n
        tensors, modes = list(kets), [[q] for q in range(n)]
        ham = [(np.asarray(t, dtype=self.dtype), tuple(q)) for t, q in ham_gates]
        forward, inverse = self.gate_tensors, self.gate_tensors_inverse
        if light_cone and len(inverse) == len(forward):
            keep = self._light_cone(ham)
            if len(keep) < len(forward):
                total = len(forward)
                forward = [forward[i] for i in keep]
                # circuit.invert() lists the daggered gates in reverse order
                inverse = [inverse[total - 1 - i] for i in reversed(keep)]
        g_t, g_m, nxt = self._thread(forward, frontier, n, hyper)
        b_t, b_m, _ = self._thread(ham + inverse, frontier, nxt, hyper)
        tensors += g_t + b_t + kets
        modes += g_m + b_m + [[frontier[q]] for q in range(n)]
        return TensorNetwork(tensors, modes, [])

    def expectation_operands(self, ham_gates):
        """<0| U^dagger (prod of ham_gates) U |0> over all ``nqubits`` (:198-246).
        Returned without a trailing output list, as the reference does, and with every gate of
        the circuit and of its inverse present, as the reference does (the engine's own drivers use
        ``expectation_network``, which drops the pairs outside the observable's light cone)."""
        return self.expectation_network(ham_gates, light_cone=False).interleaved()[:-1]

    def get_pauli_gates(self, pauli_map, dtype=None):
        """{qubit: letter} -> [(sigma, (qubit,))]; kept for the older
        ``pauli_string_pattern`` flow (:172-196)."""
        from .observables import PAULI

        dt = self.dtype if dtype is None else np_dtype(dtype)
        gates = []
        for qubit, letter in pauli_map.items():
            if letter not in PAULI:
                raise ValueError("pauli string character must be one of I/X/Y/Z")
            gates.append((PAULI[letter].astype(dt), (qubit,)))
        return gates


def _is_measurement(gate):
    name = getattr(gate, "name", "")
    return name in ("measure", "M") or type(gate).__name__ == "M"
