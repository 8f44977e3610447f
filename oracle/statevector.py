"""Independent dense state-vector simulator (oracle; test infrastructure only).

Shares no code with the tensor-network restatement: it applies each gate
matrix to a ``(2,)*n`` array with ``numpy.tensordot``.  Qubit 0 is the most
significant bit, qibo's convention.
"""

import numpy as np


def _gate_list(circuit):
    for g in circuit.queue:
        qs = tuple(g.control_qubits) + tuple(g.target_qubits)
        try:
            m = np.asarray(g.matrix(), dtype=np.complex128)
        except NotImplementedError:      # measurement marker
            continue
        yield m, qs


def apply_matrix(state, mat, qubits, n):
    k = len(qubits)
    psi = state.reshape((2,) * n)
    g = mat.reshape((2,) * (2 * k))
    psi = np.tensordot(g, psi, axes=(list(range(k, 2 * k)), list(qubits)))
    # tensordot puts the gate's output axes first; move them back to `qubits`
    psi = np.moveaxis(psi, list(range(k)), list(qubits))
    return psi.reshape(-1)


def simulate(circuit, initial_state=None):
    n = circuit.nqubits
    if initial_state is None:
        state = np.zeros(1 << n, dtype=np.complex128)
        state[0] = 1.0
    else:
        state = np.asarray(initial_state, dtype=np.complex128).reshape(-1).copy()
    for m, qs in _gate_list(circuit):
        state = apply_matrix(state, m, qs, n)
    return state


_PAULI = {
    "I": np.eye(2, dtype=np.complex128),
    "X": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    "Z": np.array([[1, 0], [0, -1]], dtype=np.complex128),
}


def pauli_expectation(state, terms, n):
    """sum_t coeff_t <state| prod_f sigma_f |state>; ``terms`` = [(coeff, [(letter, q), ...])]."""
    total = 0.0 + 0.0j
    for coeff, factors in terms:
        phi = state
        for letter, q in factors:
            phi = apply_matrix(phi, _PAULI[letter], (q,), n)
        total += coeff * np.vdot(state, phi)
    return total
