"""Observable normalisation: runcard value -> list of Pauli-string terms.

Same entry points and error behaviour as the helpers at
``/root/reference/src/qibotn/eval.py:15-150`` (``check_observable``,
``build_observable``, ``create_hamiltonian_from_dict``, ``get_ham_gates``,
``extract_gates_and_qubits``).  A *term* is a list of
``(qubit, letter, coeff)`` with the coefficient on the first factor and 1.0 on
the rest, exactly what ``extract_gates_and_qubits`` produces.

Works with real ``qibo.hamiltonians.SymbolicHamiltonian`` objects (duck-typed
on ``.terms[*].coefficient/.factors``) and with the stand-ins of
``hamiltonians_compat`` when qibo is absent.  It also accepts the older
``{"pauli_string_pattern": "XZ"}`` dict that the README still documents
(``/root/reference/README.md:108-110``): pattern tiled cyclically over the
qubits, one term, coefficient 1.
"""

from __future__ import annotations

import numpy as np

try:  # pragma: no cover - qibo is not in this image
    from qibo import hamiltonians
    from qibo.symbols import I, X, Y, Z
except Exception:
    from . import hamiltonians_compat as hamiltonians
    from .hamiltonians_compat import I, X, Y, Z

PAULI = {
    "I": np.array([[1, 0], [0, 1]], dtype=np.complex128),
    "X": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    "Z": np.array([[1, 0], [0, -1]], dtype=np.complex128),
}


def build_observable(circuit_nqubit):
    """Default observable: sum_i 0.5 X_i Z_{(i+1) mod n} (eval.py:28-35)."""
    form = 0
    for i in range(circuit_nqubit):
        form += 0.5 * X(i % circuit_nqubit) * Z((i + 1) % circuit_nqubit)
    return hamiltonians.SymbolicHamiltonian(form=form)


def create_hamiltonian_from_dict(data, circuit_nqubit):
    """``{"terms": [{"coefficient": c, "operators": [("Z", 0), ...]}, ...]}`` ->
    SymbolicHamiltonian with an explicit I on every unnamed qubit (eval.py:38-83)."""
    paulis = {"X": X, "Y": Y, "Z": Z}
    if "terms" not in data and "pauli_string_pattern" in data:
        pattern = data["pauli_string_pattern"]
        if not pattern or any(ch not in PAULI for ch in pattern):
            raise ValueError("pauli string character must be one of I/X/Y/Z")
        ops = {"I": I, **paulis}
        expr = None
        for q in range(circuit_nqubit):
            f = ops[pattern[q % len(pattern)]](q)
            expr = f if expr is None else expr * f
        return hamiltonians.SymbolicHamiltonian(1.0 * expr)
    terms = []
    for term in data["terms"]:
        named = {q: paulis[g] for g, q in term["operators"]}      # KeyError on a bad letter
        expr = None
        for q in range(circuit_nqubit):
            f = named[q](q) if q in named else I(q)
            expr = f if expr is None else expr * f
        terms.append(term["coefficient"] * expr)
    if not terms:
        raise ValueError("No valid Hamiltonian terms were added.")
    return hamiltonians.SymbolicHamiltonian(sum(terms))


def check_observable(observable, circuit_nqubit):
    """None -> default; dict -> from dict; SymbolicHamiltonian -> as is (eval.py:15-25)."""
    if observable is None:
        return build_observable(circuit_nqubit)
    if isinstance(observable, dict):
        return create_hamiltonian_from_dict(observable, circuit_nqubit)
    if isinstance(observable, hamiltonians.SymbolicHamiltonian) or (
            hasattr(observable, "terms") and not isinstance(observable, (list, tuple))):
        return observable
    raise TypeError("Invalid observable type.")


def extract_gates_and_qubits(hamiltonian):
    """Per term: ``[(qubit, letter, coeff), ...]``, coefficient on the first factor
    only; letter and qubit parsed from ``str(factor)`` such as "X12" (eval.py:114-150)."""
    if not hasattr(hamiltonian, "terms"):
        raise ValueError("Unsupported Hamiltonian type. Must be SymbolicHamiltonian or Hamiltonian.")
    extracted = []
    for term in hamiltonian.terms:
        coeff, row = term.coefficient, []
        for factor in term.factors:
            text = str(factor)
            row.append((int(text[1:]), text[0], coeff))
            coeff = 1.0
        extracted.append(row)
    return extracted


def get_ham_gates(pauli_map, dtype="complex128"):
    """One term -> ``[(coeff * sigma, (qubit,))]`` as host arrays (eval.py:86-111)."""
    from .network import np_dtype

    dt = np_dtype(dtype)
    gates = []
    for qubit, letter, coeff in pauli_map:
        sigma = PAULI.get(letter)
        if sigma is None:
            raise ValueError("pauli string character must be one of I/X/Y/Z")
        gates.append(((coeff * sigma).astype(dt), (qubit,)))
    return gates
